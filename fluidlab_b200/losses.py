"""Step losses that seed the backward pass.  Mirrors fluidlab/fluidengine/losses/loss.py (`Loss` :13-78),
losses/shapematching_loss.py (`ShapeMatchingLoss` :13-130: index-matched squared distance, temporal range
curriculum) and losses/latteart_loss.py (`LatteArtLoss`).  Targets stay resident in HBM (the reference
re-uploads the step's target every step, shapematching_loss.py:72-78); the per-step loss and its seed are one
kernel each (fmpm_loss_chamfer / fmpm_loss_chamfer_grad)."""
import pickle as pkl
import numpy as np
import torch
from .macros import MILK, ICECREAM, ICECREAM1


class Loss:
    def __init__(self, max_loss_steps, weights=None, target_file=None, target=None):
        self.weights = weights
        self.target_file = target_file
        self.target = target
        self.inf = 1e8
        self.max_loss_steps = max_loss_steps

    def build(self, sim):
        self.sim = sim
        self.res, self.n_grid, self.dx, self.dim = sim.res, sim.n_grid, sim.dx, sim.dim
        self.agent = sim.agent
        self.n_particles = sim.n_particles
        self.step_loss = torch.zeros((self.max_loss_steps,), dtype=torch.float32, device=sim.device)
        self._step_grad_on = np.zeros(self.max_loss_steps, dtype=bool)
        if self.target_file is not None:
            self.load_target(self.target_file)
        elif self.target is not None:
            self.set_target(self.target)
        self.reset()

    def reset_grad(self):  # loss.py:49-51 (total_loss.grad = 1)
        self._step_grad_on[:] = False

    def load_target(self, path):
        pass

    def clear_loss(self):
        self.step_loss.zero_()
        self.total_loss = 0.0
        self._step_grad_on[:] = False

    def reset(self):
        self.clear_loss()

    def step(self):  # loss.py:72-74
        self.compute_step_loss(self.sim.cur_step_global - 1, self.sim.cur_substep_local)

    def step_grad(self):  # loss.py:76-78
        self.compute_step_loss_grad(self.sim.cur_step_global - 1, self.sim.cur_substep_local)


class ShapeMatchingLoss(Loss):
    def __init__(self, matching_mat, temporal_range_type='expand', temporal_init_range_end=50, plateau_count_limit=5,
                 temporal_expand_speed=50, plateau_thresh=(0.01, 0.5), **kwargs):
        super().__init__(**kwargs)
        self.matching_mat = matching_mat
        self.temporal_range_type = temporal_range_type
        self.temporal_init_range_end = temporal_init_range_end
        self.plateau_count_limit = plateau_count_limit
        self.temporal_expand_speed = temporal_expand_speed
        self.plateau_thresh = list(plateau_thresh)

    def build(self, sim):
        self.chamfer_weight = self.weights['chamfer']
        if self.temporal_range_type == 'last':
            self.temporal_range = [self.max_loss_steps - 1, self.max_loss_steps]
        elif self.temporal_range_type == 'all':
            self.temporal_range = [0, self.max_loss_steps]
        elif self.temporal_range_type == 'expand':
            self.temporal_range = [0, self.temporal_init_range_end]
            self.best_loss = self.inf
            self.plateau_count = 0
        super().build(sim)
        self.row_mask = sim.material_row_mask(self.matching_mat)

    def load_target(self, path):  # shapematching_loss.py:52-57
        target = pkl.load(open(path, 'rb'))
        self.set_target(target['x'])

    def set_target(self, xs):
        """xs: sequence of max_loss_steps arrays (N,3), original particle order."""
        assert self.max_loss_steps == len(xs)
        assert self.n_particles == len(xs[0])
        self.tgt = torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(a, dtype=np.float32) for a in xs]))).to(self.sim.device)

    def compute_step_loss(self, s, f):  # shapematching_loss.py:64-66, 80-88
        self.sim.chamfer_loss(self.tgt[s], self.row_mask, self.chamfer_weight, self.step_loss[s:s + 1], f)

    def compute_step_loss_grad(self, s, f):  # shapematching_loss.py:68-70
        if self._step_grad_on[s]:
            self.sim.add_x_grad_chamfer(self.tgt[s], self.row_mask, self.chamfer_weight, f)

    def _total(self):
        return float(self.step_loss[self.temporal_range[0]:self.temporal_range[1]].sum().item())

    def get_final_loss(self):  # shapematching_loss.py:95-106
        self.total_loss = self._total()
        self.expand_temporal_range()
        return {'loss': self.total_loss, 'last_step_loss': float(self.step_loss[self.max_loss_steps - 1].item()),
                'temporal_range': self.temporal_range[1]}

    def get_final_loss_grad(self):  # shapematching_loss.py:107-108
        self._step_grad_on[:] = False
        self._step_grad_on[self.temporal_range[0]:self.temporal_range[1]] = True

    def expand_temporal_range(self):  # shapematching_loss.py:110-130
        if self.temporal_range_type == 'expand':
            loss_improved = self.best_loss - self.total_loss
            loss_improved_rate = loss_improved / self.best_loss
            if loss_improved_rate < self.plateau_thresh[0] or loss_improved < self.plateau_thresh[1]:
                self.plateau_count += 1
            else:
                self.plateau_count = 0
            if self.best_loss > self.total_loss:
                self.best_loss = self.total_loss
            if self.plateau_count >= self.plateau_count_limit:
                self.plateau_count = 0
                self.best_loss = self.inf
                self.temporal_range[1] = min(self.max_loss_steps, self.temporal_range[1] + self.temporal_expand_speed)


class LatteArtLoss(ShapeMatchingLoss):
    def __init__(self, type='diff', **kwargs):
        super().__init__(matching_mat=MILK, temporal_range_type='all', **kwargs)

    def get_step_loss(self):  # latteart_loss.py:25-33
        cur = float(self.step_loss[self.sim.cur_step_global - 1].item())
        return {'reward': 0.025 * (121.3 - cur), 'loss': 0.025 * cur}

    def get_final_loss(self):  # latteart_loss.py:35-45
        info = super().get_final_loss()
        info['reward'] = float(np.sum((121.3 - self.step_loss.cpu().numpy()) * 0.025))
        return info


class _ScaledShapeLoss(ShapeMatchingLoss):
    """ShapeMatchingLoss with a task's material, curriculum start and reward scaling (the reference repeats the class per task: the data differ, the
    kernels do not).  `type='diff'`: the expanding temporal range used by the gradient-based solver; `'default'`: the whole horizon."""
    MAT, INIT_END, OFFSET, SCALE, STEP_LOSS_SCALE = None, 50, 0.0, 1.0, 1.0

    def __init__(self, type='diff', **kwargs):
        if type == 'diff':
            super().__init__(matching_mat=self.MAT, temporal_init_range_end=self.INIT_END, temporal_range_type='expand', **kwargs)
        else:
            assert type == 'default', type
            super().__init__(matching_mat=self.MAT, temporal_range_type='all', **kwargs)

    def get_step_loss(self):
        cur = float(self.step_loss[self.sim.cur_step_global - 1].item())
        return {'reward': self.SCALE * (self.OFFSET - cur), 'loss': self.STEP_LOSS_SCALE * cur}

    def get_final_loss(self):
        info = super().get_final_loss()
        info['reward'] = float(np.sum((self.OFFSET - self.step_loss.cpu().numpy()) * self.SCALE))
        return info


class IceCreamDynamicLoss(_ScaledShapeLoss):
    """losses/icecreamdynamic_loss.py:14-60: ICECREAM particles against the recorded target, range expanding from 200 steps, reward 0.001 (1700 - loss)"""
    MAT, INIT_END, OFFSET, SCALE, STEP_LOSS_SCALE = ICECREAM, 200, 1700.0, 0.001, 0.001


class IceCreamStaticLoss(_ScaledShapeLoss):
    """losses/icecreamstatic_loss.py:13-55: ICECREAM1 particles, range expanding from 100 steps, reward 0.001 (900 - loss), the step loss unscaled"""
    MAT, INIT_END, OFFSET, SCALE, STEP_LOSS_SCALE = ICECREAM1, 100, 900.0, 0.001, 1.0


class CirculationLoss(Loss):
    """temperature loss of the air-circulation task (losses/circulation_loss.py:14-147): 15 detector cells of the smoke field at height 64;
    the first five should stay hot (|q - 1|), the others reach `target_temp` (|q - 0|).  The per-step value is 15 numbers gathered on the
    device; its seed adds sign(q - target) * weight to the smoke field's q adjoint."""
    DETECTORS = [[25, 85], [35, 85], [15, 85], [25, 75], [25, 95], [25, 42], [35, 42], [15, 42], [25, 32], [25, 52], [107, 65], [115, 65], [99, 65], [107, 45], [107, 85]]

    def __init__(self, type='diff', **kwargs):
        super().__init__(**kwargs)
        self.temporal_range_type = 'all'
        self.target_temp, self.detector_h = 0.0, 64

    def build(self, sim):
        self.temp_weight = self.weights['temp']
        self.temporal_range = [0, self.max_loss_steps]
        self.smoke_field = sim.smoke_field
        assert self.smoke_field is not None, 'CirculationLoss needs a smoke field (losses/loss.py:41-42)'
        n = self.smoke_field.n_grid
        cells = [(x * n + self.detector_h) * n + z for x, z in self.DETECTORS]
        self._cells = torch.tensor(cells, dtype=torch.long, device=sim.device)
        self._target = torch.tensor([1.0] * 5 + [self.target_temp] * 10, dtype=torch.float32, device=sim.device)
        super().build(sim)

    def compute_step_loss(self, s, f):  # circulation_loss.py:85-105: step_loss[s] += w * sum |q[s_local, cell][0] - target|
        q = self.smoke_field._q[self.sim.cur_step_local, 0]
        self.step_loss[s] += self.temp_weight * (q[self._cells] - self._target).abs().sum()

    def compute_step_loss_grad(self, s, f):
        if self._step_grad_on[s]:
            sf = self.smoke_field
            sf._ensure_grad_buffers()
            q = sf._q[self.sim.cur_step_local, 0]
            sf._gq[self.sim.cur_step_local, 0].index_add_(0, self._cells, self.temp_weight * torch.sign(q[self._cells] - self._target))

    def get_final_loss(self):  # circulation_loss.py:118-128
        self.total_loss = float(self.step_loss[self.temporal_range[0]:self.temporal_range[1]].sum().item())
        return {'loss': self.total_loss, 'last_step_loss': float(self.step_loss[self.max_loss_steps - 1].item()), 'temporal_range': self.temporal_range[1]}

    def get_final_loss_grad(self):
        self._step_grad_on[:] = False
        self._step_grad_on[self.temporal_range[0]:self.temporal_range[1]] = True

    def get_step_loss(self):  # circulation_loss.py:136-144
        cur = float(self.step_loss[self.sim.cur_step_global - 1].item())
        return {'reward': 1.0 * (11 - cur), 'loss': 1.0 * cur}
