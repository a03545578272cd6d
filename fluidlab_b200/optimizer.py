"""Trajectory optimisation above the simulator: the reference's Solver / TrainablePolicy / Adam (fluidlab/optimizer/solver.py:14-59,
optimizer/policies.py:131-169, optimizer/optim.py:3-41) with the optimiser state resident on the GPU.

The composite action table ((horizon + 1) x action_dim float64: the action_v rows, then action_p), Adam's two moment tables and the gradient
(`agent.get_grad_device`, float32) stay on the device; one launch of `fmpm_adam_step` (csrc/fmpm_io.cu: k_adam_step) replaces the NumPy update
and its result is bit-identical to the reference's (same dtypes, same rounding per operation).  The host keeps a float64 mirror of the table
(one small device->host copy per iteration) because `TaichiEnv.step(action)` takes host actions, exactly like the reference.

There is no CPU implementation here: an unbound policy (no simulator / library handle) raises."""
import ctypes as C

import numpy as np
import torch

from . import _lib


class Optimizer:  # optim.py:3-20
    def __init__(self, parameters_shape, cfg):
        self.cfg = cfg
        self.lr = self.cfg.lr
        self.init_lr = self.cfg.lr
        self.parameters_shape = tuple(parameters_shape)
        self._sim = None
        self.initialize()

    def initialize(self):
        raise NotImplementedError

    def bind(self, sim):
        """sim: the MPMSimulator whose library handle, device and stream the update kernel uses"""
        self._sim = sim
        self._to_device()

    def step(self, parameters, grads):
        return self._step(parameters, grads)


class Adam(Optimizer):  # optim.py:22-41
    def initialize(self):
        self.momentum_buffer = None
        self.v_buffer = None
        self.iter = 0

    def _to_device(self):
        dev = self._sim.device
        self.momentum_buffer = torch.zeros(self.parameters_shape, dtype=torch.float64, device=dev)
        self.v_buffer = torch.zeros(self.parameters_shape, dtype=torch.float64, device=dev)

    def step_device(self, table, grads, trainable=None, fix_dim_mask=0, clip=(-np.inf, np.inf)):
        """in-place update of the device table (float64 [rows, cols]) from device grads (float32 [rows, cols]); the last row is not clipped"""
        if self._sim is None:
            raise RuntimeError('Adam: not bound to a simulator (Optimizer.bind) - the update runs on the GPU, there is no CPU fallback')
        assert tuple(table.shape) == tuple(grads.shape) == self.parameters_shape and table.dtype == torch.float64 and grads.dtype == torch.float32
        assert table.is_contiguous() and grads.is_contiguous()
        b1, b2 = float(self.cfg.beta_1), float(self.cfg.beta_2)
        c = _lib.FmpmAdamCfg(lr=float(self.lr), beta_1=b1, beta_2=b2, epsilon=float(self.cfg.epsilon), bias_1=1 - b1 ** (self.iter + 1),
                             bias_2=1 - b2 ** (self.iter + 1), clip_lo=float(clip[0]), clip_hi=float(clip[1]), rows=self.parameters_shape[0],
                             cols=self.parameters_shape[1], fix_dim_mask=int(fix_dim_mask), reserved=0)
        sim = self._sim
        sim._ck(sim._lib.fmpm_adam_step(sim._h, C.byref(c), table.data_ptr(), self.momentum_buffer.data_ptr(), self.v_buffer.data_ptr(), grads.data_ptr(),
                                        None if trainable is None else trainable.data_ptr(), sim._stream()), 'fmpm_adam_step')
        self.iter += 1

    def _step(self, parameters, grads):
        """the reference's call shape (host arrays in, new host array out); the arithmetic still runs on the device"""
        if self._sim is None:
            raise RuntimeError('Adam: not bound to a simulator (Optimizer.bind) - the update runs on the GPU, there is no CPU fallback')
        dev = self._sim.device
        table = torch.from_numpy(np.ascontiguousarray(parameters, dtype=np.float64).reshape(self.parameters_shape)).to(dev)
        g = torch.from_numpy(np.ascontiguousarray(grads, dtype=np.float32)).to(dev)
        self.step_device(table, g, clip=(-np.inf, np.inf))
        return table.cpu().numpy()


class ActionsPolicy:  # policies.py:10-19
    def __init__(self, comp_actions):
        self.actions_v = comp_actions[:-1]
        self.actions_p = comp_actions[-1]

    def get_actions_p(self):
        return self.actions_p

    def get_action_v(self, i, **kwargs):
        return self.actions_v[i]


class TrainablePolicy:  # policies.py:131-164
    def __init__(self, optim_cfg, init_range, action_dim, horizon, action_range, fix_dim=None, sim=None):
        self.horizon = horizon
        self.action_dim = action_dim
        self.actions_v = np.random.uniform(init_range.v[0], init_range.v[1], size=(horizon, action_dim))
        self.actions_p = np.random.uniform(init_range.p[0], init_range.p[1], size=(action_dim))
        self.action_range = action_range
        self.comp_actions_shape = (horizon + 1, action_dim)
        self.trainable = np.full(self.comp_actions_shape[0], True)
        self.fix_dim = fix_dim
        self.freeze_till = 0
        self.optim = {'Adam': Adam}[optim_cfg.type](self.comp_actions_shape, optim_cfg)
        self._sim = None
        self._table = None
        if sim is not None:
            self.bind(sim)

    def bind(self, sim):
        self._sim = sim
        self.optim.bind(sim)
        self._table = torch.from_numpy(self.comp_actions.astype(np.float64)).to(sim.device).contiguous()
        return self

    @property
    def comp_actions(self):
        return np.vstack([self.actions_v, self.actions_p[None, :]])

    def get_actions_p(self):
        return self.actions_p

    def get_action_v(self, i, **kwargs):
        return self.actions_v[i]

    def optimize(self, grads, loss_info=None):
        """grads: device float32 tensor (agent.get_grad_device) or a host array (agent.get_grad); either way the update is one kernel on the
        device table and the host mirror is refreshed with one small copy."""
        if self._sim is None:
            raise RuntimeError('TrainablePolicy: not bound to a simulator (bind(sim)) - optimize runs on the GPU, there is no CPU fallback')
        dev = self._sim.device
        g = grads if torch.is_tensor(grads) else torch.from_numpy(np.ascontiguousarray(grads, dtype=np.float32))
        g = g.to(device=dev, dtype=torch.float32).contiguous()
        assert tuple(g.shape) == self.comp_actions_shape
        # the host mirror is authoritative between iterations (callers may edit actions_v / actions_p, as reference scripts do)
        self._table.copy_(torch.from_numpy(self.comp_actions.astype(np.float64)))
        trainable = torch.from_numpy(np.ascontiguousarray(self.trainable, dtype=np.uint8)).to(dev)
        mask = 0
        if self.fix_dim is not None:
            ncol = int(self.comp_actions_shape[-1])
            assert ncol <= 32, 'fix_dim_mask covers at most 32 action columns'
            for d in np.atleast_1d(self.fix_dim):   # the reference's `grads[:, fix_dim] = 0` takes negative indices too
                assert -ncol <= int(d) < ncol, f'fix_dim {d} out of range for {ncol} action columns'
                mask |= 1 << (int(d) % ncol)
        self.optim.step_device(self._table, g, trainable=trainable, fix_dim_mask=mask, clip=self.action_range)
        new = self._table.cpu().numpy()
        self.actions_p = new[-1]
        self.actions_v = new[:-1]


class _TaskPolicy(TrainablePolicy):
    """The task policies of policies.py that differ from TrainablePolicy only by a rule table: which rows train, a gradient clip, and
    learning-rate / freeze schedules keyed on loss_info['temporal_range'] (the loss's current horizon)."""
    TRAINABLE = None        # None = every row; else slice bounds (start, stop) of the rows that train
    GRAD_CLIP = None        # clip of the gradient before the update
    LR_SCHEDULE = ()        # ((temporal_range >, lr factor), ...) first match wins
    FREEZE_SCHEDULE = ()    # ((temporal_range >, freeze rows [:n]), ...) first match wins

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.TRAINABLE is not None:
            self.trainable = np.full(self.comp_actions_shape[0], False)
            self.trainable[slice(*self.TRAINABLE)] = True

    def optimize(self, grads, loss_info=None):
        if self.GRAD_CLIP is not None:
            grads = grads.clamp(*self.GRAD_CLIP) if torch.is_tensor(grads) else np.clip(grads, *self.GRAD_CLIP)
        super().optimize(grads, loss_info)
        tr = (loss_info or {}).get('temporal_range', 0)
        for above, factor in self.LR_SCHEDULE:
            if tr > above:
                self.optim.lr = self.optim.init_lr * factor
                break
        for above, n in self.FREEZE_SCHEDULE:
            if tr > above:
                self.trainable[:n] = False
                break


class _PhasedPolicy(TrainablePolicy):
    """Gathering / GatheringO / Mixing (policies.py:218-339): the horizon is cut into periods; inside a period only the first phase trains, the
    other phases are scripted in get_action_v(update=True) - lift by a fixed velocity, travel back towards a home position at the speed that
    arrives at the phase's end (reads the effector's latest position), lower again."""
    PHASES = ()          # phase end offsets inside one period; the last one is the period
    LIFT = 0.008
    KEEP_HEIGHT = True   # the travel-back phase keeps y
    HOME = None          # None = actions_p

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        n, period = self.comp_actions_shape[0], self.PHASES[-1]
        self.stage_step = list(self.PHASES)
        self.trainable = np.full(n, False)
        self.status = np.full(n, 0)
        for i in range(self.horizon):
            self.status[i] = int(np.searchsorted(np.asarray(self.PHASES), i % period, side='right'))
            self.trainable[i] = self.status[i] == 0

    def _script(self, i, agent):
        raise NotImplementedError

    def get_action_v(self, i, agent=None, update=False):
        if update and self.status[i] != 0:
            self.actions_v[i] = self._script(i, agent)
        return self.actions_v[i]

    def _travel_back(self, i, agent, end):
        home = self.actions_p if self.HOME is None else np.array(self.HOME)
        a = (home - agent.rigid.latest_pos.to_numpy()[0]) / (end - (i % self.PHASES[-1]))
        if self.KEEP_HEIGHT:
            a[1] = 0
        return a


class GatheringOPolicy(_PhasedPolicy):  # policies.py:262-303
    PHASES = (50, 65, 105, 120)

    def _script(self, i, agent):
        st = self.status[i]
        if st == 2:
            return self._travel_back(i, agent, self.PHASES[2])
        return np.array([0, self.LIFT if st == 1 else -self.LIFT, 0])


class GatheringPolicy(GatheringOPolicy):  # policies.py:218-259
    def optimize(self, grads, loss_info=None):
        tr = (loss_info or {}).get('temporal_range', 0)
        for step in (720, 600, 480, 360, 240, 120):
            if tr > step:
                self.freeze_till = tr - 120
                self.trainable[:self.freeze_till] = False
                break
        super().optimize(grads, loss_info)


class MixingPolicy(_PhasedPolicy):  # policies.py:306-339
    PHASES = (50, 80)
    KEEP_HEIGHT = False
    HOME = (0.5, 0.73, 0.5)

    def _script(self, i, agent):
        return self._travel_back(i, agent, self.PHASES[1])

    def optimize(self, grads, loss_info=None):
        super().optimize(grads, loss_info)
        tr = (loss_info or {}).get('temporal_range', 0)
        for step in range(1920, 79, -80):
            if tr > step:
                self.freeze_till = tr - 160
                self.trainable[:self.freeze_till] = False
                break


class LatteArtPolicy(TrainablePolicy):  # policies.py:167-169
    pass


class LatteArtStirPolicy(_TaskPolicy):  # policies.py:172-192
    LR_SCHEDULE = ((250, 0.2), (150, 0.5))
    FREEZE_SCHEDULE = tuple((step, step - 100) for step in (400, 350, 300, 250, 200, 150, 100))


class IceCreamDynamicPolicy(_TaskPolicy):  # policies.py:195-200
    TRAINABLE = (169, -1)


class IceCreamStaticPolicy(_TaskPolicy):  # policies.py:203-215
    TRAINABLE = (None, -1)
    GRAD_CLIP = (-1e5, 1e5)
    LR_SCHEDULE = ((450, 0.1),)


class CirculationPolicy(TrainablePolicy):  # policies.py:341-344
    pass


class PouringPolicy(TrainablePolicy):  # policies.py:357-359
    pass


class TransportingPolicy(_TaskPolicy):  # policies.py:362-366
    TRAINABLE = (None, -1)


def trainable_policy(taichi_env, policy_cls, optim_cfg, init_range, horizon_action, action_range, fix_dim=None):
    """what every env's trainable_policy(optim_cfg, init_range) does (e.g. envs/latteart_env.py:97-101): the task's policy sized by the agent's
    action_dim, already bound to the simulator that runs its update kernel"""
    return policy_cls(optim_cfg, init_range, taichi_env.agent.action_dim, horizon_action, action_range, fix_dim=fix_dim, sim=taichi_env.simulator)


def forward_backward(taichi_env, sim_state, policy, horizon_action, device_grad=True):
    """solver.py:23-59: one rollout with gradients -> (loss_info, dLoss/d(comp_actions)).  device_grad: return the gradient as a device tensor
    (agent.get_grad_device) so that policy.optimize consumes it without a round trip."""
    taichi_env.set_state(sim_state, grad_enabled=True)
    taichi_env.apply_agent_action_p(policy.get_actions_p())
    cur_horizon = taichi_env.loss.temporal_range[1]
    for i in range(cur_horizon):
        taichi_env.step(policy.get_action_v(i, agent=taichi_env.agent, update=True) if i < horizon_action else None)
    loss_info = taichi_env.get_final_loss()
    taichi_env.reset_grad()
    taichi_env.get_final_loss_grad()
    for i in range(cur_horizon - 1, policy.freeze_till - 1, -1):
        taichi_env.step_grad(policy.get_action_v(i) if i < horizon_action else None)
    taichi_env.apply_agent_action_p_grad(policy.get_actions_p())
    agent = taichi_env.agent
    return loss_info, (agent.get_grad_device(horizon_action) if device_grad else agent.get_grad(horizon_action))


class Solver:  # solver.py:10-67 (rendering and file logging stay with the caller)
    def __init__(self, env, logger=None, cfg=None):
        self.cfg = cfg
        self.env = env
        self.logger = logger

    def solve(self, policy=None, callback=None):
        """env: anything with .taichi_env, .horizon_action and .trainable_policy(optim_cfg, init_range) (the reference's FluidEnv interface);
        returns the optimised policy.  callback(iteration, loss_info) replaces the reference's logger calls when no logger is given."""
        taichi_env = self.env.taichi_env
        if policy is None:
            policy = self.env.trainable_policy(self.cfg.optim, self.cfg.init_range)
        if getattr(policy, '_sim', None) is None:
            policy.bind(taichi_env.simulator)
        taichi_env_state = taichi_env.get_state()
        for iteration in range(self.cfg.n_iters):
            if self.logger is not None and hasattr(self.logger, 'save_policy'):
                self.logger.save_policy(policy, iteration)
            loss_info, grad = forward_backward(taichi_env, taichi_env_state['state'], policy, self.env.horizon_action)
            loss_info['iteration'] = iteration
            policy.optimize(grad, loss_info)
            loss_info['lr'] = policy.optim.lr
            if self.logger is not None:
                self.logger.log(iteration, loss_info)
            if callback is not None:
                callback(iteration, loss_info)
        return policy
