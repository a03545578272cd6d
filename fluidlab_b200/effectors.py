"""Effectors: per-substep pose state + action buffers on the device, driven through libfluidmpm.so.

Mirrors fluidlab/fluidengine/effectors/effector.py (`Effector`: fields :34-51, move_kernel :157-161, set_action
:262-268, set_velocity :252-260, apply_action_p :223-231, get_action_grad :276-283, ckpt :83-139) and
effectors/injector.py (`Injector` :12-105, `BallInjector` :215-256).  The pose chain of one step (set_action +
n_substeps move_kernel calls) is one tiny kernel launch (fmpm_effector_step); its adjoint likewise.
`AirCon` (effectors/aircon.py) adds the strength / radius channels the smoke solver reads.
"""
import ctypes as C
import numpy as np
import torch
from scipy.spatial.transform import Rotation

from . import _lib
from .boundaries import create_boundary, _tup
from .macros import DTYPE_NP


def _xyzw_to_wxyz(q):
    return np.array([q[3], q[0], q[1], q[2]])


class _HostField:
    def __init__(self, val):
        self._val = val

    def to_numpy(self):
        return self._val


class Effector:
    state_dim = 7

    def __init__(self, max_substeps_local, max_substeps_global, max_action_steps_global, ckpt_dest, dim=3, action_dim=3,
                 action_scale_p=(1.0, 1.0, 1.0), action_scale_v=(1.0, 1.0, 1.0), init_pos=(0.5, 0.5, 0.5), init_euler=(0.0, 0.0, 0.0)):
        self.dim = dim
        self.max_substeps_local = max_substeps_local
        self.max_substeps_global = max_substeps_global
        self.max_action_steps_global = max_action_steps_global
        self.ckpt_dest = ckpt_dest
        self.action_dim = action_dim
        assert action_dim in (0, 3, 6, 8), 'action layouts of the reference: none, xyz, xyz + rotation, AirCon (+ strength, radius)'
        self.init_pos = np.array(_tup(init_pos))
        self.init_rot = _xyzw_to_wxyz(Rotation.from_euler('zyx', _tup(init_euler)[::-1], degrees=True).as_quat())
        self.action_scale_v = np.array(list(_tup(action_scale_v)) + [1.0] * 8, dtype=DTYPE_NP)[:8]
        self.action_scale_p = np.array(list(_tup(action_scale_p)) + [1.0] * 8, dtype=DTYPE_NP)[:8]
        self.boundary = None
        self.mesh = None
        self.sim = None

    def setup_boundary(self, **kwargs):
        self.boundary = create_boundary(**kwargs)

    def setup_mesh(self, **kwargs):
        self.mesh_cfg = kwargs  # visual only for injectors (has_dynamics=False, injector.py:32)

    @property
    def init_state(self):
        return np.append(self.init_pos, self.init_rot)

    # ---- device state
    def build(self, sim):
        self.sim = sim
        dev, T, f32 = sim.device, self.max_substeps_local, torch.float32
        z = lambda *s: torch.zeros(s, dtype=f32, device=dev)
        self.pos, self.quat, self.v, self.w = z(T + 1, 3), z(T + 1, 4), z(T + 1, 3), z(T + 1, 3)
        self.gpos, self.gquat, self.gv, self.gw = z(T + 1, 3), z(T + 1, 4), z(T + 1, 3), z(T + 1, 3)
        ad = max(self.action_dim, 1)
        self.action_buffer, self.action_buffer_grad = z(self.max_action_steps_global + 1, ad), z(self.max_action_steps_global + 1, ad)
        self.action_buffer_p, self.action_buffer_p_grad = z(ad), z(ad)
        self._act_dev = z(64, ad)
        self._act_host = torch.zeros((64, ad), dtype=f32).pin_memory()
        self._act_events = [None] * 64
        self._act_slot = 0
        if self.boundary is None:
            self.boundary = create_boundary()
        e = _lib.FmpmEffector()
        for name, t in (('pos', self.pos), ('quat', self.quat), ('v', self.v), ('w', self.w), ('gpos', self.gpos), ('gquat', self.gquat),
                        ('gv', self.gv), ('gw', self.gw), ('act', self.action_buffer), ('gact', self.action_buffer_grad),
                        ('act_p', self.action_buffer_p), ('gact_p', self.action_buffer_p_grad)):
            setattr(e, name, t.data_ptr())
        e.action_dim = self.action_dim
        e.scale_v = (C.c_float * 6)(*[float(x) for x in self.action_scale_v[:6]]); e.scale_p = (C.c_float * 6)(*[float(x) for x in self.action_scale_p[:6]])
        b = self.boundary
        e.boundary_type = b.type_id
        e.b_lower = (C.c_float * 3)(*[float(x) for x in b.lower]); e.b_upper = (C.c_float * 3)(*[float(x) for x in b.upper])
        e.cyl_center = (C.c_float * 2)(*[float(x) for x in b.xz_center]); e.cyl_radius = float(b.xz_radius)
        self._c = e
        self.ckpt_ram = dict()
        self.set_state(0, self.init_state)

    def reset_grad(self):  # effector.py:75-81
        for t in (self.gpos, self.gquat, self.gv, self.gw, self.action_buffer_grad, self.action_buffer_p_grad):
            t.zero_()

    def _stage_action(self, action):
        i = self._act_slot
        self._act_slot = (i + 1) % 64
        if self._act_events[i] is not None:
            self._act_events[i].synchronize()
        a = np.zeros(self._act_host.shape[1], dtype=np.float32); a[:len(action)] = action
        self._act_host[i].copy_(torch.from_numpy(a))
        self._act_dev[i].copy_(self._act_host[i], non_blocking=True)
        ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream(self.sim.device)); self._act_events[i] = ev
        return self._act_dev[i]

    # ---- pose chain
    def set_action(self, s, s_global, n_substeps, action):  # effector.py:262-268 (+ the step's move_kernel calls, :146-161)
        assert s_global <= self.max_action_steps_global
        assert s * n_substeps <= self.max_substeps_local
        sim = self.sim
        a = self._stage_action(np.asarray(action, dtype=np.float32).reshape(-1))
        sim._ck(sim._lib.fmpm_effector_step(sim._h, C.byref(self._c), s, s_global, a.data_ptr(), sim._stream()), 'fmpm_effector_step')
        self._latest_f = (s + 1) * n_substeps - 1     # update_latest_pos(f) of the step's last move (effector.py:146-152)

    @property
    def latest_pos(self):
        """effector.py:58,150-152: position at the last moved substep, as an object with .to_numpy() -> float32 [1, 3] (zeros before any move)"""
        f = getattr(self, '_latest_f', None)
        val = np.zeros((1, 3), np.float32) if f is None else self.pos[f].detach().cpu().numpy().astype(np.float32)[None]
        return _HostField(val)

    def set_action_grad(self, s, s_global, n_substeps, action):  # effector.py:270-274 (+ move_kernel.grad)
        assert s_global <= self.max_action_steps_global
        assert s * n_substeps <= self.max_substeps_local
        sim = self.sim
        sim._ck(sim._lib.fmpm_effector_step_grad(sim._h, C.byref(self._c), s, s_global, sim._stream()), 'fmpm_effector_step_grad')

    def apply_action_p(self, action_p):  # effector.py:228-231
        ap = np.zeros(self.action_buffer_p.shape[0], dtype=np.float32); ap[:len(action_p)] = np.asarray(action_p, dtype=np.float32)
        self.action_buffer_p.copy_(torch.from_numpy(ap))
        sim = self.sim
        sim._ck(sim._lib.fmpm_effector_apply_action_p(sim._h, C.byref(self._c), sim._stream()), 'fmpm_effector_apply_action_p')

    def apply_action_p_grad(self, action_p):  # effector.py:233-234
        sim = self.sim
        sim._ck(sim._lib.fmpm_effector_apply_action_p_grad(sim._h, C.byref(self._c), sim._stream()), 'fmpm_effector_apply_action_p_grad')

    def get_action_grad(self, s, n):  # effector.py:276-283
        if self.action_dim > 0:
            grad = np.zeros((n + 1, self.action_dim), dtype=DTYPE_NP)
            grad[:n] = self.action_buffer_grad[s:s + n, :self.action_dim].cpu().numpy()
            grad[n] = self.action_buffer_p_grad[:self.action_dim].cpu().numpy()
            return grad
        return None

    def get_action_grad_device(self, s, n):
        """get_action_grad without the device->host copy: float32 [n + 1, action_dim] on the device (or None)"""
        if self.action_dim > 0:
            return torch.cat([self.action_buffer_grad[s:s + n, :self.action_dim], self.action_buffer_p_grad[None, :self.action_dim]], dim=0)
        return None

    def move(self, f):
        pass  # folded into set_action (one kernel per step)

    def move_grad(self, f):
        pass  # folded into set_action_grad

    # ---- state / ring helpers (effector.py:164-213)
    def get_state(self, f):
        return torch.cat([self.pos[f], self.quat[f]]).cpu().numpy().astype(DTYPE_NP)

    def set_state(self, f, state):
        ss = self.get_state(f)
        ss[:len(state)] = state
        t = torch.from_numpy(np.asarray(ss[:7], dtype=np.float32)).to(self.pos.device)
        self.pos[f] = t[:3]; self.quat[f] = t[3:7]

    def copy_frame(self, source, target):
        for t in (self.pos, self.quat, self.v, self.w):
            t[target] = t[source]

    def copy_grad(self, source, target):
        for t in (self.gpos, self.gquat, self.gv, self.gw):
            t[target] = t[source]

    def reset_grad_till_frame(self, f):
        for t in (self.gpos, self.gquat, self.gv, self.gw):
            t[:f].zero_()

    def get_ckpt(self):
        return {'pos': self.pos[0].clone(), 'quat': self.quat[0].clone(), 'v': self.v[0].clone(), 'w': self.w[0].clone()}

    def set_ckpt(self, ckpt):
        self.pos[0] = ckpt['pos']; self.quat[0] = ckpt['quat']; self.v[0] = ckpt['v']; self.w[0] = ckpt['w']


class Injector(Effector):
    state_dim = 7
    kind = 1

    def __init__(self, radius=1.0, flux=1, inject_v=(0.0, 0.0, 0.0), inject_p=(0.0, 0.0, 0.0), randomize_inject_v=False,
                 locally_random=False, **kwargs):
        super().__init__(**kwargs)
        self.randomize_inject_v = bool(randomize_inject_v)   # injector.py:96-97 (Injector.act only; BallInjector.act ignores it, :240-256)
        self.radius = radius
        self.n_particles = flux
        self.locally_random = locally_random
        self.inject_v = np.array(_tup(inject_v), dtype=DTYPE_NP)
        self.inject_p = np.array(_tup(inject_p), dtype=DTYPE_NP)
        self.has_dynamics = False
        self.act_id = [0] * (self.max_substeps_local + 1)
        self.init_random_vector()

    def init_random_vector(self):  # injector.py:54-60 — drawn from the global NumPy RNG, like the reference
        random_length = self.max_substeps_local if self.locally_random else self.max_substeps_global
        self.random_vector_np = np.random.uniform(size=(random_length, self.n_particles, self.dim)).astype(DTYPE_NP)

    def set_act_range(self, used):  # injector.py:62-68
        act_range = np.where(np.asarray(used) == 0)[0].astype(np.int32)
        self.act_range_np = act_range
        self.act_id[0] = 0  # index into act_range (the reference stores act_range[0] there and then indexes with it: same when it is 0)

    def build(self, sim):
        super().build(sim)
        self._random_vector = torch.from_numpy(self.random_vector_np.copy()).to(sim.device)

    def finalize(self):
        sim = self.sim
        self._act_range = torch.from_numpy(self.act_range_np.copy()).to(sim.device)
        inj = _lib.FmpmInjector()
        inj.kind, inj.flux, inj.radius = self.kind, int(self.n_particles), float(self.radius)
        inj.inject_v = (C.c_float * 3)(*[float(x) for x in self.inject_v]); inj.inject_p = (C.c_float * 3)(*[float(x) for x in self.inject_p])
        inj.random_vector = self._random_vector.data_ptr(); inj.act_range = self._act_range.data_ptr(); inj.n_act_range = len(self.act_range_np)
        inj.randomize_inject_v = int(self.randomize_inject_v and self.kind == 1)
        self._inj = inj

    def act(self, f, f_global):  # injector.py:80-105
        sim = self.sim
        assert self.act_id[f] + self.n_particles <= len(self.act_range_np), 'too many particles added'
        row = f if self.locally_random else f_global
        sim._ck(sim._lib.fmpm_inject(sim._h, f, C.byref(self._inj), C.byref(self._c), self.act_id[f], row, sim._frame_ord[f + 1].inv_ptr(),
                                     sim._stream()), 'fmpm_inject')
        self.act_id[f + 1] = self.act_id[f] + self.n_particles

    def act_grad(self, f, f_global, gin):
        sim = self.sim
        sim._ck(sim._lib.fmpm_inject_grad(sim._h, f, gin, C.byref(self._inj), C.byref(self._c), self.act_id[f], sim._frame_ord[f].inv_ptr(),
                                          sim._stream()), 'fmpm_inject_grad')

    def get_state(self, f):  # injector.py:197-205
        out = np.zeros(8, dtype=DTYPE_NP)
        out[:7] = super().get_state(f)
        out[7] = self.act_id[f]
        return out

    def set_state(self, f, state):
        ss = self.get_state(f)
        ss[:len(state)] = state
        t = torch.from_numpy(np.asarray(ss[:7], dtype=np.float32)).to(self.pos.device)
        self.pos[f] = t[:3]; self.quat[f] = t[3:7]
        self.act_id[f] = int(ss[7])

    def copy_frame(self, source, target):
        super().copy_frame(source, target)
        self.act_id[target] = self.act_id[source]

    def get_ckpt(self):
        c = super().get_ckpt(); c['act_id'] = self.act_id[0]; return c

    def set_ckpt(self, ckpt):
        super().set_ckpt(ckpt); self.act_id[0] = ckpt['act_id']


class BallInjector(Injector):
    kind = 2

    def init_random_vector(self):  # injector.py:224-238
        random_length = self.max_substeps_local if self.locally_random else self.max_substeps_global
        chunks, n_generated = [], 0
        while True:
            rand_pos = np.random.uniform(high=self.radius, low=-self.radius, size=(self.n_particles * random_length, 3))
            rand_pos = rand_pos[np.linalg.norm(rand_pos, axis=1) <= self.radius]
            n_generated += rand_pos.shape[0]
            chunks.append(rand_pos)
            if n_generated >= self.n_particles * random_length:
                break
        self.random_vector_np = np.concatenate(chunks)[:self.n_particles * random_length].reshape([random_length, self.n_particles, 3]).astype(DTYPE_NP)


class Rigid(Effector):
    """Rigid end-effector with an SDF mesh (effectors/rigid.py:11-38): `collide` = mesh.collide (meshes/dynamic.py:93-121),
    evaluated inside the CUDA kernels (csrc/fmpm_sdf.cuh) at particle and/or grid level."""

    def setup_mesh(self, **kwargs):
        from .meshes import Dynamic
        self.mesh = Dynamic(container=self, has_dynamics=True, **kwargs)


class AirCon(Effector):
    """The air conditioner of the circulation task (effectors/aircon.py:12-26): a 6-DOF pose chain plus a blowing strength `s` and a
    radius `r` per substep, set from action components 6 and 7 (aircon.py:225-241).  The smoke solver reads pos/quat/s/r[f] and
    accumulates their adjoints (csrc/fsmk_smoke.cu); the pose part of the chain is the ordinary effector kernel."""
    state_dim = 9

    def __init__(self, inject_v=(-0.3, 0.0, 1.0), **kwargs):
        super().__init__(**kwargs)
        self.inject_v = np.array(_tup(inject_v), dtype=DTYPE_NP)
        self.has_dynamics = False

    def build(self, sim):
        T, dev = self.max_substeps_local, sim.device
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        self.s, self.r, self.gs, self.gr = z(T + 1), z(T + 1), z(T + 1), z(T + 1)
        super().build(sim)

    def reset_grad(self):  # aircon.py:52-60
        super().reset_grad()
        self.gs.zero_(); self.gr.zero_()

    def set_action(self, s, s_global, n_substeps, action):  # aircon.py:225-241
        super().set_action(s, s_global, n_substeps, action)
        if self.action_dim > 6:
            j0, j1 = s * n_substeps, (s + 1) * n_substeps
            self.s[j0:j1] = self.action_buffer[s_global, 6] * float(self.action_scale_v[6])
            self.r[j0:j1] = self.action_buffer[s_global, 7] * float(self.action_scale_v[7])

    def set_action_grad(self, s, s_global, n_substeps, action):
        if self.action_dim > 6:
            j0, j1 = s * n_substeps, (s + 1) * n_substeps
            self.action_buffer_grad[s_global, 6] += self.gs[j0:j1].sum() * float(self.action_scale_v[6])
            self.action_buffer_grad[s_global, 7] += self.gr[j0:j1].sum() * float(self.action_scale_v[7])
        super().set_action_grad(s, s_global, n_substeps, action)

    def get_state(self, f):  # aircon.py:188-211
        out = np.zeros(9, dtype=DTYPE_NP)
        out[:7] = super().get_state(f)
        out[7], out[8] = float(self.s[f]), float(self.r[f])
        return out

    def set_state(self, f, state):
        ss = self.get_state(f)
        ss[:len(state)] = state
        t = torch.from_numpy(np.asarray(ss, dtype=np.float32)).to(self.pos.device)
        self.pos[f] = t[:3]; self.quat[f] = t[3:7]; self.s[f] = t[7]; self.r[f] = t[8]

    def copy_frame(self, source, target):  # aircon.py:154-161
        super().copy_frame(source, target)
        self.s[target] = self.s[source]; self.r[target] = self.r[source]

    def copy_grad(self, source, target):
        super().copy_grad(source, target)
        self.gs[target] = self.gs[source]; self.gr[target] = self.gr[source]

    def reset_grad_till_frame(self, f):
        super().reset_grad_till_frame(f)
        self.gs[:f].zero_(); self.gr[:f].zero_()

    def get_ckpt(self):
        c = super().get_ckpt(); c['s'] = self.s[0].clone(); c['r'] = self.r[0].clone(); return c

    def set_ckpt(self, ckpt):
        super().set_ckpt(ckpt); self.s[0] = ckpt['s']; self.r[0] = ckpt['r']
