"""Eulerian smoke solver behind the reference's `SmokeField` interface (fluidlab/fluidengine/simulators/smoke_field.py, abbrev. SF):
`MPMSimulator.step_` / `step_grad` run it once per STEP for the air-circulation task (mpm_simulator.py:744-747, 765-767).

All arithmetic is in libfluidmpm.so (csrc/fsmk_smoke.cu, C ABI include/fluidsmoke.h): free-space mask, RK3 semi-Lagrangian advection +
air-conditioner impulse, divergence, time-blocked Jacobi sweeps, projection, and the hand-written adjoint of every kernel the reference
differentiates with Taichi autodiff.  This class owns the device tensors (torch = allocator), keeps the reference's method names and
state formats (`get_state` / `set_state`: v, v_tmp (n,n,n,3), div, p (n,n,n), q (n,n,n,q_dim)), and has no CPU path: without the CUDA
library or a GPU it fails at build()."""
import ctypes as C
import numpy as np
import torch

from . import _lib
from .macros import DTYPE_NP


class SmokeField:
    def __init__(self, dim, ckpt_dest, res=128, dt=0.03, solver_iters=500, q_dim=3, decay=0.99):  # SF:14-33
        assert dim == 3
        self.dim, self.ckpt_dest = dim, ckpt_dest
        self.n_grid = int(res)
        self.dx = 1 / self.n_grid
        self.res = (self.n_grid,) * dim
        self.dt, self.solver_iters, self.q_dim, self.decay = float(dt), int(solver_iters), int(q_dim), decay
        self.high_T, self.low_T = 1.0, 0.0
        self.lower_y, self.higher_y = 60, 68
        self._h = None

    # ------------------------------------------------------------------------------------------ build, SF:36-93
    def build(self, mpm_sim, agent):
        self.mpm_sim = mpm_sim
        self.max_steps_local = mpm_sim.max_steps_local
        self.agent = mpm_sim.agent if agent is None else agent
        self.device = mpm_sim.device
        self._lib = _lib.load()
        n, S, G, qd, dev, f32 = self.n_grid, self.max_steps_local, self.n_grid ** 3, self.q_dim, self.device, torch.float32
        z = lambda *s, dt=f32: torch.zeros(s, dtype=dt, device=dev)
        self._v, self._vt, self._div, self._p, self._q = z(S + 1, G, 4), z(S + 1, G, 4), z(S + 1, G), z(S + 1, G), z(S + 1, qd, G)
        self._free = z(S + 1, G, dt=torch.uint8)
        self._tmp_a, self._tmp_b, self._acc = z(G), z(G), z(G)
        self._gv = self._gvt = self._gdiv = self._gp = self._gq = None
        cfg = _lib.FsmkConfig()
        cfg.res, cfg.max_steps_local, cfg.q_dim, cfg.solver_iters, cfg.dt = n, S, qd, self.solver_iters, self.dt
        cfg.lower_y, cfg.higher_y, cfg.low_T = int(self.lower_y), int(self.higher_y), float(self.low_T)
        aircon = getattr(self.agent, 'aircon', None)
        assert aircon is not None, 'SmokeField needs an agent with an AirCon effector (agents/agent_circulation.py:14-19)'
        cfg.inject_v = (C.c_float * 3)(*[float(x) for x in aircon.inject_v])
        cfg.device = dev.index if dev.type == 'cuda' and dev.index is not None else 0
        h = C.c_void_p()
        rc = self._lib.fsmk_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise RuntimeError(f'fsmk_create failed (rc={rc}): the smoke solver needs a CUDA device; there is no CPU fallback')
        self._h = h
        self._bind()
        statics = [s for s in (mpm_sim.statics or []) if getattr(s, 'has_dynamics', False)]   # Static.is_collide, meshes/static.py:106-114
        assert len(statics) <= 4, 'at most 4 statics with dynamics'
        arr = (_lib.FmpmSdfMesh * max(len(statics), 1))()
        for i, st in enumerate(statics):
            arr[i] = st.device_struct(_lib, dev)
        self._statics_keepalive = (statics, arr)
        self._ck(self._lib.fsmk_set_statics(self._h, len(statics), arr), 'fsmk_set_statics')
        self.bind_aircon(aircon)
        self.init_fields()
        self.ckpt_ram = dict()

    def bind_aircon(self, aircon):
        """(re)register the air conditioner's per-substep device arrays (they exist once the effector is built)"""
        if self._h is None or getattr(aircon, 'pos', None) is None:
            return
        a = _lib.FsmkAircon()
        a.pos, a.quat, a.s, a.r = aircon.pos.data_ptr(), aircon.quat.data_ptr(), aircon.s.data_ptr(), aircon.r.data_ptr()
        a.gpos, a.gquat, a.gs, a.gr = aircon.gpos.data_ptr(), aircon.gquat.data_ptr(), aircon.gs.data_ptr(), aircon.gr.data_ptr()
        self._ck(self._lib.fsmk_set_aircon(self._h, C.byref(a)), 'fsmk_set_aircon')

    def _bind(self):
        b = _lib.FsmkBuffers()
        p = lambda t: None if t is None else t.data_ptr()
        b.v, b.v_tmp, b.div, b.p, b.q, b.is_free = p(self._v), p(self._vt), p(self._div), p(self._p), p(self._q), p(self._free)
        b.gv, b.gv_tmp, b.gdiv, b.gp, b.gq = p(self._gv), p(self._gvt), p(self._gdiv), p(self._gp), p(self._gq)
        b.tmp_a, b.tmp_b, b.acc = p(self._tmp_a), p(self._tmp_b), p(self._acc)
        self._ck(self._lib.fsmk_bind(self._h, C.byref(b)), 'fsmk_bind')

    def _ensure_grad_buffers(self):
        if self._gv is None:
            self._gv, self._gvt, self._gdiv, self._gp, self._gq = [torch.zeros_like(t) for t in (self._v, self._vt, self._div, self._p, self._q)]
            self._bind()

    def _ck(self, rc, what):
        if rc != 0:
            raise _lib.FmpmError(f'{what}: {self._lib.fsmk_last_error(self._h).decode()}')

    def _stream(self):
        return self.mpm_sim._stream()

    def __del__(self):
        try:
            if self._h is not None:
                self._lib.fsmk_destroy(self._h)
        except Exception:
            pass

    def init_fields(self):  # SF:87-93: the band starts hot (first temperature component only)
        n = self.n_grid
        q0 = self._q[0].view(self.q_dim, n, n, n)
        q0[0, :, max(self.lower_y + 1, 0):max(self.higher_y, 0), :] = self.high_T

    # ------------------------------------------------------------------------------------------ step, SF:95-127
    def step(self, s, f):
        self._ck(self._lib.fsmk_step(self._h, int(s), int(f), self._stream()), 'fsmk_step')

    def step_grad(self, s, f):
        self._ensure_grad_buffers()
        self._ck(self._lib.fsmk_step_grad(self._h, int(s), int(f), self._stream()), 'fsmk_step_grad')

    # ------------------------------------------------------------------------------------------ ring helpers, SF:162-188
    def _state_tensors(self):
        return (self._v, self._vt, self._div, self._p, self._q)

    def _grad_tensors(self):
        self._ensure_grad_buffers()
        return (self._gv, self._gvt, self._gdiv, self._gp, self._gq)

    def copy_frame(self, source, target):
        for t in self._state_tensors():
            t[target].copy_(t[source])

    def copy_grad(self, source, target):
        for t in self._grad_tensors():
            t[target].copy_(t[source])

    def reset_grad(self):
        if self._gv is not None:
            for t in self._grad_tensors():
                t.zero_()

    def reset_grad_till_frame(self, s):
        for t in self._grad_tensors():
            t[:s].zero_()

    # ------------------------------------------------------------------------------------------ state io, SF:362-439
    def _frame_to_ref(self, s, dev_tensors):
        n, qd = self.n_grid, self.q_dim
        v, vt, div, p, q = dev_tensors
        return {'v': v[s, :, :3].reshape(n, n, n, 3), 'v_tmp': vt[s, :, :3].reshape(n, n, n, 3), 'div': div[s].reshape(n, n, n), 'p': p[s].reshape(n, n, n),
                'q': q[s].reshape(qd, n, n, n).permute(1, 2, 3, 0)}

    def get_state(self, s):
        return {k: t.contiguous().cpu().numpy().astype(DTYPE_NP) for k, t in self._frame_to_ref(s, self._state_tensors()).items()}

    def get_grad(self, s):
        """adjoint of frame s in the same format (tests / diagnostics; the reference has no accessor for it)"""
        return {k: t.contiguous().cpu().numpy().astype(DTYPE_NP) for k, t in self._frame_to_ref(s, self._grad_tensors()).items()}

    def _write_frame(self, s, state, tensors):
        n, qd, dev = self.n_grid, self.q_dim, self.device
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32) if not torch.is_tensor(a) else a).to(dev, torch.float32)
        v, vt, div, p, q = tensors
        v[s, :, :3] = t(state['v']).reshape(-1, 3); vt[s, :, :3] = t(state['v_tmp']).reshape(-1, 3)
        div[s] = t(state['div']).reshape(-1); p[s] = t(state['p']).reshape(-1)
        q[s] = t(state['q']).reshape(n, n, n, qd).permute(3, 0, 1, 2).reshape(qd, -1)

    def set_state(self, s, state):
        self._write_frame(s, state, self._state_tensors())

    def set_grad(self, s, state):
        self._write_frame(s, state, self._grad_tensors())

    def is_free(self, s):
        n = self.n_grid
        return self._free[s].reshape(n, n, n).cpu().numpy().astype(np.int32)

    def get_ckpt(self, ckpt_name=None):  # SF:384-416 — frame 0 on the checkpoint device
        d = {'gpu': self.device, 'cpu': torch.device('cpu'), 'disk': torch.device('cpu')}[self.ckpt_dest]
        ckpt = {k: t[0].to(d, copy=True) for k, t in zip(('v', 'v_tmp', 'div', 'p', 'q'), self._state_tensors())}
        if ckpt_name is not None:
            self.ckpt_ram[ckpt_name] = ckpt
        return ckpt

    def set_ckpt(self, ckpt=None, ckpt_name=None):  # SF:418-425
        if ckpt is None:
            ckpt = self.ckpt_ram[ckpt_name]
        for k, t in zip(('v', 'v_tmp', 'div', 'p', 'q'), self._state_tensors()):
            t[0].copy_(ckpt[k])
