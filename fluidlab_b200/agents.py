"""Agents: groups of effectors.  Mirrors fluidlab/fluidengine/agents/agent.py (`Agent` :9-152),
agents/agent_injector.py (`AgentInjector` :8-39), agents/agent_rigid.py, agents/agent_icecreamdynamic.py and the collector agents
agents/agent_pouring.py / agents/agent_jetbot.py.  `collide` of injector agents is the identity (agent_injector.py:34-36), so no
collision kernel is involved for them."""
import ctypes as C
import numpy as np
import torch
from . import _lib
from .boundaries import create_boundary
from .macros import WATER
from .effectors import Effector, Injector, BallInjector, Rigid, AirCon  # noqa: F401 (names are eval'ed from yaml, agent.py:32)


class Agent:
    def __init__(self, max_substeps_local, max_substeps_global, max_action_steps_global, ckpt_dest, collide_type='particle'):
        self.max_substeps_local = max_substeps_local
        self.max_substeps_global = max_substeps_global
        self.max_action_steps_global = max_action_steps_global
        self.ckpt_dest = ckpt_dest
        self.collide_type = collide_type
        assert self.collide_type in ['particle', 'grid', 'both']
        self.effectors = []
        self.action_dims = [0]

    def add_effector(self, type, params, mesh_cfg, boundary_cfg):
        cls = eval(type) if isinstance(type, str) else type
        effector = cls(max_substeps_local=self.max_substeps_local, max_substeps_global=self.max_substeps_global,
                       max_action_steps_global=self.max_action_steps_global, ckpt_dest=self.ckpt_dest, **dict(params))
        if mesh_cfg is not None:
            effector.setup_mesh(**dict(mesh_cfg))
        effector.setup_boundary(**dict(boundary_cfg))
        self.effectors.append(effector)
        self.action_dims.append(self.action_dims[-1] + effector.action_dim)

    def build(self, sim):
        self.n_effectors = len(self.effectors)
        self.sim = sim
        for effector in self.effectors:
            effector.build(sim)

    def reset_grad(self):
        for e in self.effectors:
            e.reset_grad()

    def act(self, f, f_global):
        return

    def act_grad(self, f, f_global, gin=0):
        return

    def collect(self, f):
        """collector agents only: runs BEFORE the substep kernels of frame f (the reference's agent.act precedes p2g, MPM:521)"""
        return

    @property
    def action_dim(self):
        return self.action_dims[-1]

    @property
    def state_dim(self):
        return sum(e.state_dim for e in self.effectors)

    def set_action(self, s, s_global, n_substeps, action):
        action = np.asarray(action).reshape(-1)
        assert len(action) == self.action_dims[-1], 'Action length does not match agent specifications.'
        for i in range(self.n_effectors):
            self.effectors[i].set_action(s, s_global, n_substeps, action[self.action_dims[i]:self.action_dims[i + 1]])

    def set_action_grad(self, s, s_global, n_substeps, action):
        action = np.asarray(action).reshape(-1)
        assert len(action) == self.action_dims[-1]
        for i in range(self.n_effectors - 1, -1, -1):
            self.effectors[i].set_action_grad(s, s_global, n_substeps, action[self.action_dims[i]:self.action_dims[i + 1]])

    def apply_action_p(self, action_p):
        action_p = np.asarray(action_p).reshape(-1)
        for i in range(self.n_effectors):
            self.effectors[i].apply_action_p(action_p[self.action_dims[i]:self.action_dims[i + 1]])

    def apply_action_p_grad(self, action_p):
        action_p = np.asarray(action_p).reshape(-1)
        for i in range(self.n_effectors - 1, -1, -1):
            self.effectors[i].apply_action_p_grad(action_p[self.action_dims[i]:self.action_dims[i + 1]])

    def get_grad(self, n):
        grads = [g for g in (e.get_action_grad(0, n) for e in self.effectors) if g is not None]
        return np.concatenate(grads, axis=1)

    def get_grad_device(self, n):
        """get_grad as a device tensor (float32 [n + 1, action_dim]): feeds optimizer.TrainablePolicy.optimize without leaving the GPU"""
        grads = [g for g in (e.get_action_grad_device(0, n) for e in self.effectors) if g is not None]
        return torch.cat(grads, dim=1).contiguous()

    def move(self, f):
        for e in self.effectors:
            e.move(f)

    def move_grad(self, f):
        for e in reversed(self.effectors):
            e.move_grad(f)

    def get_state(self, f):
        return [e.get_state(f) for e in self.effectors]

    def set_state(self, f, state):
        for e, s in zip(self.effectors, state):
            e.set_state(f, s)

    def copy_frame(self, source, target):
        for e in self.effectors:
            e.copy_frame(source, target)

    def copy_grad(self, source, target):
        for e in self.effectors:
            e.copy_grad(source, target)

    def reset_grad_till_frame(self, f):
        for e in self.effectors:
            e.reset_grad_till_frame(f)

    def get_ckpt(self, ckpt_name=None):
        return [e.get_ckpt() for e in self.effectors]

    def set_ckpt(self, ckpt=None, ckpt_name=None):
        for e, c in zip(self.effectors, ckpt):
            e.set_ckpt(c)


class AgentInjector(Agent):
    """Agent with one Injector (agents/agent_injector.py)."""

    def build(self, sim):
        super().build(sim)
        assert self.n_effectors == 1
        assert isinstance(self.effectors[0], Injector)
        self.injector = self.effectors[0]
        self.injector.set_act_range(sim.get_used(0))
        self.injector.finalize()

    def act(self, f, f_global):
        self.injector.act(f, f_global)

    def act_grad(self, f, f_global, gin=0):
        self.injector.act_grad(f, f_global, gin)


class AgentRigid(Agent):
    """Agent with one Rigid (agents/agent_rigid.py): its mesh collides with the material (collide_type particle/grid/both)."""

    def build(self, sim):
        super().build(sim)
        assert self.n_effectors == 1
        assert isinstance(self.effectors[0], Rigid)
        self.rigid = self.effectors[0]
        assert self.rigid.mesh is not None, 'Rigid effector without a mesh'
        if sim.has_particles:   # a particle-free scene (e.g. smoke only) has no MPM handle to register colliders with
            sim.register_colliders()


class AgentIceCreamDynamic(Agent):
    """Static (Ball)Injector + controllable Rigid cone (agents/agent_icecreamdynamic.py): injects while
    f_global < inject_till, the cone collides only above y = 0.25, actions are clipped to [-1, 1] / [0.05, 0.95]."""
    collide_y_min = 0.25

    def __init__(self, inject_till=0, **kwargs):
        super().__init__(**kwargs)
        self.inject_till = inject_till

    def build(self, sim):
        super().build(sim)
        assert self.n_effectors == 2
        assert isinstance(self.effectors[0], Injector)
        self.injector = self.effectors[0]
        assert isinstance(self.effectors[1], Rigid)
        self.rigid = self.effectors[1]
        self.injector.set_act_range(sim.get_used(0))
        self.injector.finalize()
        if sim.has_particles:   # a particle-free scene (e.g. smoke only) has no MPM handle to register colliders with
            sim.register_colliders()

    def act(self, f, f_global):
        if f_global < self.inject_till:
            self.injector.act(f, f_global)
        else:
            self.injector.act_id[f + 1] = self.injector.act_id[f]

    def act_grad(self, f, f_global, gin=0):
        if f_global < self.inject_till:
            self.injector.act_grad(f, f_global, gin)

    @property
    def action_dim(self):
        return self.rigid.action_dim

    @property
    def state_dim(self):
        return self.rigid.state_dim

    def set_action(self, s, s_global, n_substeps, action):
        action = np.asarray(action).reshape(-1).clip(-1, 1)
        assert len(action) == self.rigid.action_dim
        # `move` runs for every effector (agent_icecreamdynamic.py:70-72); the injector's action buffer stays zero
        self.injector.set_action(s, s_global, n_substeps, np.zeros(max(self.injector.action_dim, 1)))
        self.rigid.set_action(s, s_global, n_substeps, action)

    def set_action_grad(self, s, s_global, n_substeps, action):
        self.rigid.set_action_grad(s, s_global, n_substeps, action)

    def apply_action_p(self, action_p):
        self.rigid.apply_action_p(np.asarray(action_p).reshape(-1).clip(0.05, 0.95))

    def apply_action_p_grad(self, action_p):
        self.rigid.apply_action_p_grad(np.asarray(action_p).reshape(-1).clip(0.05, 0.95))

    def get_grad(self, n):
        return self.rigid.get_action_grad(0, n)

    def get_grad_device(self, n):
        return self.rigid.get_action_grad_device(0, n).contiguous()


class _Collector:
    """collector_act_kernel (agents/agent_pouring.py:31-41, agents/agent_jetbot.py:30-40): used particles outside
    `collector_boundary` are parked at NOWHERE and leave the simulation.  `material`: None = every material."""

    def _setup_collector(self, collector_boundary, material=None):
        self.collector_boundary = create_boundary(**dict(collector_boundary))
        self._collector_material = material
        self._collector = None

    def _build_collector(self, sim):
        b = self.collector_boundary
        c = _lib.FmpmCollector()
        c.boundary_type = b.type_id
        c.lower = (C.c_float * 3)(*[float(v) for v in b.lower]); c.upper = (C.c_float * 3)(*[float(v) for v in b.upper])
        c.cyl_center = (C.c_float * 2)(*[float(v) for v in b.xz_center]); c.cyl_radius = float(b.xz_radius)
        c.row_mask = 0xffffffff if self._collector_material is None else sim.material_row_mask(self._collector_material)
        self._collector = c

    def collect(self, f):
        sim = self.sim
        if sim.has_particles:
            sim._ck(sim._lib.fmpm_collect(sim._h, f, C.byref(self._collector), sim._stream()), 'fmpm_collect')


class AgentPouring(_Collector, AgentRigid):
    """Agent with one Rigid and a collector (agents/agent_pouring.py): the mesh collides at grid AND particle level."""

    def __init__(self, collector_boundary, **kwargs):
        kwargs.pop('collide_type', None)
        super().__init__(collide_type='both', **kwargs)
        self._setup_collector(collector_boundary, material=None)

    def build(self, sim):
        super().build(sim)
        self._build_collector(sim)


class AgentJetBot(_Collector, AgentInjector):
    """Agent with one Injector and a collector of WATER particles (agents/agent_jetbot.py)."""

    def __init__(self, collector_boundary, **kwargs):
        super().__init__(**kwargs)
        self._setup_collector(collector_boundary, material=WATER)

    def build(self, sim):
        super().build(sim)
        self._build_collector(sim)


class AgentCirculation(Agent):
    """agent of the air-circulation env (agents/agent_circulation.py:8-23): one AirCon effector, no collision with the MPM particles."""

    def build(self, sim):
        super().build(sim)
        assert self.n_effectors == 1 and isinstance(self.effectors[0], AirCon)
        self.aircon = self.effectors[0]
        if getattr(sim, 'smoke_field', None) is not None:
            sim.smoke_field.bind_aircon(self.aircon)   # the effector's device arrays exist now
