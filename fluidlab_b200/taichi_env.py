"""Engine facade with the reference's `TaichiEnv` interface (fluidlab/fluidengine/taichi_env.py:17-222) so the
reference's envs/ and optimizer/ drive the B200 simulator unchanged: setup_agent / setup_boundary / add_body /
setup_loss / build / step / step_grad / get_state / set_state / reset_grad / loss and action accessors.
Renderers are outside this path; `setup_smoke_field` attaches the B200 smoke solver (smoke.py)."""
import numpy as np
from .simulator import MPMSimulator
from .bodies import Bodies
from .macros import DTYPE_NP
from .meshes import Statics
from . import agents as _agents


class _Cfg(dict):
    """dict with attribute access: accepts a yacs CfgNode (it is a dict) or a plain dict for agent configs."""
    __getattr__ = dict.get


class TaichiEnv:
    def __init__(self, dim=3, quality=1, particle_density=1e6, max_substeps_local=50, max_substeps_global=100000, horizon=100,
                 ckpt_dest='gpu', gravity=(0.0, -10.0, 0.0), device=None, sort_every=1):
        self.particle_density = particle_density
        self.dim = dim
        self.max_substeps_local = max_substeps_local
        self.max_substeps_global = max_substeps_global
        self.horizon = horizon
        self.ckpt_dest = ckpt_dest
        self.t = 0
        self.simulator = MPMSimulator(dim=dim, quality=quality, horizon=horizon, max_substeps_local=max_substeps_local,
                                      max_substeps_global=max_substeps_global, gravity=gravity, ckpt_dest=ckpt_dest,
                                      device=device, sort_every=sort_every)
        self.agent = None
        self.statics = Statics()
        self.particle_bodies = Bodies(dim=dim, particle_density=particle_density)
        self.renderer = None
        self.loss = None
        self.smoke_field = None

    def setup_agent(self, agent_cfg):  # taichi_env.py:59-75
        agent_cfg = _Cfg(agent_cfg)
        cls = getattr(_agents, agent_cfg['type'])
        self.agent = cls(max_substeps_local=self.max_substeps_local, max_substeps_global=self.max_substeps_global,
                         max_action_steps_global=self.horizon, ckpt_dest=self.ckpt_dest, **dict(agent_cfg.get('params', {}) or {}))
        for effector_cfg in agent_cfg['effectors']:
            effector_cfg = _Cfg(effector_cfg)
            self.agent.add_effector(type=effector_cfg['type'], params=effector_cfg['params'], mesh_cfg=effector_cfg.get('mesh', None),
                                    boundary_cfg=effector_cfg['boundary'])

    def setup_renderer(self, **kwargs):
        self.renderer = None  # visualisation is out of scope

    def setup_boundary(self, **kwargs):
        self.simulator.setup_boundary(**kwargs)

    def setup_smoke_field(self, **kwargs):  # taichi_env.py:95-100
        from .smoke import SmokeField
        self.smoke_field = SmokeField(dim=self.dim, ckpt_dest=self.ckpt_dest, **kwargs)

    def add_static(self, **kwargs):  # taichi_env.py:89-90
        self.statics.add_static(**kwargs)  # statics without dynamics are visual only and never reach the kernels

    def add_body(self, **kwargs):
        self.particle_bodies.add_body(**kwargs)

    def setup_loss(self, loss_cls, **kwargs):
        self.loss = loss_cls(max_loss_steps=self.horizon, **kwargs)

    def build(self):  # taichi_env.py:108-134
        self.particles = self.particle_bodies.get()
        if self.particles is not None:
            self.n_particles = len(self.particles['x']); self.has_particles = True
        else:
            self.n_particles = 0; self.has_particles = False
        self.simulator.build(self.agent, self.smoke_field, self.statics, self.particles)
        if self.agent is not None:
            self.agent.build(self.simulator)
        if self.smoke_field is not None:   # taichi_env.py:125-126
            self.smoke_field.build(self.simulator, self.agent)
        if self.loss is not None:
            self.loss.build(self.simulator)
        self.t = 0

    def reset_grad(self):
        self.simulator.reset_grad()
        if self.agent is not None:
            self.agent.reset_grad()
        if self.smoke_field is not None:   # taichi_env.py:142-143
            self.smoke_field.reset_grad()
        if self.loss is not None:
            self.loss.reset_grad()

    def enable_grad(self):
        self.simulator.enable_grad()

    def disable_grad(self):
        self.simulator.disable_grad()

    @property
    def grad_enabled(self):
        return self.simulator.grad_enabled

    def get_state_RL(self):
        return self.simulator.get_state_RL()

    def get_obs_RL(self, n_obs_ptcls_per_body=200):
        """FluidEnv._get_obs's vector (envs/fluid_env.py:99-125) assembled on the device: one small D2H instead of the full state"""
        return self.simulator.get_obs_RL(n_obs_ptcls_per_body)

    def step(self, action=None):  # taichi_env.py:165-174
        if action is not None:
            assert self.agent is not None, 'Environment has no agent to execute action.'
            action = np.array(action).astype(DTYPE_NP)
        self.simulator.step(action=action)
        if self.loss:
            self.loss.step()
        self.t += 1

    def step_grad(self, action=None):  # taichi_env.py:176-183
        if self.loss:
            self.loss.step_grad()
        if action is not None:
            assert self.agent is not None, 'Environment has no agent to execute action.'
            action = np.array(action).astype(DTYPE_NP)
        self.simulator.step_grad(action=action)

    def get_step_loss(self):
        assert self.loss is not None
        return self.loss.get_step_loss()

    def get_final_loss(self):
        assert self.loss is not None
        return self.loss.get_final_loss()

    def get_final_loss_grad(self):
        assert self.loss is not None
        self.loss.get_final_loss_grad()

    def get_state(self):
        return {'state': self.simulator.get_state(), 'grad_enabled': self.grad_enabled}

    def set_state(self, state, grad_enabled=False):  # taichi_env.py:203-214
        self.t = 0
        self.simulator.cur_substep_global = 0
        self.simulator.set_state(0, state)
        if grad_enabled:
            self.enable_grad()
        else:
            self.disable_grad()
        if self.loss:
            self.loss.reset()

    def apply_agent_action_p(self, action_p):
        assert self.agent is not None, 'Environment has no agent to execute action.'
        self.agent.apply_action_p(action_p)

    def apply_agent_action_p_grad(self, action_p):
        assert self.agent is not None, 'Environment has no agent to execute action.'
        self.agent.apply_action_p_grad(action_p)
