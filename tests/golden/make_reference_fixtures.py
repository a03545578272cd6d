"""Fixtures produced by the REAL reference code (run in the build container only; /root/reference does not exist on the GPU box).

The host-side, NumPy-only parts of zhouxian/FluidLab that feed the hot path — the material tables of `fluidlab/configs/macros.py` and
the particle samplers of `fluidlab/fluidengine/bodies/bodies.py` — import fine once `taichi` (used there only for a dtype alias) and the
rendering / mesh dependencies are stubbed.  This script runs them UNMODIFIED and stores what `fluidlab_b200.macros` / `fluidlab_b200.bodies`
must reproduce: the tables verbatim and, per sampled body, the particle count, three checksums and the first / last 16 rows.
(The Taichi kernels themselves cannot run: the simulation path stays pinned by the oracle only)

    python tests/golden/make_reference_fixtures.py        # writes tests/golden/reference_host_fixtures.npz
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('FLUIDLAB_REFERENCE', '/root/reference')

# every sampler configuration the shipped envs use (envs/*.py add_body calls), plus the natural fillings
BODY_CASES = [
    ('latteart_coffee', dict(type='cylinder', center=(0.5, 0.55, 0.5), height=0.1, radius=0.42, material='COFFEE')),
    ('nowhere', dict(type='nowhere', n_particles=1000, material='MILK')),
    ('cube_random', dict(type='cube', lower=(0.2, 0.3, 0.2), upper=(0.45, 0.5, 0.6), material='WATER')),
    ('cube_size_grid', dict(type='cube', lower=(0.3, 0.3, 0.3), size=(0.2, 0.1, 0.15), material='ELASTIC', filling='grid')),
    ('cube_natural_euler', dict(type='cube', lower=(0.45, 0.45, 0.45), size=(0.1, 0.1, 0.1), euler=(45.0, 45.0, 45.0), material='RIGID_HEAVY', filling='natural')),
    ('cylinder_natural', dict(type='cylinder', center=(0.5, 0.4, 0.5), height=0.12, radius=0.2, material='WATER', filling='natural')),
    ('cylinder_grid', dict(type='cylinder', center=(0.5, 0.4, 0.5), height=0.12, radius=0.2, material='WATER', filling='grid')),
    ('ball_random', dict(type='ball', center=(0.5, 0.5, 0.5), radius=0.1, material='ICECREAM')),
    ('ball_natural', dict(type='ball', center=(0.4, 0.6, 0.5), radius=0.07, material='WATER', filling='natural')),
    ('cube_euler_random', dict(type='cube', lower=(0.3, 0.3, 0.3), upper=(0.5, 0.4, 0.5), euler=(0.0, -75.0, 10.0), material='ELASTIC')),
]
TABLES = ('MU', 'LAMDA', 'RHO', 'MAT_CLASS', 'FRICTION')


MESH_CASES = [   # (name, baked sdf pickle, pos, euler, scale): the shipped collider configs (envs/configs/agent_*.yaml, envs/*_env.py add_static)
    ('stirrer', 'stirrer-128.sdf', (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (0.6, 0.4, 0.6)),
    ('cone_tip', 'cone_tip-128.sdf', (0.0, 0.0, 0.0), (-90.0, 0.0, 30.0), (0.726, 0.726, 0.726)),
    ('glass', 'glass-128.sdf', (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (0.75, 0.65, 0.75)),
    ('laddle', 'laddle-128.sdf', (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (0.35, 0.45, 0.35)),
    ('static_tank', 'bowl-128.sdf', (0.5, 0.4, 0.5), (0.0, 45.0, 10.0), (1.2, 0.8, 1.2)),
]
EFFECTOR_CASES = [   # constructor kwargs of effectors/effector.py:18-52 and effectors/injector.py:12-60,220-240 (no Taichi arithmetic involved)
    ('injector_local', 'Injector', dict(radius=0.0075, flux=2, init_pos=(0.5, 0.5, 0.5), init_euler=(10.0, 20.0, 30.0), action_dim=3, inject_v=(0.0, -3.0, 0.0),
                                        locally_random=True)),
    ('injector_global', 'Injector', dict(radius=0.015, flux=4, init_pos=(0.5, 0.8, 0.5), init_euler=(0.0, 0.0, 0.0), action_dim=6, inject_v=(-3.0, 0.0, 0.0),
                                         inject_p=(-0.07, 0.0, 0.0), locally_random=False)),
    ('ball_injector', 'BallInjector', dict(radius=0.035, flux=4, init_pos=(0.5, 0.6, 0.5), action_dim=3, inject_v=(0.0, -0.4, 0.0), locally_random=True)),
]
EFFECTOR_COMMON = dict(max_substeps_local=50, max_substeps_global=400, max_action_steps_global=40, ckpt_dest='cpu')


def load_reference():
    class _Obj:   # stands in for every Taichi object (fields, vectors, decorators); fields remember what from_numpy() received
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            if len(a) == 1 and not k and callable(a[0]) and not isinstance(a[0], _Obj):
                return a[0]          # decorator use: @ti.kernel, @ti.func, @ti.data_oriented
            return _Obj()

        def __getattr__(self, k):
            if k.startswith('__'):
                raise AttributeError(k)
            if k == 'grad_for':          # @ti.ad.grad_for(fn) is a decorator FACTORY
                return lambda *_: (lambda g: g)
            o = _Obj(); object.__setattr__(self, k, o); return o

        def from_numpy(self, a):
            object.__setattr__(self, 'np_value', np.array(a))

        def __getitem__(self, k):
            return _Obj()

        def __setitem__(self, k, v):
            pass

        def __iter__(self):
            return iter(())

    class _Stub(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith('__'):
                raise AttributeError(k)
            o = _Obj(); setattr(self, k, o); return o
    ti = _Stub('taichi'); ti.f32, ti.f64, ti.i32 = 'f32', 'f64', 'i32'
    ti.static = lambda x: x          # loops over ti.static(range(n)) keep iterating
    sys.modules['taichi'] = ti
    for name in ('trimesh', 'yacs', 'yacs.config', 'gym', 'gym.spaces', 'mesh_to_sdf', 'skimage', 'skimage.measure', 'matplotlib', 'matplotlib.pyplot',
                 'imageio', 'pyrender', 'open3d', 'cv2', 'OpenGL', 'OpenGL.GL', 'pyglet'):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Stub(name)
    sys.path.insert(0, REF)
    sys.modules['fluidlab.fluidengine.renderers.gl_renderer_src'] = _Stub('fluidlab.fluidengine.renderers.gl_renderer_src')   # compiled FleX binding
    macros = importlib.import_module('fluidlab.configs.macros')
    spec = importlib.util.spec_from_file_location('ref_bodies', os.path.join(REF, 'fluidlab/fluidengine/bodies/bodies.py'))
    bodies = importlib.util.module_from_spec(spec); spec.loader.exec_module(bodies)
    return macros, bodies


TRACE_T, TRACE_ACTIONS = 20, [[0.1, 0.0, 0.0], [0.0, 0.2, 0.0], [0.0, 0.0, 0.3], None, None]   # 5 steps over a 2-step ring: two wrap-arounds


class TraceAgent:
    """records the calls the simulator's step / step_grad / checkpoint logic makes on its agent (agents/agent.py API)"""

    def __init__(self, trace, sim_ref):
        self.trace, self.sim_ref = trace, sim_ref

    def _ev(self, name, *args):
        self.trace.append([name, int(self.sim_ref().cur_substep_global)] + [None if a is None else int(a) for a in args])

    def set_action(self, s, s_global, n_substeps, action):
        self._ev('agent.set_action', s, s_global, n_substeps)

    def set_action_grad(self, s, s_global, n_substeps, action):
        self._ev('agent.set_action_grad', s, s_global, n_substeps)

    def copy_frame(self, a, b):
        self._ev('agent.copy_frame', a, b)

    def copy_grad(self, a, b):
        self._ev('agent.copy_grad', a, b)

    def reset_grad_till_frame(self, f):
        self._ev('agent.reset_grad_till_frame', f)

    def get_ckpt(self, ckpt_name=None):
        self._ev('agent.get_ckpt')
        return {}

    def set_ckpt(self, ckpt=None, ckpt_name=None):
        self._ev('agent.set_ckpt')


def drive(sim):
    """the call sequence of optimizer/solver.py:23-59 at simulator level: forward all steps, then backward in reverse order"""
    sim.enable_grad()
    for a in TRACE_ACTIONS:
        sim.step(None if a is None else np.array(a, dtype=np.float32))
    for a in reversed(TRACE_ACTIONS):
        sim.step_grad(None if a is None else np.array(a, dtype=np.float32))


class Recorder:
    """stand-in for simulator / agent / loss objects: every method call is appended to the shared trace"""

    def __init__(self, name, trace):
        object.__setattr__(self, '_name', name); object.__setattr__(self, '_trace', trace)

    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)

        def call(*a, **kw):
            def brief(v):
                if v is None or isinstance(v, (int, float, bool, str)):
                    return v
                if isinstance(v, dict):
                    return 'dict'
                arr = np.asarray(v, dtype=np.float64).reshape(-1)
                return [round(float(q), 6) for q in arr] if arr.size <= 12 else 'arr%d' % arr.size
            args = [brief(v) for v in list(a) + [kw[q] for q in sorted(kw)]]
            self._trace.append([f'{self._name}.{k}'] + args)
        return call

    def __setattr__(self, k, v):
        self._trace.append([f'{self._name}.{k}=', v if isinstance(v, (int, float, bool)) else 'obj'])

    def __bool__(self):
        return True


class MockEffector(Recorder):
    def __init__(self, name, trace, action_dim, tag):
        Recorder.__init__(self, name, trace)
        object.__setattr__(self, 'action_dim', action_dim); object.__setattr__(self, 'tag', tag); object.__setattr__(self, 'state_dim', 7)

    def get_action_grad(self, s, n):
        self._trace.append([f'{self._name}.get_action_grad', s, n])
        return None if self.action_dim == 0 else np.full((n + 1, self.action_dim), float(self.tag))

    def get_state(self, f):
        self._trace.append([f'{self._name}.get_state', f])
        return np.full(7, float(self.tag))


def drive_agent(agent, trace):
    """Agent (agents/agent.py:68-131): how actions are split over the effectors and in which order the effectors are visited"""
    agent.effectors = [MockEffector('eff0', trace, 3, 1), MockEffector('eff1', trace, 0, 2), MockEffector('eff2', trace, 6, 3)]
    agent.action_dims = [0, 3, 3, 9]
    agent.n_effectors = 3
    act = np.arange(9) * 0.5
    agent.set_action(1, 4, 10, act); agent.set_action_grad(1, 4, 10, act)
    agent.apply_action_p(act + 1); agent.apply_action_p_grad(act + 1)
    g = agent.get_grad(2)
    trace.append(['get_grad.result', list(np.asarray(g).shape), [float(v) for v in np.asarray(g)[0]]])
    agent.move(7); agent.move_grad(7)
    st = agent.get_state(3)
    trace.append(['get_state.result', [float(np.asarray(s)[0]) for s in st]])
    agent.set_state(3, [np.arange(7.0), np.arange(7.0) + 1, np.arange(7.0) + 2])
    agent.copy_frame(20, 0); agent.copy_grad(0, 20); agent.reset_grad_till_frame(20); agent.reset_grad()
    trace.append(['dims', int(agent.action_dim), int(agent.state_dim)])


def reference_agent_trace():
    ag = importlib.import_module('fluidlab.fluidengine.agents.agent')
    trace = []
    a = ag.Agent(max_substeps_local=20, max_substeps_global=1000, max_action_steps_global=50, ckpt_dest='cpu')
    drive_agent(a, trace)
    return trace


def drive_env(env):
    """optimizer/solver.py:23-59 at TaichiEnv level"""
    a = np.array([0.1, 0.2, 0.3]); ap = np.array([0.5, 0.6, 0.5])
    env.set_state({'x': 0}, grad_enabled=True)
    env.apply_agent_action_p(ap)
    for act in (a, a, None):
        env.step(act)
    env.get_final_loss()
    env.reset_grad(); env.get_final_loss_grad()
    for act in (None, a, a):
        env.step_grad(act)
    env.apply_agent_action_p_grad(ap)
    env.set_state({'x': 0}, grad_enabled=False)
    env.step(a)
    env.get_state_RL()


def make_env_with_recorders(cls, trace):
    env = cls.__new__(cls)
    env.simulator, env.agent, env.loss = Recorder('simulator', trace), Recorder('agent', trace), Recorder('loss', trace)
    env.smoke_field, env.renderer, env.t = None, None, 0
    return env


def reference_env_trace():
    te = importlib.import_module('fluidlab.fluidengine.taichi_env')
    trace = []
    drive_env(make_env_with_recorders(te.TaichiEnv, trace))
    return trace


def reference_temporal_range_schedule(losses):
    """ShapeMatchingLoss.expand_temporal_range (losses/shapematching_loss.py:110-130), the reference's own code, fed a loss sequence"""
    sm = importlib.import_module('fluidlab.fluidengine.losses.shapematching_loss')
    L = sm.ShapeMatchingLoss.__new__(sm.ShapeMatchingLoss)
    L.temporal_range_type, L.temporal_range, L.best_loss, L.plateau_count, L.inf = 'expand', [0, 50], 1e8, 0, 1e8
    L.plateau_thresh, L.plateau_count_limit, L.temporal_expand_speed, L.max_loss_steps = [0.01, 0.5], 5, 50, 220
    out = []
    import io, contextlib
    for v in losses:
        L.total_loss = {None: float(v)}
        with contextlib.redirect_stdout(io.StringIO()):
            L.expand_temporal_range()
        out.append([L.temporal_range[1], L.plateau_count, float(L.best_loss)])
    return out


LOSS_SEQUENCE = [100.0, 90.0, 89.5, 89.4, 89.39, 89.38, 89.37, 89.36, 120.0, 80.0, 79.9, 79.95, 79.9, 79.9, 79.9, 79.9, 300.0, 299.0, 299.0, 299.0,
                 299.0, 299.0, 298.0, 100.0, 99.9, 99.9, 99.9, 99.9, 99.9, 99.9]


def reference_step_trace():
    """run the REAL MPMSimulator.step / step_ / step_grad / memory_to_cache / memory_from_cache (MPM:721-912) with every Taichi kernel
    replaced by a recorder: the orchestration (ring indices, checkpoint names, re-simulation, agent calls) is pure Python."""
    sim_mod = importlib.import_module('fluidlab.fluidengine.simulators.mpm_simulator')
    S = sim_mod.MPMSimulator
    trace = []
    sim = S(dim=3, quality=0.25, gravity=(0.0, -10.0, 0.0), horizon=10, max_substeps_local=TRACE_T, max_substeps_global=1000, ckpt_dest='cpu')
    rec = lambda name: (lambda *a, **k: trace.append([name, int(sim.cur_substep_global)] + [int(v) if isinstance(v, (int, np.integer)) else None for v in a[:2]]))
    sim.substep = lambda f, none: trace.append(['substep', int(sim.cur_substep_global), int(f), int(bool(none))])
    sim.substep_grad = lambda f, none: trace.append(['substep_grad', int(sim.cur_substep_global), int(f), int(bool(none))])
    for name in ('copy_frame', 'copy_grad', 'reset_grad_till_frame', 'readframe', 'setframe'):
        setattr(sim, name, rec(name))
    import weakref
    agent = TraceAgent(trace, weakref.ref(sim))
    n = 4
    sim.build(agent, None, [], dict(x=np.full((n, 3), 0.5), used=np.ones(n), mat=np.zeros(n), rho=np.ones(n), body_id=np.zeros(n), bodies={'n': 1}))
    drive(sim)
    return trace


def summarize(x):
    x = np.asarray(x, dtype=np.float64)
    return dict(n=len(x), cs=np.array([x.sum(), np.abs(x).sum(), (x * np.arange(1, x.size + 1).reshape(x.shape)).sum()]), head=x[:16].copy(), tail=x[-16:].copy())


def main():
    RM, RB = load_reference()
    out = {}
    names = [k for k in dir(RM) if k.isupper() and isinstance(getattr(RM, k), int) and k not in ('DTYPE_NP',)]
    out['int_names'] = np.array(names); out['int_values'] = np.array([getattr(RM, k) for k in names], dtype=np.int64)
    for t in TABLES:
        tab = getattr(RM, t)
        keys = sorted(tab)
        out[f'tab_{t}_keys'] = np.array(keys, dtype=np.int64); out[f'tab_{t}_vals'] = np.array([tab[k] for k in keys], dtype=np.float64)
    out['NOWHERE'] = np.array(RM.NOWHERE, dtype=np.float64); out['EPS'] = np.float64(RM.EPS)
    np.random.seed(12345)   # the samplers re-seed to 0 internally and must restore this state
    for name, kw in BODY_CASES:
        kw = dict(kw); kw['material'] = getattr(RM, kw['material'])
        b = RB.Bodies(dim=3, particle_density=1e6)
        b.add_body(**kw)
        g = b.get()
        s = summarize(g['x'])
        out[f'body_{name}_n'] = s['n']; out[f'body_{name}_cs'] = s['cs']; out[f'body_{name}_head'] = s['head']; out[f'body_{name}_tail'] = s['tail']
        out[f'body_{name}_used'] = int(np.asarray(g['used']).sum()); out[f'body_{name}_rho'] = float(np.asarray(g['rho'])[0])
    out['rng_after'] = np.random.uniform(size=4)   # the global RNG stream must be untouched by add_body
    # two bodies in one container: ids and concatenation order
    b = RB.Bodies(dim=3, particle_density=1e6)
    b.add_body(type='nowhere', n_particles=60000, material=RM.MILK)
    b.add_body(type='cylinder', center=(0.5, 0.55, 0.5), height=0.1, radius=0.42, material=RM.COFFEE)
    g = b.get()
    out['latteart_n'] = len(g['x']); out['latteart_used'] = int(np.asarray(g['used']).sum())
    out['latteart_body_id_counts'] = np.bincount(np.asarray(g['body_id']).astype(np.int64))
    # meshes: the reference's own Mesh.init_transform (meshes/mesh.py:97-127) composes T_mesh_to_voxels <- T_file @ inv(T_init); run it on an
    # instance that carries only what those lines read (vertices / colours are rendering data and replaced by one dummy vertex)
    import pickle as pkl
    mesh_mod = importlib.import_module('fluidlab.fluidengine.meshes.mesh')
    for name, fn, pos, euler, scale in MESH_CASES:
        path = os.path.join(REF, 'fluidlab/assets/meshes/processed', fn)
        if not os.path.exists(path):
            continue
        sdf = pkl.load(open(path, 'rb'))
        m = mesh_mod.Mesh.__new__(mesh_mod.Mesh)
        m.scale, m.pos, m.euler = scale, pos, np.array(euler)   # what Mesh.__init__ stores (meshes/mesh.py:16-39, eval_str'ed tuples)
        m.raw_vertices = np.zeros((1, 3), np.float32); m.raw_vertex_normals_np = np.zeros((1, 3), np.float32)
        m.faces_np = np.zeros(3, np.int32); m.colors_np = np.zeros((1, 4), np.float32); m.n_vertices, m.n_faces = 1, 3
        m.has_dynamics = True
        m.sdf_voxels_np = sdf['voxels'].astype(np.float32); m.T_mesh_to_voxels_np = sdf['T_mesh_to_voxels'].astype(np.float32)
        m.init_transform()
        out[f'mesh_{name}_T_file'] = np.asarray(sdf['T_mesh_to_voxels'], dtype=np.float64)
        out[f'mesh_{name}_T'] = np.asarray(m.T_mesh_to_voxels_np, dtype=np.float64)
        out[f'mesh_{name}_res'] = int(sdf['voxels'].shape[0]); out[f'mesh_{name}_vox_cs'] = np.float64(np.asarray(sdf['voxels'], dtype=np.float64).sum())
    # effectors: initial pose quaternion (scipy 'zyx' euler convention, effector.py:40-44) and the injectors' pre-drawn random tables
    inj = importlib.import_module('fluidlab.fluidengine.effectors.injector')
    for name, cls, kw in EFFECTOR_CASES:
        np.random.seed(777)
        e = getattr(inj, cls)(**EFFECTOR_COMMON, **kw)
        out[f'eff_{name}_init_state'] = np.asarray(e.init_state, dtype=np.float64)
        rv = getattr(e, 'random_vector_np', None)
        if rv is None or isinstance(rv, list):
            rv = e.random_vector.np_value
        out[f'eff_{name}_rv_shape'] = np.array(rv.shape); out[f'eff_{name}_rv_head'] = np.asarray(rv, dtype=np.float64).reshape(-1, 3)[:24]
        out[f'eff_{name}_rv_sum'] = np.float64(np.asarray(rv, dtype=np.float64).sum())
        out[f'eff_{name}_rng_after'] = np.random.uniform(size=2)
    import json
    out['step_trace_json'] = np.array(json.dumps(reference_step_trace()))
    out['env_trace_json'] = np.array(json.dumps(reference_env_trace()))
    out['agent_trace_json'] = np.array(json.dumps(reference_agent_trace()))
    out['temporal_range_schedule'] = np.array(reference_temporal_range_schedule(LOSS_SEQUENCE), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, 'reference_host_fixtures.npz'), **out)
    print('wrote', os.path.getsize(os.path.join(HERE, 'reference_host_fixtures.npz')), 'bytes;', {k: int(out[f'body_{k}_n']) for k, _ in BODY_CASES})


if __name__ == '__main__':
    main()
