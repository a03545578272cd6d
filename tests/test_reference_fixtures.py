"""Host-side pieces pinned against the REAL reference: `tests/golden/reference_host_fixtures.npz` was produced by running the
unmodified `fluidlab/configs/macros.py` and `fluidlab/fluidengine/bodies/bodies.py` of zhouxian/FluidLab
(tests/golden/make_reference_fixtures.py, build container only).  fluidlab_b200.macros / fluidlab_b200.bodies must reproduce them."""
import os
import sys
import numpy as np

from fluidlab_b200 import macros as M
from fluidlab_b200.bodies import Bodies

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
sys.path.insert(0, G)
D = np.load(os.path.join(G, 'reference_host_fixtures.npz'))


def test_material_tables_equal_the_reference_macros():
    for name, val in zip(D['int_names'], D['int_values']):
        name = str(name)
        if hasattr(M, name) and isinstance(getattr(M, name), int):
            assert getattr(M, name) == int(val), name
    for must in ('WATER', 'ELASTIC', 'MILK', 'COFFEE', 'ICECREAM', 'RIGID', 'RIGID_HEAVY', 'MAT_LIQUID', 'MAT_ELASTIC', 'MAT_RIGID', 'MAT_PLASTO_ELASTIC'):
        assert hasattr(M, must), must
    for t in ('MU', 'LAMDA', 'RHO', 'MAT_CLASS', 'FRICTION'):
        tab = getattr(M, t)
        for k, v in zip(D[f'tab_{t}_keys'], D[f'tab_{t}_vals']):
            assert float(tab[int(k)]) == float(v), (t, int(k), tab[int(k)], v)
        assert len(tab) == len(D[f'tab_{t}_keys']), t
    assert list(M.NOWHERE) == list(D['NOWHERE'])


def test_samplers_reproduce_the_reference_particle_sets():
    import make_reference_fixtures as mk
    np.random.seed(12345)
    for name, kw in mk.BODY_CASES:
        kw = dict(kw); kw['material'] = getattr(M, kw['material'])
        b = Bodies(dim=3, particle_density=1e6)
        b.add_body(**kw)
        g = b.get()
        s = mk.summarize(g['x'])
        assert s['n'] == int(D[f'body_{name}_n']), (name, s['n'])
        assert int(np.asarray(g['used']).sum()) == int(D[f'body_{name}_used']) and float(np.asarray(g['rho'])[0]) == float(D[f'body_{name}_rho']), name
        for k in ('head', 'tail'):
            assert np.abs(s[k] - D[f'body_{name}_{k}']).max() <= 1e-15 * max(1.0, np.abs(s[k]).max()), (name, k)
        assert np.allclose(s['cs'], D[f'body_{name}_cs'], rtol=1e-13, atol=1e-9), (name, s['cs'], D[f'body_{name}_cs'])
    assert np.array_equal(np.random.uniform(size=4), D['rng_after']), 'add_body must leave the global NumPy RNG stream untouched'
    b = Bodies(dim=3, particle_density=1e6)
    b.add_body(type='nowhere', n_particles=60000, material=M.MILK)
    b.add_body(type='cylinder', center=(0.5, 0.55, 0.5), height=0.1, radius=0.42, material=M.COFFEE)
    g = b.get()
    assert len(g['x']) == int(D['latteart_n']) == 115480 and int(np.asarray(g['used']).sum()) == int(D['latteart_used']) == 55480
    assert np.array_equal(np.bincount(np.asarray(g['body_id']).astype(np.int64)), D['latteart_body_id_counts'])


def test_effector_initial_pose_and_injector_random_tables_equal_the_reference():
    """initial pose (scipy 'zyx' euler -> wxyz quaternion, effectors/effector.py:40-52) and the random tables the injectors draw from the
    global NumPy RNG at construction (effectors/injector.py:54-60, 220-240) — values AND RNG-stream consumption."""
    import make_reference_fixtures as mk
    from fluidlab_b200 import effectors as E
    for name, cls, kw in mk.EFFECTOR_CASES:
        np.random.seed(777)
        e = getattr(E, cls)(**mk.EFFECTOR_COMMON, **kw)
        assert np.allclose(np.asarray(e.init_state, dtype=np.float64)[:7], D[f'eff_{name}_init_state'][:7], rtol=0, atol=1e-7), name
        rv = np.asarray(e.random_vector_np)
        assert list(rv.shape) == list(D[f'eff_{name}_rv_shape']), name
        assert np.array_equal(rv.astype(np.float32).reshape(-1, 3)[:24], D[f'eff_{name}_rv_head'].astype(np.float32)), name
        assert abs(float(rv.astype(np.float64).sum()) - float(D[f'eff_{name}_rv_sum'])) < 1e-6 * rv.size, name
        assert np.array_equal(np.random.uniform(size=2), D[f'eff_{name}_rng_after']), (name, 'RNG stream consumption differs')


def test_mesh_transform_equals_the_reference_init_transform():
    """T_mesh_to_voxels <- T_file @ inv(trans_quat_to_T(pos, quat) @ scale_to_T(scale)) as the reference's own Mesh.init_transform
    (meshes/mesh.py:97-127) evaluates it in float32 for the shipped collider configs; fluidlab_b200.meshes.Mesh must give the same 4x4."""
    import make_reference_fixtures as mk
    from fluidlab_b200.meshes import Static
    n = 0
    for name, fn, pos, euler, scale in mk.MESH_CASES:
        if f'mesh_{name}_T' not in D.files:
            continue
        res = int(D[f'mesh_{name}_res'])
        m = Static(file=fn.replace('-128.sdf', '.obj'), material=M.CUP, pos=pos, euler=euler, scale=scale, has_dynamics=True,
                   sdf=dict(voxels=np.zeros((2, 2, 2), np.float32), T_mesh_to_voxels=D[f'mesh_{name}_T_file']))
        ref = D[f'mesh_{name}_T']
        err = np.abs(np.asarray(m.T_mesh_to_voxels_np, dtype=np.float64) - ref).max() / np.abs(ref).max()
        assert err < 2e-6, (name, err)        # both sides are float32 matrix products / inverses (different LAPACK call order)
        assert res == 128
        n += 1
    assert n >= 4


def test_step_and_checkpoint_orchestration_equals_the_reference():
    """The ring / checkpoint / re-simulation orchestration of MPMSimulator.step, step_, step_grad, memory_to_cache and memory_from_cache
    (MPM:721-912) is pure Python: the fixture holds the event trace of the REAL reference code run with recorder kernels (5 steps over a
    2-step ring, forward then backward).  fluidlab_b200.MPMSimulator — built without CUDA, kernels replaced by the same recorders — must
    emit the same sequence of substeps, adjoint substeps, agent calls and frame-0 restarts, with the same global substep counters.
    (Its particle adjoints live in a ping-pong pair, so the reference's particle copy_frame(0,T) / copy_grad / reset_grad_till_frame in
    memory_from_cache have no counterpart and are filtered out.)"""
    import json
    import weakref
    import torch
    import make_reference_fixtures as mk
    from fluidlab_b200.simulator import MPMSimulator, _IDENTITY
    T = mk.TRACE_T
    ref = json.loads(str(D['step_trace_json']))
    keep = lambda e: (e[0] in ('substep', 'substep_grad') or e[0].startswith('agent.') or (e[0] == 'copy_frame' and e[2:4] == [T, 0]))
    ref = [e for e in ref if keep(e)]
    trace = []
    sim = MPMSimulator.__new__(MPMSimulator)
    sim.dim, sim.n_substeps, sim.max_substeps_local, sim.max_substeps_global, sim.horizon = 3, 10, T, 1000, 10
    sim.ckpt_dest, sim.device = 'cpu', torch.device('cpu')
    sim.has_particles, sim.smoke_field, sim.sort_every, sim.use_graphs, sim.store_grids = True, None, 0, False, False
    sim.actions_buffer, sim.ckpt_ram, sim.cur_substep_global, sim.grad_enabled = [], {}, 0, False
    sim._pa, sim._pf, sim._pf8 = torch.zeros((T + 1, 4, 4, 4)), torch.zeros((T + 1, 2, 4, 4)), torch.zeros((T + 1, 4))
    sim._frame_ord = [_IDENTITY] * (T + 1)
    sim._h = sim._lib = None
    sim.substep = lambda f, none: trace.append(['substep', int(sim.cur_substep_global), int(f), int(bool(none))])
    sim.substep_grad = lambda f, none: trace.append(['substep_grad', int(sim.cur_substep_global), int(f), int(bool(none))])
    sim.copy_frame = lambda a, b: trace.append(['copy_frame', int(sim.cur_substep_global), int(a), int(b)])
    sim.agent = mk.TraceAgent(trace, weakref.ref(sim))
    mk.drive(sim)
    assert len(trace) == len(ref), (len(trace), len(ref))
    for i, (a, b) in enumerate(zip(trace, ref)):
        assert a == b, (i, a, b)


def test_taichi_env_call_sequence_equals_the_reference():
    """TaichiEnv.set_state / step / step_grad / reset_grad / get_final_loss(_grad) / apply_agent_action_p(_grad) (fluidengine/taichi_env.py:136-222)
    only sequence calls on the simulator, the agent and the loss: the fixture is the trace of the REAL class driven with recorder objects."""
    import json
    import make_reference_fixtures as mk
    from fluidlab_b200.taichi_env import TaichiEnv
    ref = json.loads(str(D['env_trace_json']))
    trace = []
    mk.drive_env(mk.make_env_with_recorders(TaichiEnv, trace))
    assert trace == ref, [(i, a, b) for i, (a, b) in enumerate(zip(trace, ref)) if a != b][:3]


def test_temporal_range_schedule_equals_the_reference():
    """ShapeMatchingLoss.expand_temporal_range (losses/shapematching_loss.py:110-130): plateau counting and horizon expansion driven by a
    loss sequence — the fixture was produced by the reference's own method."""
    import io
    import contextlib
    import make_reference_fixtures as mk
    from fluidlab_b200.losses import ShapeMatchingLoss
    L = ShapeMatchingLoss.__new__(ShapeMatchingLoss)
    L.temporal_range_type, L.temporal_range, L.best_loss, L.plateau_count, L.inf = 'expand', [0, 50], 1e8, 0, 1e8
    L.plateau_thresh, L.plateau_count_limit, L.temporal_expand_speed, L.max_loss_steps = [0.01, 0.5], 5, 50, 220
    got = []
    for v in mk.LOSS_SEQUENCE:
        L.total_loss = float(v)
        with contextlib.redirect_stdout(io.StringIO()):
            L.expand_temporal_range()
        got.append([L.temporal_range[1], L.plateau_count, float(L.best_loss)])
    assert np.array_equal(np.array(got, dtype=np.float64), D['temporal_range_schedule'])


def test_agent_effector_dispatch_equals_the_reference():
    """Agent (agents/agent.py:68-131): splitting of the action vector over the effectors, visiting order forward / reverse, get_grad
    concatenation, state get / set, frame and adjoint shuffles — trace of the REAL class with recorder effectors."""
    import json
    import make_reference_fixtures as mk
    from fluidlab_b200.agents import Agent
    ref = json.loads(str(D['agent_trace_json']))
    trace = []
    a = Agent(max_substeps_local=20, max_substeps_global=1000, max_action_steps_global=50, ckpt_dest='cpu')
    mk.drive_agent(a, trace)
    assert trace == ref, [(i, x, y) for i, (x, y) in enumerate(zip(trace, ref)) if x != y][:3]
