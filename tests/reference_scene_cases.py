"""The PRODUCT (fluidlab_b200: MPMSimulator + agents, real kernels) against runs of the reference's own agent scenes
(tests/golden/reference_run_{latteart,jetbot}.npz — the unmodified reference classes stepped on the Taichi emulation,
tests/golden/make_reference_run.py).  tests/test_reference_run.py pins the ORACLE to these fixtures; here the device path is compared with
them directly.  Shared by tests/test_cuda_emu_mpm.py (emulated device) and tests/test_gpu_parity.py (B200)."""
import os
import numpy as np

from conftest import make_particles

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
BARS = dict(x=1e-5, F=2e-5, v=2e-4, C=2e-3)


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-12))


def _check(fr, d):
    assert np.array_equal(fr['used'], d['ref_used']), 'used flags differ from the reference run'
    u = fr['used'] != 0
    for k, bar in BARS.items():
        assert rel(fr[k][u], d['ref_' + k][u]) < bar, (k, rel(fr[k][u], d['ref_' + k][u]))
    assert np.array_equal(fr['x'][~u], d['ref_x'][~u]), 'parked particles differ'


FUSE = [False]   # tests flip it to run the same scenes through the g2p2g path


def _sim(d, device, gravity, bnd):
    from fluidlab_b200 import MPMSimulator
    s = MPMSimulator(dim=3, quality=int(d['n_grid']) / 64, gravity=gravity, horizon=10, max_substeps_local=int(d['T']), max_substeps_global=1000,
                     ckpt_dest='gpu' if device is None else 'cpu', device=device)
    s.setup_boundary(**bnd)
    s.fuse_g2p2g = FUSE[0]
    return s


def _drive(s, agent, inj, d):
    inj.random_vector_np = np.asarray(d['random_vector'], dtype=np.float32)     # the reference drew it from the global NumPy RNG
    P = make_particles(d['x0'], d['mat'], int(d['n_grid']), used=d['used0'])
    s.build(agent, None, [], P)
    agent.build(s)
    assert np.abs(inj.get_state(0)[:7] - d['init_state'][:7]).max() < 1e-6, 'init_pos / init_euler -> pose (effector.py:63-73)'
    agent.apply_action_p(d['action_p'])
    assert s._can_fuse_injector() == FUSE[0]
    for a in d['actions']:
        s.step(a)
    f = s.cur_substep_local
    assert np.abs(inj.get_state(f)[:7] - d['ref_pose'][:7]).max() < 2e-6, (inj.get_state(f), d['ref_pose'])
    _check(s.readframe(f), d)


def run_latteart_case(device=None):
    """AgentInjector + locally-random Injector whose own boundary is a y-pinned cylinder (radial clamp hit), MILK injected into COFFEE,
    cylinder domain, gravity -20, 3 steps"""
    from fluidlab_b200 import AgentInjector
    d = np.load(os.path.join(G, 'reference_run_latteart.npz'))
    s = _sim(d, device, (0.0, -20.0, 0.0), dict(type='cylinder', xz_radius=0.2, xz_center=(0.5, 0.5), y_range=(0.34, 0.9)))
    common = dict(max_substeps_local=int(d['T']), max_substeps_global=1000, max_action_steps_global=20, ckpt_dest='cpu')
    agent = AgentInjector(**common)
    agent.add_effector(type='Injector', params=dict(radius=0.0075, flux=int(d['flux']), init_pos=(0.5, 0.5, 0.5), action_dim=3, inject_v=(0.0, -3.0, 0.0),
                                                    action_scale_p=(1.0, 1.0, 1.0), action_scale_v=(1.0, 1.0, 1.0), locally_random=True),
                       mesh_cfg=None, boundary_cfg=dict(type='cylinder', xz_radius=0.12, xz_center=(0.5, 0.5), y_range=(0.55, 0.55)))
    _drive(s, agent, agent.effectors[0], d)


def run_jetbot_case(device=None, randv=False):
    """AgentJetBot: a 6-DOF Injector (rotated pose, pose chain with quaternions) injecting WATER into a pool + the collector, 3 steps"""
    from fluidlab_b200 import AgentJetBot
    d = np.load(os.path.join(G, 'reference_run_jetbot_randv.npz' if randv else 'reference_run_jetbot.npz'))   # randv: Injector(randomize_inject_v=True), injector.py:96-97
    cube = lambda lo, hi: dict(type='cube', lower=tuple(lo), upper=tuple(hi))
    s = _sim(d, device, (0.0, -10.0, 0.0), cube(d['b_lower'], d['b_upper']))
    common = dict(max_substeps_local=int(d['T']), max_substeps_global=1000, max_action_steps_global=20, ckpt_dest='cpu')
    agent = AgentJetBot(collector_boundary=cube(d['c_lower'], d['c_upper']), **common)
    agent.add_effector(type='Injector', params=dict(radius=0.015, flux=int(d['flux']), init_pos=(0.58, 0.55, 0.5), init_euler=(20.0, 35.0, -10.0), inject_v=(-3.0, 0.0, 0.0),
                                                    inject_p=(-0.07, 0.0, 0.0), randomize_inject_v=randv, action_dim=6, action_scale_p=(1.0,) * 6, action_scale_v=(1.0, 1.0, 1.0, 5.0, 5.0, 5.0)),
                       mesh_cfg=None, boundary_cfg=cube(d['e_lower'], d['e_upper']))
    _drive(s, agent, agent.effectors[0], d)
    assert int(d['ref_used'].sum()) < int(d['used0'].sum()) + int(d['flux']) * 10 * int(d['n_steps']), 'the reference run collected nothing'


def run_jetbot_randv_case(device=None):
    """the AgentJetBot scene with Injector(randomize_inject_v=True): every injected particle's velocity gets (2 random_vector - 1) * |inject_v| * 2 on top of the
    rotated inject_v (effectors/injector.py:96-97); a run of the reference's own class"""
    run_jetbot_case(device, randv=True)


def run_pouring_case(device=None):
    """AgentPouring: a 6-DOF Rigid effector whose Dynamic SDF mesh (rotated, anisotropically scaled box) collides at grid AND particle level with
    friction + soft influence, plus the collector; elastic blob; 2 steps.  The mesh goes through the product's own transform code
    (meshes.py: T_mesh_to_voxels @ inv(T_init)) and must land on the reference's matrix."""
    from fluidlab_b200 import AgentPouring, macros as M
    d = np.load(os.path.join(G, 'reference_run_pouring.npz'))
    cube = lambda lo, hi: dict(type='cube', lower=tuple(lo), upper=tuple(hi))
    s = _sim(d, device, (0.0, -10.0, 0.0), cube(d['b_lower'], d['b_upper']))
    res, he = 32, 0.2                                   # the volume's mesh frame, as in tests/golden/make_reference_run.py: scene_pouring
    sc = (res - 1) / (2 * he)
    Tm = np.eye(4); Tm[0, 0] = Tm[1, 1] = Tm[2, 2] = sc; Tm[:3, 3] = sc * he
    common = dict(max_substeps_local=int(d['T']), max_substeps_global=1000, max_action_steps_global=20, ckpt_dest='cpu')
    agent = AgentPouring(collector_boundary=cube(d['c_lower'], d['c_upper']), **common)
    agent.add_effector(type='Rigid', params=dict(init_pos=(0.5, 0.64, 0.5), init_euler=(0.0, 23.0, 5.0), action_dim=6, action_scale_p=(1.0,) * 6, action_scale_v=(1.0,) * 6),
                       mesh_cfg=dict(file='box.obj', material=M.STIRRER, softness=100.0, scale=(1.0, 0.9, 1.1), euler=(0.0, 10.0, 0.0),
                                     sdf=dict(voxels=d['vox'], T_mesh_to_voxels=Tm)), boundary_cfg=cube(d['e_lower'], d['e_upper']))
    rigid = agent.effectors[0]
    assert np.abs(rigid.mesh.T_mesh_to_voxels_np - d['T_final']).max() < 1e-4 * np.abs(d['T_final']).max()
    assert abs(rigid.mesh.friction - float(d['friction'])) < 1e-7
    P = make_particles(d['x0'], d['mat'], int(d['n_grid']))
    s.build(agent, None, [], P)
    agent.build(s)
    agent.apply_action_p(d['action_p'])
    assert s._can_fuse_injector() == FUSE[0]
    for a in d['actions']:
        s.step(a)
    f = s.cur_substep_local
    assert np.abs(rigid.get_state(f)[:7] - d['ref_pose'][:7]).max() < 2e-6
    assert np.abs(d['ref_v']).max() > 5.0 and int(d['ref_used'].sum()) < len(d['x0']), 'the reference collider / collector did nothing'
    _check(s.readframe(f), d)


def run_icecream_case(device=None):
    """AgentIceCreamDynamic: BallInjector of plasto-elastic ICECREAM that stops at inject_till + a Rigid sphere collider gated at y > 0.25 + a
    Static mesh with dynamics colliding in grid_op; every particle starts parked; 3 steps"""
    from conftest import sphere_sdf, box_sdf
    from fluidlab_b200 import AgentIceCreamDynamic, Statics, macros as M
    d = np.load(os.path.join(G, 'reference_run_icecream.npz'))
    cube = dict(type='cube', lower=tuple(d['lower']), upper=tuple(d['upper']))
    s = _sim(d, device, (0.0, -10.0, 0.0), cube)
    vox, Tm = sphere_sdf(0.10, 0.2)
    bv, bT = box_sdf((0.3, 0.05, 0.3), 0.4)
    assert np.array_equal(vox.reshape(d['vox'].shape), d['vox']) and np.array_equal(bv.reshape(d['svox'].shape), d['svox'])
    common = dict(max_substeps_local=int(d['T']), max_substeps_global=1000, max_action_steps_global=20, ckpt_dest='cpu')
    agent = AgentIceCreamDynamic(inject_till=int(d['inject_till']), **common)
    agent.add_effector(type='BallInjector', params=dict(locally_random=True, radius=0.035, flux=int(d['flux']), init_pos=(0.5, 0.62, 0.5), inject_v=(0.0, -0.4, 0.0), action_dim=3),
                       mesh_cfg=None, boundary_cfg=cube)
    agent.add_effector(type='Rigid', params=dict(init_pos=(0.5, 0.46, 0.5), action_dim=3),
                       mesh_cfg=dict(file='cone.obj', material=M.CONE, softness=100.0, sdf=dict(voxels=vox.reshape(d['vox'].shape), T_mesh_to_voxels=Tm)), boundary_cfg=cube)
    statics = Statics()
    statics.add_static(file='plate.obj', material=M.CUP, has_dynamics=True, pos=(0.5, 0.36, 0.5), sdf=dict(voxels=bv.reshape(d['svox'].shape), T_mesh_to_voxels=bT))
    inj, rigid = agent.effectors
    assert np.abs(rigid.mesh.T_mesh_to_voxels_np - d['T_rigid']).max() < 1e-5 * np.abs(d['T_rigid']).max()
    assert np.abs(statics[0].T_mesh_to_voxels_np - d['T_static']).max() < 1e-5 * np.abs(d['T_static']).max()
    inj.random_vector_np = np.asarray(d['random_vector'], dtype=np.float32)
    N = len(d['x0'])
    P = make_particles(d['x0'], d['mat'], int(d['n_grid']), used=np.zeros(N, np.int32))
    s.build(agent, None, statics, P)
    agent.build(s)
    agent.apply_action_p(d['action_p'])
    assert s._can_fuse_injector() == FUSE[0]
    for a in d['actions']:
        s.step(a)
    f = s.cur_substep_local
    assert np.abs(rigid.get_state(f)[:7] - d['ref_pose'][:7]).max() < 2e-6
    assert int(d['ref_used'].sum()) == int(d['flux']) * (int(d['inject_till']) + 1) or int(d['ref_used'].sum()) > 0
    _check(s.readframe(f), d)


def run_cloud_adjoint_case(device=None):
    """the device ADJOINT (fp32) against central finite differences taken through the reference's own forward kernels in float64
    (tests/golden/reference_fd.npz, make_reference_fd.py): L = sum w . state_3 of a WATER / ELASTIC / ICECREAM / MILK_VIS cloud, 20 picked
    entries of dL/d(x, v, C, F)_0.  fp32 against an fp64 finite difference: 2e-3 of the largest picked derivative."""
    from fluidlab_b200 import MPMSimulator
    FD = np.load(os.path.join(G, 'reference_fd.npz'), allow_pickle=True)
    n_grid, n_sub = int(FD['cloud_n_grid']), int(FD['cloud_n_sub'])
    P = make_particles(FD['cloud_x'], FD['cloud_mat'], n_grid)
    s = MPMSimulator(dim=3, quality=n_grid / 64, gravity=(0.0, -10.0, 0.0), horizon=10, max_substeps_local=10, max_substeps_global=1000,
                     ckpt_dest='gpu' if device is None else 'cpu', device=device)
    s.setup_boundary(type='cube', lower=tuple(FD['cloud_lower']), upper=tuple(FD['cloud_upper']))
    s.build(None, None, [], P)
    s.setframe(0, FD['cloud_x'], FD['cloud_v'], FD['cloud_C'], FD['cloud_F'], np.ones(len(FD['cloud_x']), np.int32))
    s.enable_grad()
    for f in range(n_sub):
        s.substep(f, True); s.cur_substep_global += 1
    fr = s.readframe(n_sub)
    w = {k: FD['cloud_w_' + k] for k in ('x', 'v', 'C', 'F')}
    loss = sum((w[k] * fr[k].astype(np.float64)).sum() for k in w)
    assert abs(loss - float(FD['cloud_loss'])) < 1e-4 * max(1.0, abs(loss)), (loss, float(FD['cloud_loss']))   # fp32 state vs fp64 reference; depends on the order of the float atomics
    s.reset_grad(); s.set_grad(w['x'], w['v'], w['C'], w['F'])
    for f in reversed(range(n_sub)):
        s.cur_substep_global -= 1; s.substep_grad(f, True)
    g = s.get_grad()
    scale = float(np.abs(FD['cloud_fd']).max())
    for key, idx, fd in zip(FD['cloud_pick_key'], FD['cloud_pick_idx'], FD['cloud_fd']):
        an = float(g[str(key)].reshape(-1)[int(idx)])
        assert abs(an - fd) <= 2e-3 * scale, (str(key), int(idx), an, float(fd), scale)


def run_latteart_fused_case(device=None):
    """the LatteArt reference run again, stepped through the fused path (MPMSimulator.fuse_g2p2g with an AgentInjector: g2p2g kernels + the
    separate scatter of the freshly injected particles)"""
    from fluidlab_b200 import AgentInjector
    d = np.load(os.path.join(G, 'reference_run_latteart.npz'))
    FUSE[0] = True
    try:
        return _latteart(d, device)
    finally:
        FUSE[0] = False


def _latteart(d, device):
    from fluidlab_b200 import AgentInjector
    s = _sim(d, device, (0.0, -20.0, 0.0), dict(type='cylinder', xz_radius=0.2, xz_center=(0.5, 0.5), y_range=(0.34, 0.9)))
    common = dict(max_substeps_local=int(d['T']), max_substeps_global=1000, max_action_steps_global=20, ckpt_dest='cpu')
    agent = AgentInjector(**common)
    agent.add_effector(type='Injector', params=dict(radius=0.0075, flux=int(d['flux']), init_pos=(0.5, 0.5, 0.5), action_dim=3, inject_v=(0.0, -3.0, 0.0),
                                                    action_scale_p=(1.0, 1.0, 1.0), action_scale_v=(1.0, 1.0, 1.0), locally_random=True),
                       mesh_cfg=None, boundary_cfg=dict(type='cylinder', xz_radius=0.12, xz_center=(0.5, 0.5), y_range=(0.55, 0.55)))
    _drive(s, agent, agent.effectors[0], d)
