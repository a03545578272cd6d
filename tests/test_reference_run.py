"""The CPU oracle against FORWARD RUNS OF THE REAL REFERENCE KERNELS.

`tests/golden/reference_run_*.npz` were produced in the build container by executing the unmodified source of zhouxian/FluidLab's
MPMSimulator / boundaries / effectors / agents / SDF meshes on a NumPy emulation of the Taichi API (tests/golden/taichi_emu.py,
tests/golden/make_reference_run.py: what is emulated — `ti.svd`, fp32 summation order, value-semantics of `a = b` — is listed there).
The oracle must reproduce the resulting particle states; this is the pin of its forward restatement to the reference itself.
Bars: fp32-vs-fp32 implementations of the same arithmetic after 10-30 substeps: x 1e-5, F 1e-5, v 1e-4, C 5e-4 (relative, max-norm)."""
import os
import numpy as np
import pytest

from conftest import make_particles
from oracle import oracle as orc
from fluidlab_b200 import macros as M

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
BARS = dict(x=1e-5, F=1e-5, v=1e-4, C=5e-4)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


def cube(lo, hi):
    return dict(type='cube', lower=tuple(lo), upper=tuple(hi))


def check(fr, d, sel=None):
    assert np.array_equal(fr['used'], d['ref_used']), 'used flags differ from the reference run'
    sel = (fr['used'] != 0) if sel is None else sel
    for k, bar in BARS.items():
        assert rel(fr[k][sel], d['ref_' + k][sel]) < bar, (k, rel(fr[k][sel], d['ref_' + k][sel]))
    off = fr['used'] == 0
    assert np.array_equal(fr['x'][off].astype(np.float32), d['ref_x'][off]), 'parked particles differ'


@pytest.mark.parametrize('prec', [32, 64])
def test_all_material_classes_and_cube_walls(prec):
    d = np.load(os.path.join(G, 'reference_run_multimat.npz'))
    P = make_particles(d['x0'], d['mat'], int(d['n_grid']), used=d['used0'])
    o = orc.OracleSim(int(d['n_grid']), P, gravity=tuple(d['gravity']), boundary=cube(d['b_lower'], d['b_upper']), precision=prec, max_substeps_local=20)
    o.set_frame(0, d['x0'], d['v0'], d['C0'], d['F0'], d['used0'])
    for f in range(int(d['n_sub'])):
        o.substep(f)
    fr = o.get_frame(int(d['n_sub']))
    assert len(set(int(m) for m in d['mat'])) == 5
    check(fr, d)
    off = d['used0'] == 0     # process_unused_particles: unused slots carry their whole state forward (MPM:309-316)
    for k in ('v', 'C', 'F'):
        assert np.array_equal(fr[k][off].astype(np.float32), d['ref_' + k][off]), k


@pytest.mark.parametrize('prec', [32, 64])
def test_mat_rigid_bodies_and_cylinder_boundary(prec):
    d = np.load(os.path.join(G, 'reference_run_rigid_bodies.npz'))
    P = make_particles(d['x0'], d['mat'], int(d['n_grid']), used=d['used0'])
    o = orc.OracleSim(int(d['n_grid']), P, gravity=tuple(d['gravity']), precision=prec, max_substeps_local=20,
                      boundary=dict(type='cylinder', xz_radius=float(d['xz_radius']), xz_center=tuple(d['xz_center']), y_range=tuple(d['y_range'])))
    o.set_bodies(d['body_id'], 4)
    o.set_frame(0, d['x0'], d['v0'], d['C0'], d['F0'], d['used0'])
    for f in range(int(d['n_sub'])):
        o.substep(f)
    check(o.get_frame(int(d['n_sub'])), d)


@pytest.mark.parametrize('prec', [32, 64])
def test_boundary_restitution_and_lock_dims(prec):
    d = np.load(os.path.join(G, 'reference_run_locked.npz'))
    N = len(d['x0'])
    P = make_particles(d['x0'], d['mat'], int(d['n_grid']))
    bnd = dict(type='cube', lower=tuple(d['b_lower']), upper=tuple(d['b_upper']), restitution=float(d['restitution']), lock_dims=[int(v) for v in d['lock_dims']])
    o = orc.OracleSim(int(d['n_grid']), P, gravity=(0, -10, 0), boundary=bnd, precision=prec, max_substeps_local=20)
    o.set_bodies(d['body_id'], 2)
    o.set_frame(0, d['x0'], d['v0'], d['C0'], d['F0'], np.ones(N, np.int32))
    for f in range(int(d['n_sub'])):
        o.substep(f)
    fr = o.get_frame(int(d['n_sub']))
    assert np.all(d['ref_v'][:, 2] == 0.0), 'lock_dims=[2] must zero the z velocity of every particle'
    assert d['ref_x'][:, 0].max() > 0.62 and d['ref_x'][:, 1].min() < 0.37, 'the scene must reach the walls'
    check(fr, d)


@pytest.mark.parametrize('randv', [False, True], ids=['plain', 'randomize_inject_v'])
@pytest.mark.parametrize('prec', [32, 64])
def test_agent_jetbot_6dof_injector_and_collector(prec, randv):
    d = np.load(os.path.join(G, 'reference_run_jetbot_randv.npz' if randv else 'reference_run_jetbot.npz'))   # randv: Injector(randomize_inject_v=True), injector.py:96-97
    N = len(d['x0'])
    P = make_particles(d['x0'], d['mat'], int(d['n_grid']), used=d['used0'])
    o = orc.OracleSim(int(d['n_grid']), P, gravity=(0, -10, 0), boundary=cube(d['b_lower'], d['b_upper']), precision=prec, max_substeps_local=int(d['T']))
    o.add_effector(type=1, action_dim=6, scale_v=(1, 1, 1, 5, 5, 5), boundary=cube(d['e_lower'], d['e_upper']), radius=0.015, flux=int(d['flux']), inject_v=(-3.0, 0, 0),
                   inject_p=(-0.07, 0, 0), locally_random=False, random_vector=d['random_vector'], act_range=np.where(d['used0'] == 0)[0], max_action_steps=20, randomize_inject_v=randv)
    o.set_collector(cube(d['c_lower'], d['c_upper']), mat=M.WATER)
    o.set_frame(0, d['x0'], np.zeros((N, 3)), np.zeros((N, 3, 3)), np.tile(np.eye(3), (N, 1, 1)), d['used0'])
    o.set_effector_state(0, 0, np.concatenate([d['init_state'][:7], [0.0]])); o.apply_action_p(d['action_p'])
    total = 0.0
    for i in range(int(d['n_steps'])):
        o.step(d['actions'][i])
        total += o.loss_value(o.cur_substep_local, M.WATER, 1.0, d['tgt'][i])   # vs the reference's own ShapeMatchingLoss kernels
    assert abs(total - float(d['ref_loss'])) < 1e-5 * float(d['ref_loss']), (total, float(d['ref_loss']))
    fr = o.get_frame(o.cur_substep_local)
    n_used = int(d['ref_used'].sum())
    assert n_used < int(d['used0'].sum()) + int(d['flux']) * 10 * int(d['n_steps']), 'the reference run collected nothing'
    check(fr, d)
    assert np.abs(o.effector_state(0, o.cur_substep_local)[:7] - d['ref_pose'][:7]).max() < 2e-6     # pose chain incl. the quaternion


@pytest.mark.parametrize('prec', [32, 64])
def test_latteart_like_injector_with_cylinder_boundaries(prec):
    d = np.load(os.path.join(G, 'reference_run_latteart.npz'))
    N = len(d['x0'])
    P = make_particles(d['x0'], d['mat'], int(d['n_grid']), used=d['used0'])
    bnd = dict(type='cylinder', xz_radius=0.2, xz_center=(0.5, 0.5), y_range=(0.34, 0.9))
    ebnd = dict(type='cylinder', xz_radius=0.12, xz_center=(0.5, 0.5), y_range=(0.55, 0.55))
    o = orc.OracleSim(int(d['n_grid']), P, gravity=(0, -20, 0), boundary=bnd, precision=prec, max_substeps_local=int(d['T']))
    o.add_effector(type=1, action_dim=3, boundary=ebnd, radius=0.0075, flux=int(d['flux']), inject_v=(0, -3, 0), inject_p=(0, 0, 0), locally_random=True,
                   random_vector=d['random_vector'], act_range=np.where(d['used0'] == 0)[0], max_action_steps=20)
    o.set_frame(0, d['x0'], np.zeros((N, 3)), np.zeros((N, 3, 3)), np.tile(np.eye(3), (N, 1, 1)), d['used0'])
    o.set_effector_state(0, 0, np.concatenate([d['init_state'][:7], [0.0]])); o.apply_action_p(d['action_p'])
    for i in range(int(d['n_steps'])):
        o.step(d['actions'][i])
    pose = o.effector_state(0, o.cur_substep_local)
    assert abs(pose[1] - 0.55) < 1e-6 and abs(np.hypot(pose[0] - 0.5, pose[2] - 0.5) - 0.12) < 1e-5, 'the scene must exercise the pinned y and the radial clamp'
    assert np.abs(pose[:7] - d['ref_pose'][:7]).max() < 2e-6
    check(o.get_frame(o.cur_substep_local), d)


@pytest.mark.parametrize('prec', [32, 64])
def test_agent_pouring_6dof_rigid_collider_both_levels(prec):
    d = np.load(os.path.join(G, 'reference_run_pouring.npz'))
    N = len(d['x0'])
    P = make_particles(d['x0'], d['mat'], int(d['n_grid']))
    o = orc.OracleSim(int(d['n_grid']), P, gravity=(0, -10, 0), boundary=cube(d['b_lower'], d['b_upper']), precision=prec, max_substeps_local=int(d['T']))
    o.add_effector(type=0, action_dim=6, scale_v=(1,) * 6, boundary=cube(d['e_lower'], d['e_upper']), max_action_steps=20)
    o.set_rigid_mesh(d['vox'], d['T_final'], friction=float(d['friction']), softness=float(d['softness']), collide_type='both')
    o.set_collector(cube(d['c_lower'], d['c_upper']), mat=-1)
    o.set_frame(0, d['x0'], np.zeros((N, 3)), np.zeros((N, 3, 3)), np.tile(np.eye(3), (N, 1, 1)), np.ones(N, np.int32))
    o.set_effector_state(0, 0, np.concatenate([d['init_state'][:7], [0.0]])); o.apply_action_p(d['action_p'])
    for i in range(int(d['n_steps'])):
        o.step(d['actions'][i])
    fr = o.get_frame(o.cur_substep_local)
    assert np.abs(d['ref_v']).max() > 5.0 and int(d['ref_used'].sum()) < N, 'the reference collider / collector did nothing'
    check(fr, d)
    assert np.abs(o.effector_state(0, o.cur_substep_local)[:7] - d['ref_pose'][:7]).max() < 2e-6


@pytest.mark.parametrize('prec', [32, 64])
def test_agent_icecream_dynamic_ball_injector_static_and_rigid_colliders(prec):
    d = np.load(os.path.join(G, 'reference_run_icecream.npz'))
    N = len(d['x0'])
    P = make_particles(d['x0'], d['mat'], int(d['n_grid']), used=np.zeros(N, np.int32))
    bnd = cube(d['lower'], d['upper'])
    o = orc.OracleSim(int(d['n_grid']), P, gravity=(0, -10, 0), boundary=bnd, precision=prec, max_substeps_local=int(d['T']))
    o.add_effector(type=2, action_dim=3, boundary=bnd, radius=0.035, flux=int(d['flux']), inject_v=(0, -0.4, 0), locally_random=True, random_vector=d['random_vector'],
                   act_range=np.arange(N), max_action_steps=20, init_pos=(0.5, 0.62, 0.5))
    o.add_effector(type=0, action_dim=3, boundary=bnd, max_action_steps=20, init_pos=(0.5, 0.46, 0.5))
    o.set_rigid_mesh(d['vox'], d['T_rigid'], friction=float(d['friction_rigid']), softness=100.0, collide_type='particle')
    o.add_static(d['svox'], d['T_static'], friction=float(d['friction_static']))
    o.set_icecream_agent(int(d['inject_till']))
    o.set_frame(0, d['x0'], np.zeros((N, 3)), np.zeros((N, 3, 3)), np.tile(np.eye(3), (N, 1, 1)), np.zeros(N, np.int32))
    o.set_effector_state(0, 0, np.array([0.5, 0.62, 0.5, 1, 0, 0, 0, 0.0])); o.set_effector_state(1, 0, np.array([0.5, 0.46, 0.5, 1, 0, 0, 0, 0.0]))
    o.apply_action_p(d['action_p'])
    for i in range(int(d['n_steps'])):
        o.step(d['actions'][i])
    fr = o.get_frame(o.cur_substep_local)
    assert int(d['ref_used'].sum()) == int(d['flux']) * int(d['inject_till'])     # injection stopped at inject_till
    check(fr, d)
    assert np.abs(o.effector_state(1, o.cur_substep_local)[:7] - d['ref_pose'][:7]).max() < 2e-6


# ------------------------------------------------------------------------------------------------------------------------------------
# adjoints vs central differences THROUGH THE REFERENCE'S OWN FORWARD CODE in fp64 (tests/golden/make_reference_fd.py)
# ------------------------------------------------------------------------------------------------------------------------------------
FD = np.load(os.path.join(G, 'reference_fd.npz'))


def test_substep_adjoint_equals_finite_differences_of_the_reference_forward():
    """3 substeps of a WATER / ELASTIC / ICECREAM / MILK_VIS cloud: the oracle's fp64 adjoint of L = sum w . state_3 w.r.t. 20 entries of
    (x, v, C, F)_0 against central differences of the REFERENCE's forward kernels run in float64 (its own `dprecision = 64` switch)."""
    n_grid, n_sub = int(FD['cloud_n_grid']), int(FD['cloud_n_sub'])
    P = make_particles(FD['cloud_x'], FD['cloud_mat'], n_grid)
    # the fp64 reference run holds the material constants and the particle mass in double (make_particles rounds them to f32 like the f32 build)
    P['mu'] = np.array([M.MU[int(m)] for m in FD['cloud_mat']], dtype=np.float64); P['lam'] = np.array([M.LAMDA[int(m)] for m in FD['cloud_mat']], dtype=np.float64)
    P['mass'] = (0.5 / n_grid) ** 2 * np.array([M.RHO[int(m)] for m in FD['cloud_mat']], dtype=np.float64)
    o = orc.OracleSim(n_grid, P, gravity=(0, -10, 0), boundary=cube(FD['cloud_lower'], FD['cloud_upper']), precision=64, max_substeps_local=10)
    o.set_frame(0, FD['cloud_x'], FD['cloud_v'], FD['cloud_C'], FD['cloud_F'], P['used'])
    for f in range(n_sub):
        o.substep(f)
    fr = o.get_frame(n_sub)
    w = {k: FD['cloud_w_' + k] for k in ('x', 'v', 'C', 'F')}
    loss = sum((w[k] * fr[k]).sum() for k in w)
    assert abs(loss - float(FD['cloud_loss'])) < 1e-9 * max(1.0, abs(loss)), (loss, float(FD['cloud_loss']))   # forward, fp64 vs fp64
    o.reset_grad(); o.set_grad_frame(n_sub, w['x'], w['v'], w['C'], w['F'])
    for f in reversed(range(n_sub)):
        o.substep_grad(f)
    g = o.get_grad_frame(0)
    for key, idx, fd in zip(FD['cloud_pick_key'], FD['cloud_pick_idx'], FD['cloud_fd']):
        an = g[str(key)].reshape(-1)[int(idx)]
        assert abs(an - fd) <= 2e-6 * max(1.0, abs(fd)), (str(key), int(idx), an, fd)


def test_dloss_daction_equals_finite_differences_of_the_reference_forward():
    """6-DOF Injector (AgentInjector, pose / quaternion chain, rotated inject_p / inject_v) + the reference's own ShapeMatchingLoss: the
    oracle's dLoss/dAction (velocity, rotation and initial-position actions) against central differences of the reference forward in fp64."""
    d = FD
    N = len(d['inj_x'])
    n_steps = int(d['inj_n_steps'])
    P = make_particles(d['inj_x'], d['inj_mat'], int(d['inj_n_grid']), used=d['inj_used'])
    o = orc.OracleSim(int(d['inj_n_grid']), P, gravity=(0, -10, 0), boundary=cube(d['inj_lower'], d['inj_upper']), precision=64, max_substeps_local=int(d['inj_T']))
    o.add_effector(type=1, action_dim=6, scale_v=(1, 1, 1, 5, 5, 5), boundary=cube(d['inj_e_lower'], d['inj_e_upper']), radius=0.015, flux=int(d['inj_flux']),
                   inject_v=(-3.0, 0.5, 0.0), inject_p=(-0.07, 0.01, 0.0), locally_random=True, random_vector=d['inj_random_vector'],
                   act_range=np.where(d['inj_used'] == 0)[0], max_action_steps=20)
    o.enable_grad()
    o.set_frame(0, d['inj_x'], np.zeros((N, 3)), np.zeros((N, 3, 3)), np.tile(np.eye(3), (N, 1, 1)), d['inj_used'])
    o.set_effector_state(0, 0, np.concatenate([d['inj_init_state'][:7], [0.0]])); o.apply_action_p(d['inj_action_p'])
    total = 0.0
    for i in range(n_steps):
        o.step(d['inj_actions'][i]); total += o.loss_value(o.cur_substep_local, M.MILK, 1.0, d['inj_tgt'][i])
    assert abs(total - float(d['inj_loss'])) < 1e-9 * abs(total), (total, float(d['inj_loss']))
    o.reset_grad()
    for i in range(n_steps - 1, -1, -1):
        o.loss_seed(o.cur_substep_local, M.MILK, 1.0, d['inj_tgt'][i]); o.step_grad(d['inj_actions'][i])
    o.apply_action_p_grad()
    g = o.get_action_grad(n_steps)
    for (i, j), fd in zip(d['inj_picks'], d['inj_fd']):
        assert abs(g[int(i), int(j)] - fd) <= 2e-6 * max(1.0, abs(fd)), (int(i), int(j), g[int(i), int(j)], fd)


def test_mat_rigid_adjoint_equals_finite_differences_of_the_reference_forward():
    """advect_grad of MAT_RIGID bodies (shape matching + manual SVD adjoint, MPM:436-447, 485-489): oracle fp64 adjoint at 16 entries of rigid
    particles' (x, v, C, F)_0 against central differences of the reference's forward kernels (compute_COM / compute_H / ti.svd-emulated /
    compute_R / advect_kernel) in float64."""
    d = FD
    n_grid, n_sub = int(d['rb_n_grid']), int(d['rb_n_sub'])
    P = make_particles(d['rb_x'], d['rb_mat'], n_grid)
    P['mu'] = np.array([M.MU[int(m)] for m in d['rb_mat']], dtype=np.float64); P['lam'] = np.array([M.LAMDA[int(m)] for m in d['rb_mat']], dtype=np.float64)
    P['mass'] = (0.5 / n_grid) ** 2 * np.array([M.RHO[int(m)] for m in d['rb_mat']], dtype=np.float64)
    o = orc.OracleSim(n_grid, P, gravity=(0, -10, 0), boundary=cube(d['rb_lower'], d['rb_upper']), precision=64, max_substeps_local=10)
    o.set_bodies(d['rb_bid'], 3)
    o.set_frame(0, d['rb_x'], d['rb_v'], d['rb_C'], d['rb_F'], P['used'])
    for f in range(n_sub):
        o.substep(f)
    fr = o.get_frame(n_sub)
    w = {k: d['rb_w_' + k] for k in ('x', 'v', 'C', 'F')}
    loss = sum((w[k] * fr[k]).sum() for k in w)
    assert abs(loss - float(d['rb_loss'])) < 1e-9 * max(1.0, abs(loss)), (loss, float(d['rb_loss']))
    o.reset_grad(); o.set_grad_frame(n_sub, w['x'], w['v'], w['C'], w['F'])
    for f in reversed(range(n_sub)):
        o.substep_grad(f)
    g = o.get_grad_frame(0)
    for key, idx, fd in zip(d['rb_pick_key'], d['rb_pick_idx'], d['rb_fd']):
        an = g[str(key)].reshape(-1)[int(idx)]
        assert abs(an - fd) <= 5e-6 * max(1.0, abs(fd)), (str(key), int(idx), an, fd)


def test_dloss_daction_6dof_collider_equals_finite_differences_of_the_reference_forward():
    """6-DOF Rigid (rotated, scaled box mesh) colliding at grid AND particle level through the reference's own Dynamic.collide: the oracle's
    dLoss/dAction against central differences of the reference forward in float64.  The contact map is piecewise smooth (hit / influence
    thresholds), so three step sizes were recorded and the closest one must agree."""
    d = FD
    N, n_steps = len(d['po_x']), int(d['po_n_steps'])
    P = make_particles(d['po_x'], d['po_mat'], int(d['po_n_grid']))
    P['mu'] = np.array([M.MU[int(m)] for m in d['po_mat']], dtype=np.float64); P['lam'] = np.array([M.LAMDA[int(m)] for m in d['po_mat']], dtype=np.float64)
    P['mass'] = (0.5 / int(d['po_n_grid'])) ** 2 * np.ones(N)
    o = orc.OracleSim(int(d['po_n_grid']), P, gravity=(0, -10, 0), boundary=cube(d['po_lower'], d['po_upper']), precision=64, max_substeps_local=int(d['po_T']))
    o.add_effector(type=0, action_dim=6, scale_v=(1,) * 6, boundary=cube(d['po_e_lower'], d['po_e_upper']), max_action_steps=20)
    o.set_rigid_mesh(d['po_vox'], d['po_T_final'], friction=float(d['po_friction']), softness=100.0, collide_type='both')
    o.enable_grad()
    o.set_frame(0, d['po_x'], d['po_v0'], d['po_C0'], d['po_F0'], np.ones(N, np.int32))
    o.set_effector_state(0, 0, np.concatenate([d['po_init_state'][:7], [0.0]])); o.apply_action_p(d['po_action_p'])
    for i in range(n_steps):
        o.step(d['po_actions'][i])
    fr = o.get_frame(o.cur_substep_local)
    loss = float((d['po_w'] * fr['x']).sum())
    assert abs(loss - float(d['po_loss'])) < 1e-7 * max(1.0, abs(loss)), (loss, float(d['po_loss']))
    o.reset_grad()
    o.set_grad_frame(o.cur_substep_local, d['po_w'], np.zeros((N, 3)), np.zeros((N, 3, 3)), np.zeros((N, 3, 3)))
    for i in range(n_steps - 1, -1, -1):
        o.step_grad(d['po_actions'][i])
    o.apply_action_p_grad()
    g = o.get_action_grad(n_steps)
    scale = max(1.0, np.abs(d['po_fd']).max())
    for q, (i, j) in enumerate(d['po_picks']):
        err = min(abs(g[int(i), int(j)] - fd) for fd in d['po_fd'][:, q])
        assert err <= 1e-5 * scale, (int(i), int(j), g[int(i), int(j)], d['po_fd'][:, q])


def test_manual_svd_adjoint_equals_the_reference_function():
    """MPMSimulator.backward_svd + clamp (MPM:272-302), the reference's own function evaluated on 24 random inputs (every third with nearly
    equal singular values, where the 1e-8 clamp acts): the oracle's literal restatement returns the same matrix."""
    import ctypes as C
    L = orc.lib()
    L.orc_backward_svd.argtypes = [C.c_void_p] * 7
    for row in FD['svd_grad_cases']:
        gU, gS, gV, U = (np.ascontiguousarray(row[i * 9:(i + 1) * 9]) for i in range(4))
        sig, V, ref = np.ascontiguousarray(row[36:39]), np.ascontiguousarray(row[39:48]), row[48:57]
        out = np.zeros(9)
        L.orc_backward_svd(gU.ctypes.data, gS.ctypes.data, gV.ctypes.data, U.ctypes.data, sig.ctypes.data, V.ctypes.data, out.ctypes.data)
        assert np.abs(out - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), (out, ref)
