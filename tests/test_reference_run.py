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
def test_agent_jetbot_6dof_injector_and_collector(prec):
    d = np.load(os.path.join(G, 'reference_run_jetbot.npz'))
    N = len(d['x0'])
    P = make_particles(d['x0'], d['mat'], int(d['n_grid']), used=d['used0'])
    o = orc.OracleSim(int(d['n_grid']), P, gravity=(0, -10, 0), boundary=cube(d['b_lower'], d['b_upper']), precision=prec, max_substeps_local=int(d['T']))
    o.add_effector(type=1, action_dim=6, scale_v=(1, 1, 1, 5, 5, 5), boundary=cube(d['e_lower'], d['e_upper']), radius=0.015, flux=int(d['flux']), inject_v=(-3.0, 0, 0),
                   inject_p=(-0.07, 0, 0), locally_random=False, random_vector=d['random_vector'], act_range=np.where(d['used0'] == 0)[0], max_action_steps=20)
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
