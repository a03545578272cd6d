"""CPU tests of the oracle itself: SVD convention, conservation invariants, adjoint vs central
finite differences in fp64 (SURVEY.md §8c substitute pins 1-3).  No GPU needed."""
import numpy as np
import pytest
from conftest import make_particles, sphere_sdf, box_sdf
from oracle import oracle as orc
from fluidlab_b200 import macros as M


def test_svd_convention_and_reconstruction():
    rng = np.random.RandomState(0)
    for prec, tol in ((64, 1e-12), (32, 2e-6)):
        for it in range(200):
            A = np.eye(3) + 0.3 * rng.randn(3, 3) if it % 2 else rng.randn(3, 3)
            if it == 7:
                A = np.eye(3) * 1.3  # fully degenerate
            if it == 9:
                A = np.diag([2.0, 2.0, 0.5])
            if prec == 32:
                A = A.astype(np.float32).astype(np.float64)
            U, s, V = orc.svd3(A, prec)
            assert np.allclose(U @ np.diag(s) @ V.T, A, atol=tol * max(1, np.abs(A).max()))
            assert abs(np.linalg.det(U) - 1) < 10 * tol and abs(np.linalg.det(V) - 1) < 10 * tol
            assert abs(s[0]) >= abs(s[1]) - tol and abs(s[1]) >= abs(s[2]) - tol
            assert s[0] >= 0 and s[1] >= 0
            assert np.sign(np.linalg.det(A)) == np.sign(s[2]) or abs(s[2]) < tol
            ref = np.linalg.svd(A, compute_uv=False)
            assert np.allclose(np.abs(s), ref, atol=tol * 10)


def _cloud(n, rng, lo=0.35, hi=0.65):
    return rng.uniform(lo, hi, size=(n, 3))


def test_mass_momentum_conservation_p2g_g2p():
    rng = np.random.RandomState(1)
    n_grid = 16
    P = make_particles(_cloud(300, rng), M.WATER, n_grid)
    sim = orc.OracleSim(n_grid, P, gravity=(0, 0, 0), precision=64)
    fr = sim.get_frame(0)
    v0 = rng.randn(300, 3) * 0.1
    sim.set_frame(0, fr['x'], v0, fr['C'], fr['F'], fr['used'])
    L = sim.L
    L.orc_phase_reset_grid(sim.h); L.orc_phase_p2g(sim.h, 0, 1)
    vin, m, _ = sim.get_grid()
    assert np.isclose(m.sum(), P['mass'].sum(), rtol=1e-12)
    # F = I, C = 0 -> J = 1 -> zero stress -> grid momentum equals particle momentum
    assert np.allclose(vin.sum(0), (P['mass'][:, None] * v0).sum(0), rtol=1e-10, atol=1e-14)
    L.orc_phase_grid_op(sim.h, 0); L.orc_phase_g2p(sim.h, 0)
    f1 = sim.get_frame(1)
    assert np.allclose((P['mass'][:, None] * f1['v']).sum(0), (P['mass'][:, None] * v0).sum(0), rtol=1e-10, atol=1e-14)


def test_apic_preserves_rigid_translation():
    rng = np.random.RandomState(2)
    n_grid = 16
    P = make_particles(_cloud(400, rng), M.ELASTIC, n_grid)
    sim = orc.OracleSim(n_grid, P, gravity=(0, 0, 0), precision=64)
    fr = sim.get_frame(0)
    v0 = np.tile(np.array([0.3, -0.2, 0.1]), (400, 1))
    sim.set_frame(0, fr['x'], v0, fr['C'], fr['F'], fr['used'])
    sim.substep(0)
    f1 = sim.get_frame(1)
    assert np.allclose(f1['v'], v0, atol=1e-12)
    assert np.allclose(f1['C'], 0, atol=1e-9)
    assert np.allclose(f1['x'], fr['x'] + 2e-4 * v0, atol=1e-14)


def _random_state(sim, rng, amp_F=0.02, amp_C=5.0, amp_v=0.5):
    N = sim.N
    fr = sim.get_frame(0)
    v = rng.randn(N, 3) * amp_v
    Cm = rng.randn(N, 3, 3) * amp_C
    F = np.eye(3)[None] + rng.randn(N, 3, 3) * amp_F
    sim.set_frame(0, fr['x'], v, Cm, F, fr['used'])
    return dict(x=fr['x'].copy(), v=v, C=Cm, F=F, used=fr['used'])


def _run_loss(sim, st, n_sub, wts):
    sim.set_frame(0, st['x'], st['v'], st['C'], st['F'], st['used'])
    for f in range(n_sub):
        sim.substep(f)
    fr = sim.get_frame(n_sub)
    return sum((wts[k] * fr[k]).sum() for k in ('x', 'v', 'C', 'F'))


@pytest.mark.parametrize("mat", [M.WATER, M.ELASTIC, M.ICECREAM, M.MILK_VIS, M.PLASTIC_DEMO])
@pytest.mark.parametrize("boundary", [None, dict(type='cylinder', xz_radius=0.2, xz_center=(0.5, 0.5), y_range=(0.3, 0.7))])
def test_substep_adjoint_matches_finite_differences(mat, boundary):
    rng = np.random.RandomState(3)
    n_grid, N, n_sub = 16, 120, 3
    if boundary is None:
        boundary = dict(type='cube', lower=(0.32, 0.32, 0.32), upper=(0.68, 0.68, 0.68))
    P = make_particles(_cloud(N, rng, 0.36, 0.64), mat, n_grid)
    sim = orc.OracleSim(n_grid, P, gravity=(0, -10, 0), boundary=boundary, precision=64, max_substeps_local=10)
    st = _random_state(sim, rng)
    wts = {k: rng.randn(*st[k].shape) for k in ('x', 'v', 'C', 'F')}
    _run_loss(sim, st, n_sub, wts)
    sim.reset_grad()
    sim.set_grad_frame(n_sub, wts['x'], wts['v'], wts['C'], wts['F'])
    for f in reversed(range(n_sub)):
        sim.substep_grad(f)
    g = sim.get_grad_frame(0)
    eps = 1e-6
    checked = 0
    for key in ('x', 'v', 'C', 'F'):
        flat = st[key].reshape(-1)
        for idx in rng.choice(flat.size, 6, replace=False):
            old = flat[idx]
            flat[idx] = old + eps; lp = _run_loss(sim, st, n_sub, wts)
            flat[idx] = old - eps; lm = _run_loss(sim, st, n_sub, wts)
            flat[idx] = old
            fd = (lp - lm) / (2 * eps)
            an = g[key].reshape(-1)[idx]
            assert abs(fd - an) <= 2e-5 * max(1.0, abs(fd), abs(an)), (key, idx, fd, an)
            checked += 1
    assert checked == 24


def _latte_like(precision, n_sub_steps, flux=2):
    """tiny LatteArt-like scene: coffee pool + parked milk + injector (envs/latteart_env.py:54-74)."""
    rng = np.random.RandomState(4)
    n_grid = 16
    n_coffee, n_milk = 150, 60
    x = np.concatenate([np.tile(M.NOWHERE, (n_milk, 1)), rng.uniform((0.35, 0.36, 0.35), (0.65, 0.45, 0.65), size=(n_coffee, 3))])
    mat = np.concatenate([np.full(n_milk, M.MILK), np.full(n_coffee, M.COFFEE)])
    used = np.concatenate([np.zeros(n_milk), np.ones(n_coffee)]).astype(np.int32)
    P = make_particles(x, mat, n_grid, used=used)
    bnd = dict(type='cylinder', xz_radius=0.2, xz_center=(0.5, 0.5), y_range=(0.34, 0.9))
    sim = orc.OracleSim(n_grid, P, gravity=(0, -20, 0), boundary=bnd, precision=precision, max_substeps_local=20)
    rv = np.random.RandomState(5).uniform(size=(20, flux, 3))
    sim.add_effector(type=1, action_dim=3, boundary=dict(type='cylinder', xz_radius=0.2, xz_center=(0.5, 0.5), y_range=(0.55, 0.55)),
                     radius=0.0075, flux=flux, inject_v=(0, -3, 0), inject_p=(0, 0, 0), locally_random=True, random_vector=rv,
                     act_range=np.where(used == 0)[0], max_action_steps=64)
    return sim, P


def _latte_loss(sim, P, actions, action_p, tgt, n_steps, weight=1.0):
    fr = dict(x=P['x'], v=np.zeros_like(P['x']), C=np.zeros((len(P['x']), 3, 3)), F=np.tile(np.eye(3), (len(P['x']), 1, 1)), used=P['used'])
    sim.enable_grad()
    sim.set_frame(0, fr['x'], fr['v'], fr['C'], fr['F'], fr['used'])
    sim.set_effector_state(0, 0, np.array([0.5, 0.55, 0.5, 1, 0, 0, 0, 0.0]))
    sim.apply_action_p(action_p)
    total = 0.0
    for s in range(n_steps):
        sim.step(actions[s])
        total += sim.loss_value(sim.cur_substep_local, M.MILK, weight, tgt[s])
    return total


def test_dloss_daction_injector_chain_fd():
    """End to end dLoss/dAction (velocity actions and the initial-position action) through the injector,
    across a checkpointed chunk boundary (T=20, 3 steps = 30 substeps), vs central differences in fp64."""
    n_steps = 3
    sim, P = _latte_like(64, n_steps)
    rng = np.random.RandomState(6)
    actions = rng.uniform(-0.004, 0.004, size=(n_steps, 3))
    action_p = np.array([0.47, 0.55, 0.52])
    tgt = [P['x'] * 0 + rng.uniform(0.4, 0.6, size=P['x'].shape) for _ in range(n_steps)]
    _latte_loss(sim, P, actions, action_p, tgt, n_steps)
    sim.reset_grad()
    for s in reversed(range(n_steps)):
        sim.loss_seed(sim.cur_substep_local, M.MILK, 1.0, tgt[s])
        sim.step_grad(actions[s])
    sim.apply_action_p_grad()
    g = sim.get_action_grad(n_steps)
    assert g.shape == (n_steps + 1, 3)
    eps = 1e-6
    for (i, j) in [(0, 0), (0, 2), (1, 0), (2, 2), (3, 0), (3, 2)]:
        def run(d):
            a, ap = actions.copy(), action_p.copy()
            if i < n_steps: a[i, j] += d
            else: ap[j] += d
            return _latte_loss(sim, P, a, ap, tgt, n_steps)
        fd = (run(eps) - run(-eps)) / (2 * eps)
        assert abs(fd - g[i, j]) <= 1e-5 * max(1.0, abs(fd)), (i, j, fd, g[i, j])
    # y is pinned by the effector's own boundary (y_range lower == upper) -> zero gradient
    assert np.allclose(g[:, 1], 0)
    assert np.abs(g[:, [0, 2]]).max() > 1e-6


# ------------------------------------------------------------------ SDF colliders (meshes/static.py, meshes/dynamic.py)
from conftest import sphere_sdf, box_sdf  # noqa: E402


def _rigid_scene(precision, collide_type, friction, softness, static=False, mat=M.ELASTIC):
    rng = np.random.RandomState(31)
    n_grid, N = 16, 150
    x = rng.uniform((0.40, 0.42, 0.40), (0.60, 0.58, 0.60), size=(N, 3))
    P = make_particles(x, mat, n_grid)
    sim = orc.OracleSim(n_grid, P, gravity=(0, -10, 0), boundary=dict(type='cube', lower=(0.25, 0.25, 0.25), upper=(0.75, 0.75, 0.75)),
                        precision=precision, max_substeps_local=20)
    sim.add_effector(type=0, action_dim=3, boundary=dict(type='cube', lower=(0.05, 0.05, 0.05), upper=(0.95, 0.95, 0.95)), max_action_steps=16,
                     init_pos=(0.5, 0.64, 0.5))
    vox, T = sphere_sdf(0.09, 0.2)
    sim.set_rigid_mesh(vox, T, friction=friction, softness=softness, collide_type=collide_type)
    if static:
        bv, bT = box_sdf((0.3, 0.05, 0.3), 0.4)
        bT = bT.copy(); bT[:3, 3] -= bT[:3, :3] @ np.array([0.5, 0.33, 0.5])  # world -> mesh: box centred at (0.5, 0.33, 0.5)
        sim.add_static(bv, bT, friction=0.5)
    return sim, P


def _rigid_loss(sim, P, actions, action_p, wts, n_steps):
    N = len(P['x'])
    sim.enable_grad()
    sim.set_frame(0, P['x'], np.zeros((N, 3)), np.zeros((N, 3, 3)), np.tile(np.eye(3), (N, 1, 1)), P['used'])
    sim.set_effector_state(0, 0, np.array([0.5, 0.64, 0.5, 1, 0, 0, 0, 0.0]))
    sim.apply_action_p(action_p)
    for s in range(n_steps):
        sim.step(actions[s])
    fr = sim.get_frame(sim.cur_substep_local)
    return float((wts * fr['x']).sum())


@pytest.mark.parametrize("collide_type,friction,softness,static", [('particle', 8.0, 100.0, False), ('grid', 0.5, 0.0, True), ('both', 8.0, 100.0, True)])
def test_sdf_collider_dloss_daction_fd(collide_type, friction, softness, static):
    """dLoss/dAction through Dynamic.collide (particle / grid level, soft influence, friction and sticky branches) and
    Static.collide, vs central differences in fp64."""
    n_steps = 2
    sim, P = _rigid_scene(64, collide_type, friction, softness, static)
    rng = np.random.RandomState(32)
    actions = np.array([[0.004, -0.03, 0.002], [-0.003, -0.03, 0.004]])
    action_p = np.array([0.5, 0.64, 0.5])
    wts = rng.randn(*P['x'].shape)
    _rigid_loss(sim, P, actions, action_p, wts, n_steps)
    sim.reset_grad()
    sim.set_grad_frame(sim.cur_substep_local, wts, np.zeros_like(wts), np.zeros((len(wts), 3, 3)), np.zeros((len(wts), 3, 3)))
    for s in reversed(range(n_steps)):
        sim.step_grad(actions[s])
    sim.apply_action_p_grad()
    g = sim.get_action_grad(n_steps)
    assert np.abs(g).max() > 1e-6, 'collider never touched the material'
    # The forward map is only piecewise smooth (hit / no-hit and influence thresholds jump, meshes/dynamic.py:97): grid nodes
    # crossing a threshold inside the FD interval add noise, so the end-to-end check uses a larger step and a looser bar;
    # the exact check of the collide adjoint itself is test_sdf_collide_unit_adjoint_fd below.
    eps = 1e-5
    for (i, j) in [(0, 0), (0, 1), (1, 1), (1, 2), (2, 0), (2, 1)]:
        def run(d):
            a, ap = actions.copy(), action_p.copy()
            if i < n_steps: a[i, j] += d
            else: ap[j] += d
            return _rigid_loss(sim, P, a, ap, wts, n_steps)
        fd = (run(eps) - run(-eps)) / (2 * eps)
        assert abs(fd - g[i, j]) <= 3e-3 * max(1.0, abs(fd), np.abs(g).max()), (i, j, fd, g[i, j])


@pytest.mark.parametrize("friction,softness,dynamic", [(8.0, 100.0, 1), (8.0, 0.0, 1), (0.5, 0.0, 1), (20.0, 0.0, 1), (0.5, 0.0, 0)])
def test_sdf_collide_unit_adjoint_fd(friction, softness, dynamic):
    """one collide evaluation: adjoints of (position, velocity, effector pos[f], pos[f+1]) vs central differences, all branches."""
    import ctypes as C
    L = orc.lib()
    vox, T = sphere_sdf(0.09, 0.2)
    vox = np.ascontiguousarray(vox, dtype=np.float64); T = np.ascontiguousarray(T, dtype=np.float64)
    L.orc_sdf_collide_eval.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double] + [C.c_void_p] * 4

    def ev(io, gout=None):
        out = np.zeros(3); gio = np.zeros(12)
        L.orc_sdf_collide_eval(32, vox.ctypes.data, T.ctypes.data, friction, softness, dynamic, 2e-4, io.ctypes.data, out.ctypes.data,
                               None if gout is None else gout.ctypes.data, gio.ctypes.data)
        return out, gio
    rng = np.random.RandomState(40)
    hits = 0
    c0 = np.array([0.5, 0.5, 0.5]) if dynamic else np.zeros(3)
    for _ in range(200):
        d = rng.randn(3); d /= np.linalg.norm(d)
        io = np.concatenate([c0 + d * rng.uniform(0.05, 0.115), rng.randn(3) * 0.5, c0, c0 + rng.randn(3) * 1e-4 * dynamic])
        gout = rng.randn(3)
        out, g = ev(io, gout)
        hits += int(not np.array_equal(out, io[3:6]))
        n_in = 12 if dynamic else 6
        fd = np.zeros(12)
        for i in range(3 if not dynamic else 0, n_in):
            e = np.zeros(12); e[i] = 1e-7
            fd[i] = ((ev(io + e)[0] - ev(io - e)[0]) * gout).sum() / 2e-7
        sel = slice(3, 6) if not dynamic else slice(0, 12)   # static colliders only need the velocity adjoint (node positions are constants)
        assert np.abs(fd[sel] - g[sel]).max() <= 1e-4 * max(1.0, np.abs(fd).max()), (fd, g)
    assert hits > 30


@pytest.mark.parametrize("friction,softness", [(8.0, 100.0), (0.5, 0.0), (20.0, 0.0)])
def test_sdf_collide_unit_adjoint_fd_with_rotation(friction, softness):
    """one Dynamic.collide evaluation with free pose quaternions (6-DOF Rigid effectors, agent_pouring.yaml): adjoints of
    (p, v, pos[f], pos[f+1], quat[f], quat[f+1]) vs central differences, on a box SDF (not rotation invariant)."""
    import ctypes as C
    from conftest import box_sdf
    L = orc.lib()
    vox, T = box_sdf(np.array([0.08, 0.05, 0.11]), 0.2)
    vox = np.ascontiguousarray(vox, dtype=np.float64); T = np.ascontiguousarray(T, dtype=np.float64)
    L.orc_sdf_collide_eval_q.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double] + [C.c_void_p] * 4

    def ev(io, gout=None):
        out = np.zeros(3); gio = np.zeros(20)
        L.orc_sdf_collide_eval_q(32, vox.ctypes.data, T.ctypes.data, friction, softness, 2e-4, io.ctypes.data, out.ctypes.data,
                                 None if gout is None else gout.ctypes.data, gio.ctypes.data)
        return out, gio
    rng = np.random.RandomState(41)
    hits = skipped = 0
    c0 = np.array([0.5, 0.5, 0.5])
    for _ in range(150):
        q0 = rng.randn(4); q0 /= np.linalg.norm(q0)
        dq = np.concatenate([[1.0], rng.randn(3) * 2e-4]); dq /= np.linalg.norm(dq)
        q1 = np.array([dq[0] * q0[0] - dq[1:] @ q0[1:], *(dq[0] * q0[1:] + q0[0] * dq[1:] + np.cross(dq[1:], q0[1:]))])
        d = rng.randn(3); d /= np.linalg.norm(d)
        io = np.concatenate([c0 + d * rng.uniform(0.03, 0.13), rng.randn(3) * 0.5, c0, c0 + rng.randn(3) * 1e-4, q0, q1])
        gout = rng.randn(3)
        out, g = ev(io, gout)
        hit = np.abs(out - io[3:6]).max() > 1e-9   # (round-off-level "hits" far from the surface only produce FD noise)
        hits += int(hit)
        if not hit:
            continue
        fd, fd2 = np.zeros(20), np.zeros(20)
        for i in range(20):
            e = np.zeros(20); e[i] = 1e-7
            fd[i] = ((ev(io + e)[0] - ev(io - e)[0]) * gout).sum() / 2e-7
            fd2[i] = ((ev(io + 10 * e)[0] - ev(io - 10 * e)[0]) * gout).sum() / 2e-6
        if np.abs(fd - fd2).max() > 1e-3 * max(1.0, np.abs(fd).max()):
            # not differentiable here: on the medial planes of the box the baked SDF has a zero finite-difference gradient and the
            # reference's normal = g / |g|_eps (static.py:66-79) is round-off noise
            skipped += 1
            continue
        assert np.abs(fd - g).max() <= 2e-4 * max(1.0, np.abs(fd).max()), (fd, g)
    assert hits > 30 and skipped < 0.1 * hits, (hits, skipped)


# ------------------------------------------------------------------------------------------------
# MAT_RIGID bodies: shape matching (MPM:449-505) and its adjoint (MPM:436-447, 485-489)
# ------------------------------------------------------------------------------------------------
def _rigid_mat_scene(precision, rng, n_grid=16):
    """water pool + two rigid cuboids (different sizes so the singular values of H are separated) + one elastic blob"""
    xa = rng.uniform((0.40, 0.50, 0.40), (0.50, 0.56, 0.47), size=(60, 3))
    xb = rng.uniform((0.52, 0.48, 0.50), (0.60, 0.60, 0.56), size=(50, 3))
    xw = rng.uniform((0.36, 0.36, 0.36), (0.64, 0.46, 0.64), size=(120, 3))
    xe = rng.uniform((0.40, 0.58, 0.52), (0.48, 0.64, 0.60), size=(30, 3))
    x = np.concatenate([xw, xa, xe, xb])
    mat = np.concatenate([np.full(120, M.WATER), np.full(60, M.RIGID), np.full(30, M.ELASTIC), np.full(50, M.RIGID_HEAVY)])
    bid = np.concatenate([np.zeros(120), np.ones(60), np.full(30, 2), np.full(50, 3)]).astype(np.int32)
    P = make_particles(x, mat, n_grid)
    P['body_id'] = bid; P['bodies'] = {'n': 4}
    sim = orc.OracleSim(n_grid, P, gravity=(0, -10, 0), boundary=dict(type='cube', lower=(0.32, 0.32, 0.32), upper=(0.68, 0.68, 0.68)),
                        precision=precision, max_substeps_local=10)
    sim.set_bodies(bid, 4)
    return sim, P, bid


def test_rigid_bodies_keep_their_shape():
    rng = np.random.RandomState(11)
    sim, P, bid = _rigid_mat_scene(64, rng)
    st = _random_state(sim, rng, amp_F=0.0, amp_C=0.0, amp_v=0.8)
    x0 = st['x']
    for f in range(8):
        sim.substep(f)
    x1 = sim.get_frame(8)['x']
    for b in (1, 3):
        idx = np.where(bid == b)[0]
        d0 = np.linalg.norm(x0[idx][:, None] - x0[idx][None], axis=-1)
        d1 = np.linalg.norm(x1[idx][:, None] - x1[idx][None], axis=-1)
        assert np.abs(d1 - d0).max() < 1e-12, 'pairwise distances of a MAT_RIGID body must be preserved (x <- R (x - c0) + c1)'
        assert np.abs(x1[idx] - x0[idx]).max() > 1e-5   # ... while it actually moved
    idx = np.where(bid == 2)[0]   # the elastic blob is free to deform
    d0 = np.linalg.norm(x0[idx][:, None] - x0[idx][None], axis=-1); d1 = np.linalg.norm(x1[idx][:, None] - x1[idx][None], axis=-1)
    assert np.abs(d1 - d0).max() > 1e-7


def test_rigid_body_adjoint_matches_finite_differences():
    rng = np.random.RandomState(12)
    sim, P, bid = _rigid_mat_scene(64, rng)
    st = _random_state(sim, rng)
    n_sub = 3
    wts = {k: rng.randn(*st[k].shape) for k in ('x', 'v', 'C', 'F')}
    _run_loss(sim, st, n_sub, wts)
    sim.reset_grad()
    sim.set_grad_frame(n_sub, wts['x'], wts['v'], wts['C'], wts['F'])
    for f in reversed(range(n_sub)):
        sim.substep_grad(f)
    g = sim.get_grad_frame(0)
    eps = 1e-6
    rigid = np.where((bid == 1) | (bid == 3))[0]
    for key, width in (('x', 3), ('v', 3), ('C', 9), ('F', 9)):
        flat = st[key].reshape(-1)
        picks = [int(p) * width + int(rng.randint(width)) for p in rng.choice(rigid, 4, replace=False)] + list(rng.choice(flat.size, 3, replace=False))
        for idx in picks:
            old = flat[idx]
            flat[idx] = old + eps; lp = _run_loss(sim, st, n_sub, wts)
            flat[idx] = old - eps; lm = _run_loss(sim, st, n_sub, wts)
            flat[idx] = old
            fd = (lp - lm) / (2 * eps)
            an = g[key].reshape(-1)[idx]
            assert abs(fd - an) <= 2e-5 * max(1.0, abs(fd), abs(an)), (key, idx, fd, an)


# ------------------------------------------------------------------------------------------------
# 6-DOF effectors (agent_pouring.yaml, agent_transporting.yaml) and the collector agents
# ------------------------------------------------------------------------------------------------
def _pouring_like(precision):
    """rotating + translating box collider ('both' collide type, AgentPouring) over an elastic blob, with a collector"""
    rng = np.random.RandomState(51)
    n_grid, N = 16, 150
    x = rng.uniform((0.40, 0.42, 0.40), (0.60, 0.58, 0.60), size=(N, 3))
    P = make_particles(x, M.ELASTIC, n_grid)
    sim = orc.OracleSim(n_grid, P, gravity=(0, -10, 0), boundary=dict(type='cube', lower=(0.25, 0.25, 0.25), upper=(0.75, 0.75, 0.75)),
                        precision=precision, max_substeps_local=20)
    sim.add_effector(type=0, action_dim=6, boundary=dict(type='cube', lower=(0.05, 0.05, 0.05), upper=(0.95, 0.95, 0.95)), max_action_steps=16,
                     init_pos=(0.5, 0.64, 0.5), scale_v=(1, 1, 1, 1, 1, 1))
    vox, T = box_sdf(np.array([0.12, 0.05, 0.08]), 0.2)
    sim.set_rigid_mesh(vox, T, friction=8.0, softness=100.0, collide_type='both')
    sim.set_collector(dict(type='cube', lower=(0.0, 0.0, 0.0), upper=(1.0, 1.0, 0.585)), mat=-1)
    return sim, P


def _pose_loss(sim, P, actions, action_p, wts, n_steps, init):
    N = len(P['x'])
    sim.enable_grad()
    sim.set_frame(0, P['x'], np.zeros((N, 3)), np.zeros((N, 3, 3)), np.tile(np.eye(3), (N, 1, 1)), P['used'])
    sim.set_effector_state(0, 0, np.array(list(init) + [0.0]))
    sim.apply_action_p(action_p)
    for s in range(n_steps):
        sim.step(actions[s])
    fr = sim.get_frame(sim.cur_substep_local)
    return float((wts * fr['x'] * fr['used'][:, None]).sum()), fr


def test_6dof_rigid_collider_and_collector_dloss_daction_fd():
    """dLoss/dAction for a 6-DOF Rigid (translation + rotation actions, agent_pouring.yaml) through Dynamic.collide at grid and
    particle level, with the collector removing particles, vs central differences in fp64."""
    n_steps = 2
    sim, P = _pouring_like(64)
    rng = np.random.RandomState(52)
    actions = np.array([[0.004, -0.03, 0.002, 0.02, -0.03, 0.05], [-0.003, -0.03, 0.004, -0.04, 0.02, 0.03]])
    action_p = np.array([0.5, 0.64, 0.5, 0, 0, 0.0])
    init = (0.5, 0.64, 0.5, np.cos(0.2), 0.0, np.sin(0.2), 0.0)
    wts = rng.randn(*P['x'].shape)
    _, fr = _pose_loss(sim, P, actions, action_p, wts, n_steps, init)
    n_collected = int(P['used'].sum() - fr['used'].sum())
    assert 0 < n_collected < len(wts) // 2, n_collected
    assert np.all(fr['x'][fr['used'] == 0] == -100.0)   # parked at NOWHERE (agent_pouring.py:37-38)
    sim.reset_grad()
    sim.set_grad_frame(sim.cur_substep_local, wts * fr['used'][:, None], np.zeros_like(wts), np.zeros((len(wts), 3, 3)), np.zeros((len(wts), 3, 3)))
    for s in reversed(range(n_steps)):
        sim.step_grad(actions[s])
    sim.apply_action_p_grad()
    g = sim.get_action_grad(n_steps)
    assert g.shape == (n_steps + 1, 6)
    assert np.abs(g[:n_steps, 3:]).max() > 1e-6, 'rotation actions carry no gradient'
    assert np.all(g[n_steps, 3:] == 0)   # apply_action_p only sets the position (effector.py:223-226 "TODO: add orientation")
    # hit / influence / collector thresholds make the forward map piecewise smooth (see test_sdf_collider_dloss_daction_fd): the
    # central difference is taken at three step sizes and the closest one must agree (exact check: the unit tests above)
    for (i, j) in [(0, 0), (0, 3), (0, 4), (0, 5), (1, 1), (1, 3), (1, 4), (1, 5), (2, 0), (2, 1)]:
        def run(d):
            a, ap = actions.copy(), action_p.copy()
            if i < n_steps: a[i, j] += d
            else: ap[j] += d
            return _pose_loss(sim, P, a, ap, wts, n_steps, init)[0]
        fds = [(run(eps) - run(-eps)) / (2 * eps) for eps in (1e-5, 1e-6, 1e-7)]
        err = min(abs(fd - g[i, j]) for fd in fds)
        assert err <= 3e-3 * max(1.0, np.abs(g).max()), (i, j, fds, g[i, j])
        if i == 1 and j >= 3:   # the last step's rotation columns are smooth here: tight bar
            assert err <= 1e-5 * max(1.0, abs(g[i, j])), (i, j, fds, g[i, j])


def test_6dof_injector_and_collector_dloss_daction_fd():
    """AgentJetBot-like: an Injector whose pose rotates (inject_p / inject_v are rotated by quat[f], injector.py:93-96) plus a
    WATER collector; dLoss/dAction (6 columns) vs central differences in fp64."""
    rng = np.random.RandomState(53)
    n_grid, n_pool, n_parked, flux, n_steps = 16, 120, 80, 2, 3
    x = np.concatenate([np.tile(M.NOWHERE, (n_parked, 1)), rng.uniform((0.36, 0.36, 0.36), (0.64, 0.44, 0.64), size=(n_pool, 3))])
    used = np.concatenate([np.zeros(n_parked), np.ones(n_pool)]).astype(np.int32)
    P = make_particles(x, M.WATER, n_grid, used=used)
    sim = orc.OracleSim(n_grid, P, gravity=(0, -10, 0), boundary=dict(type='cube', lower=(0.3, 0.3, 0.3), upper=(0.7, 0.7, 0.7)),
                        precision=64, max_substeps_local=20)
    rv = np.random.RandomState(54).uniform(size=(64, flux, 3))
    sim.add_effector(type=1, action_dim=6, scale_v=(1, 1, 1, 5, 5, 5), boundary=dict(type='cube', lower=(0.1, 0.1, 0.1), upper=(0.9, 0.9, 0.9)),
                     radius=0.015, flux=flux, inject_v=(-3.0, 0, 0), inject_p=(-0.07, 0, 0), locally_random=False, random_vector=rv,
                     act_range=np.where(used == 0)[0], max_action_steps=64)
    sim.set_collector(dict(type='cube', lower=(0.0, 0.0, 0.0), upper=(1.0, 1.0, 0.60)), mat=M.WATER)
    actions = np.array([[0.003, -0.002, 0.001, 0.02, 0.03, -0.02], [-0.002, 0.001, 0.002, -0.01, 0.02, 0.03], [0.001, 0.0, -0.002, 0.03, -0.02, 0.01]])
    action_p = np.array([0.58, 0.55, 0.5, 0, 0, 0.0])
    init = (0.58, 0.55, 0.5, np.cos(0.3), np.sin(0.3) * 0.6, 0.0, np.sin(0.3) * 0.8)
    wts = rng.randn(*x.shape)
    _, fr = _pose_loss(sim, P, actions, action_p, wts, n_steps, init)
    assert int(fr['used'].sum()) < n_pool + flux * 10 * n_steps, 'nothing was collected'
    sim.reset_grad()
    sim.set_grad_frame(sim.cur_substep_local, wts * fr['used'][:, None], np.zeros_like(wts), np.zeros((len(wts), 3, 3)), np.zeros((len(wts), 3, 3)))
    for s in reversed(range(n_steps)):
        sim.step_grad(actions[s])
    sim.apply_action_p_grad()
    g = sim.get_action_grad(n_steps)
    assert g.shape == (n_steps + 1, 6) and np.abs(g[:n_steps, 3:]).max() > 1e-6
    eps = 1e-6
    for (i, j) in [(0, 0), (0, 3), (0, 4), (0, 5), (1, 2), (1, 4), (2, 3), (2, 5), (3, 0), (3, 2)]:
        def run(d):
            a, ap = actions.copy(), action_p.copy()
            if i < n_steps: a[i, j] += d
            else: ap[j] += d
            return _pose_loss(sim, P, a, ap, wts, n_steps, init)[0]
        fd = (run(eps) - run(-eps)) / (2 * eps)
        assert abs(fd - g[i, j]) <= 1e-4 * max(1.0, abs(fd), np.abs(g).max()), (i, j, fd, g[i, j])


# ------------------------------------------------------------------------------------------------
# symmetry pins of the restatement, fp64
# ------------------------------------------------------------------------------------------------
def _sym_run(x, v, C, F, mat, n_grid, gravity, boundary, n_sub=6):
    P = make_particles(x, mat, n_grid)
    sim = orc.OracleSim(n_grid, P, gravity=gravity, boundary=boundary, precision=64, max_substeps_local=10)
    sim.set_frame(0, x, v, C, F, P['used'])
    for f in range(n_sub):
        sim.substep(f)
    return sim.get_frame(n_sub)


def _sym_state(rng, N, lo, hi):
    x = rng.uniform(lo, hi, size=(N, 3))
    return x, rng.randn(N, 3) * 0.4, rng.randn(N, 3, 3) * 2.0, np.eye(3)[None] + rng.randn(N, 3, 3) * 0.02


@pytest.mark.parametrize("mat", [M.WATER, M.ELASTIC, M.ICECREAM])
def test_particle_order_does_not_matter(mat):
    rng = np.random.RandomState(61)
    n_grid, N = 16, 200
    bnd = dict(type='cube', lower=(0.3, 0.3, 0.3), upper=(0.7, 0.7, 0.7))
    x, v, C, F = _sym_state(rng, N, 0.36, 0.64)
    a = _sym_run(x, v, C, F, mat, n_grid, (0, -10, 0), bnd)
    perm = rng.permutation(N)
    b = _sym_run(x[perm], v[perm], C[perm], F[perm], mat, n_grid, (0, -10, 0), bnd)
    for k in ('x', 'v', 'C', 'F'):
        assert np.abs(a[k][perm] - b[k]).max() < 1e-11 * max(1.0, np.abs(a[k]).max()), k


@pytest.mark.parametrize("mat", [M.WATER, M.ELASTIC])
def test_grid_aligned_translation_equivariance(mat):
    """shifting particles AND walls by whole cells shifts the result by the same vector (base / fx / weights / boundary logic)"""
    rng = np.random.RandomState(62)
    n_grid, N = 32, 200
    dx = 1.0 / n_grid
    shift = np.array([3, -2, 5]) * dx
    lo, hi = np.array([0.3, 0.3, 0.3]), np.array([0.6, 0.6, 0.6])
    x, v, C, F = _sym_state(rng, N, 0.34, 0.56)
    a = _sym_run(x, v, C, F, mat, n_grid, (0, -10, 0), dict(type='cube', lower=tuple(lo), upper=tuple(hi)))
    b = _sym_run(x + shift, v, C, F, mat, n_grid, (0, -10, 0), dict(type='cube', lower=tuple(lo + shift), upper=tuple(hi + shift)))
    assert np.abs(a['x'] + shift - b['x']).max() < 1e-9      # wall positions are rounded to f32 first (boundaries.py:99-104)
    for k in ('v', 'C', 'F'):
        assert np.abs(a[k] - b[k]).max() < 1e-6 * max(1.0, np.abs(a[k]).max()), k


@pytest.mark.parametrize("mat", [M.WATER, M.ELASTIC, M.PLASTIC_DEMO])
def test_axis_permutation_equivariance(mat):
    """relabelling the axes (x,y,z) -> (z,x,y), with gravity and walls relabelled too, relabels the result: no axis is special in
    p2g / grid_op / g2p / SVD / F-update (catches transposed C or F conventions and swapped stencil indices)"""
    rng = np.random.RandomState(63)
    n_grid, N = 16, 200
    x, v, C, F = _sym_state(rng, N, 0.36, 0.64)
    g = np.array([1.0, -10.0, 3.0])
    lo, hi = np.array([0.30, 0.32, 0.28]), np.array([0.70, 0.66, 0.72])
    a = _sym_run(x, v, C, F, mat, n_grid, tuple(g), dict(type='cube', lower=tuple(lo), upper=tuple(hi)))
    p = [2, 0, 1]                      # new axis i = old axis p[i]
    Pm = np.eye(3)[p]                  # y = Pm @ x
    rot = lambda Mx: np.einsum('ij,njk,lk->nil', Pm, Mx, Pm)
    b = _sym_run(x[:, p], v[:, p], rot(C), rot(F), mat, n_grid, tuple(g[p]), dict(type='cube', lower=tuple(lo[p]), upper=tuple(hi[p])))
    assert np.abs(a['x'][:, p] - b['x']).max() < 1e-11 and np.abs(a['v'][:, p] - b['v']).max() < 1e-9
    assert np.abs(rot(a['C']) - b['C']).max() < 1e-8 and np.abs(rot(a['F']) - b['F']).max() < 1e-10
