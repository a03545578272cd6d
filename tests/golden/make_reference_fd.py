"""Central finite differences THROUGH THE REAL REFERENCE FORWARD CODE, in float64 (build container only).

The reference's macros carry a `dprecision = 64` switch (configs/macros.py:207-211); with DTYPE_TI / DTYPE_NP set to float64 before the
other modules import them, and the Taichi API emulated by tests/golden/taichi_emu.py in float64, the unmodified forward kernels become a
smooth-enough fp64 program whose derivative can be measured.  This script measures d(loss)/d(action) for the 6-DOF injector scene
(real AgentJetBot-free variant: AgentInjector, the reference's own ShapeMatchingLoss) and d(sum w.state)/d(state0) for a free multi-material
cloud, and stores them in tests/golden/reference_fd.npz.  tests/test_reference_run.py requires the oracle's hand-written ADJOINTS to match:
the adjoint side pinned to the reference's own forward function (Taichi's autodiff itself cannot run here).

    python tests/golden/make_reference_fd.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('FLUIDLAB_REFERENCE', '/root/reference')
sys.path.insert(0, HERE); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))


def load_reference_f64():
    import taichi_emu
    taichi_emu.default_fp = np.float64
    ti = taichi_emu.install()
    taichi_emu.stub_optional_dependencies()
    taichi_emu.install_value_semantics(REF)
    sys.path.insert(0, REF)
    macros = importlib.import_module('fluidlab.configs.macros')
    import torch
    macros.dprecision, macros.DTYPE_TI, macros.DTYPE_NP, macros.DTYPE_TC = 64, ti.f64, np.float64, torch.float64   # the reference's own fp64 switch
    return dict(macros=macros, sim=importlib.import_module('fluidlab.fluidengine.simulators.mpm_simulator'),
                agents=importlib.import_module('fluidlab.fluidengine.agents'),
                loss=importlib.import_module('fluidlab.fluidengine.losses.shapematching_loss'))


def read_frame(S, f):
    N = S.n_particles
    x, v = np.zeros((N, 3)), np.zeros((N, 3))
    C, F, u = np.zeros((N, 3, 3)), np.zeros((N, 3, 3)), np.zeros((N,), np.int32)
    S.readframe(f, x, v, C, F, u)
    return dict(x=x, v=v, C=C, F=F, used=u)


# ------------------------------------------------------------------------------------------------ A: substep adjoint of a free cloud
def cloud_inputs():
    from fluidlab_b200 import macros as M
    rng = np.random.RandomState(301)
    N, n_grid, n_sub = 60, 16, 3
    mats = [M.WATER, M.ELASTIC, M.ICECREAM, M.MILK_VIS]
    x = rng.uniform(0.36, 0.64, size=(N, 3))
    mat = np.array([mats[i % 4] for i in range(N)], dtype=np.int32)
    v = rng.randn(N, 3) * 0.5; C = rng.randn(N, 3, 3) * 4.0; F = np.eye(3)[None] + rng.randn(N, 3, 3) * 0.02
    w = {k: rng.randn(*a.shape) for k, a in (('x', x), ('v', v), ('C', C), ('F', F))}
    picks = [(k, int(i)) for k in ('x', 'v', 'C', 'F') for i in rng.choice(dict(x=x, v=v, C=C, F=F)[k].size, 5, replace=False)]
    return dict(N=N, n_grid=n_grid, n_sub=n_sub, x=x, v=v, C=C, F=F, mat=mat, w=w, picks=picks, lower=(0.32, 0.32, 0.32), upper=(0.68, 0.68, 0.68))


def cloud_loss(R, c, st):
    S = R['sim'].MPMSimulator(dim=3, quality=c['n_grid'] / 64, gravity=(0.0, -10.0, 0.0), horizon=10, max_substeps_local=10, max_substeps_global=1000, ckpt_dest='cpu')
    S.setup_boundary(type='cube', lower=c['lower'], upper=c['upper'])
    N = c['N']
    rho = np.array([R['macros'].RHO[int(m)] for m in c['mat']])
    S.build(None, None, [], dict(x=st['x'], used=np.ones(N), mat=c['mat'], rho=rho, body_id=np.zeros(N), bodies={'n': 1}))
    S.setframe(0, st['x'], st['v'], st['C'], st['F'], np.ones(N, np.int32))
    for f in range(c['n_sub']):
        S.substep(f, True)
    fr = read_frame(S, c['n_sub'])
    return float(sum((c['w'][k] * fr[k]).sum() for k in ('x', 'v', 'C', 'F')))


# ------------------------------------------------------------------------------------------------ B: dLoss/dAction, 6-DOF injector
def injector_inputs():
    from fluidlab_b200 import macros as M
    rng = np.random.RandomState(302)
    n_grid, n_pool, n_parked, flux, n_steps, T = 16, 70, 60, 2, 2, 20
    x = np.concatenate([np.tile(M.NOWHERE, (n_parked, 1)), rng.uniform((0.36, 0.36, 0.36), (0.64, 0.44, 0.64), size=(n_pool, 3))])
    used = np.concatenate([np.zeros(n_parked), np.ones(n_pool)]).astype(np.int32)
    mat = np.concatenate([np.full(n_parked, M.MILK), np.full(n_pool, M.COFFEE)]).astype(np.int32)
    actions = np.array([[0.003, -0.002, 0.001, 0.02, 0.03, -0.02], [-0.002, 0.001, 0.002, -0.01, 0.02, 0.03]])
    action_p = np.array([0.58, 0.55, 0.5, 0.0, 0.0, 0.0])
    tgt = rng.uniform(0.4, 0.6, size=(n_steps, len(x), 3))
    return dict(n_grid=n_grid, flux=flux, n_steps=n_steps, T=T, x=x, used=used, mat=mat, actions=actions, action_p=action_p, tgt=tgt,
                lower=(0.3, 0.3, 0.3), upper=(0.7, 0.7, 0.7), e_lower=(0.1, 0.1, 0.1), e_upper=(0.9, 0.9, 0.9),
                picks=[(0, 0), (0, 2), (0, 3), (0, 4), (0, 5), (1, 1), (1, 3), (1, 5), (2, 0), (2, 1), (2, 2)])


def injector_loss(R, c, actions, action_p, want_aux=False):
    import taichi as ti
    from fluidlab_b200 import macros as M
    N = len(c['x'])
    common = dict(max_substeps_local=c['T'], max_substeps_global=1000, max_action_steps_global=20, ckpt_dest='cpu')
    np.random.seed(55)
    agent = R['agents'].AgentInjector(**common)
    agent.add_effector(type='Injector', params=dict(radius=0.015, flux=c['flux'], init_pos=(0.58, 0.55, 0.5), init_euler=(20.0, 35.0, -10.0), inject_v=(-3.0, 0.5, 0.0),
                                                     inject_p=(-0.07, 0.01, 0.0), action_dim=6, action_scale_p=(1.0,) * 6, action_scale_v=(1.0, 1.0, 1.0, 5.0, 5.0, 5.0),
                                                     locally_random=True),
                       mesh_cfg=None, boundary_cfg=dict(type='cube', lower=c['e_lower'], upper=c['e_upper']))
    S = R['sim'].MPMSimulator(dim=3, quality=c['n_grid'] / 64, gravity=(0.0, -10.0, 0.0), horizon=10, max_substeps_local=c['T'], max_substeps_global=1000, ckpt_dest='cpu')
    S.setup_boundary(type='cube', lower=c['lower'], upper=c['upper'])
    rho = np.array([R['macros'].RHO[int(m)] for m in c['mat']])
    S.build(agent, None, [], dict(x=c['x'], used=c['used'], mat=c['mat'], rho=rho, body_id=np.zeros(N), bodies={'n': 1}))
    agent.build(S)
    loss = R['loss'].ShapeMatchingLoss(matching_mat=M.MILK, temporal_range_type='all', max_loss_steps=c['n_steps'], weights={'chamfer': 1.0}, target_file=None)
    loss.build(S)
    loss.target = {'x': [c['tgt'][i] for i in range(c['n_steps'])]}
    loss.tgt_particles_x = ti.Vector.field(3, dtype=R['macros'].DTYPE_TI, shape=N)
    agent.apply_action_p(action_p)
    for i in range(c['n_steps']):
        S.step(actions[i]); loss.step()
    val = float(loss.get_final_loss()['loss'])
    if want_aux:
        inj = agent.effectors[0]
        return val, dict(random_vector=inj.random_vector.to_numpy(), init_state=np.asarray(inj.init_state, dtype=np.float64))
    return val


# ------------------------------------------------------------------------------------------------ C: MAT_RIGID bodies (advect_grad)
def rigid_inputs():
    from fluidlab_b200 import macros as M
    rng = np.random.RandomState(303)
    n_grid, n_sub = 16, 3
    xw = rng.uniform((0.38, 0.36, 0.38), (0.62, 0.46, 0.62), size=(40, 3))
    xa = rng.uniform((0.40, 0.50, 0.40), (0.50, 0.56, 0.47), size=(24, 3))
    xb = rng.uniform((0.50, 0.48, 0.50), (0.58, 0.60, 0.56), size=(20, 3))
    x = np.concatenate([xw, xa, xb])
    mat = np.concatenate([np.full(40, M.WATER), np.full(24, M.RIGID), np.full(20, M.RIGID_HEAVY)]).astype(np.int32)
    bid = np.concatenate([np.zeros(40), np.ones(24), np.full(20, 2)]).astype(np.int32)
    N = len(x)
    v = rng.randn(N, 3) * 0.5; C = rng.randn(N, 3, 3) * 3.0; F = np.eye(3)[None] + rng.randn(N, 3, 3) * 0.02
    w = {k: rng.randn(*a.shape) for k, a in (('x', x), ('v', v), ('C', C), ('F', F))}
    rigid = np.where(bid > 0)[0]
    picks = [(k, int(p) * (3 if k in 'xv' else 9) + int(rng.randint(3 if k in 'xv' else 9))) for k in ('x', 'v', 'C', 'F') for p in rng.choice(rigid, 4, replace=False)]
    return dict(N=N, n_grid=n_grid, n_sub=n_sub, x=x, v=v, C=C, F=F, mat=mat, bid=bid, w=w, picks=picks, lower=(0.32, 0.32, 0.32), upper=(0.68, 0.68, 0.68))


def rigid_loss(R, c, st):
    S = R['sim'].MPMSimulator(dim=3, quality=c['n_grid'] / 64, gravity=(0.0, -10.0, 0.0), horizon=10, max_substeps_local=10, max_substeps_global=1000, ckpt_dest='cpu')
    S.setup_boundary(type='cube', lower=c['lower'], upper=c['upper'])
    N = c['N']
    rho = np.array([R['macros'].RHO[int(m)] for m in c['mat']])
    S.build(None, None, [], dict(x=st['x'], used=np.ones(N), mat=c['mat'], rho=rho, body_id=c['bid'], bodies={'n': 3}))
    S.setframe(0, st['x'], st['v'], st['C'], st['F'], np.ones(N, np.int32))
    for f in range(c['n_sub']):
        S.substep(f, True)
    fr = read_frame(S, c['n_sub'])
    return float(sum((c['w'][k] * fr[k]).sum() for k in ('x', 'v', 'C', 'F')))


# ------------------------------------------------------------------------------------------------ D: dLoss/dAction, 6-DOF Rigid collider
def pouring_inputs():
    from fluidlab_b200 import macros as M
    rng = np.random.RandomState(304)
    n_grid, N, n_steps, T = 16, 70, 2, 20
    x = rng.uniform((0.40, 0.42, 0.40), (0.60, 0.58, 0.60), size=(N, 3))
    res, he, half, ctr = 32, 0.2, np.array([0.12, 0.05, 0.08]), np.array([0.0043, 0.0031, -0.0052])
    ax = np.linspace(-he, he, res)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing='ij')
    q = np.stack([np.abs(X - ctr[0]) - half[0], np.abs(Y - ctr[1]) - half[1], np.abs(Z - ctr[2]) - half[2]], -1)
    vox = (np.linalg.norm(np.maximum(q, 0), axis=-1) + np.minimum(q.max(-1), 0)).astype(np.float32)
    sc = (res - 1) / (2 * he)
    Tm = np.eye(4); Tm[0, 0] = Tm[1, 1] = Tm[2, 2] = sc; Tm[:3, 3] = sc * he
    return dict(n_grid=n_grid, N=N, n_steps=n_steps, T=T, x=x, mat=np.full(N, M.ELASTIC, dtype=np.int32), vox=vox, Tm=Tm, w=rng.randn(N, 3),
                v0=rng.randn(N, 3) * 0.05, C0=rng.randn(N, 3, 3) * 0.5, F0=np.eye(3)[None] + rng.randn(N, 3, 3) * 0.02,
                actions=np.array([[0.004, -0.03, 0.002, 0.02, -0.03, 0.05], [-0.003, -0.03, 0.004, -0.04, 0.02, 0.03]]), action_p=np.array([0.5, 0.64, 0.5, 0.0, 0.0, 0.0]),
                lower=(0.25, 0.25, 0.25), upper=(0.75, 0.75, 0.75), e_lower=(0.05, 0.05, 0.05), e_upper=(0.95, 0.95, 0.95),
                picks=[(0, 1), (0, 3), (0, 4), (0, 5), (1, 0), (1, 3), (1, 5), (2, 1)])


def pouring_loss(R, c, actions, action_p, want_aux=False):
    from fluidlab_b200 import macros as M
    import make_reference_run as mr
    mr.SDF_REGISTRY['box.obj'] = dict(voxels=c['vox'], T_mesh_to_voxels=c['Tm'])
    common = dict(max_substeps_local=c['T'], max_substeps_global=1000, max_action_steps_global=20, ckpt_dest='cpu')
    agent = R['agents'].AgentRigid(collide_type='both', **common)
    agent.add_effector(type='Rigid', params=dict(init_pos=(0.5, 0.64, 0.5), init_euler=(0.0, 23.0, 5.0), action_dim=6, action_scale_p=(1.0,) * 6, action_scale_v=(1.0,) * 6),
                       mesh_cfg=dict(file='box.obj', material=M.STIRRER, softness=100.0, scale=(1.0, 0.9, 1.1), euler=(0.0, 10.0, 0.0)),
                       boundary_cfg=dict(type='cube', lower=c['e_lower'], upper=c['e_upper']))
    S = R['sim'].MPMSimulator(dim=3, quality=c['n_grid'] / 64, gravity=(0.0, -10.0, 0.0), horizon=10, max_substeps_local=c['T'], max_substeps_global=1000, ckpt_dest='cpu')
    S.setup_boundary(type='cube', lower=c['lower'], upper=c['upper'])
    N = c['N']
    S.build(agent, None, [], dict(x=c['x'], used=np.ones(N), mat=c['mat'], rho=np.ones(N), body_id=np.zeros(N), bodies={'n': 1}))
    agent.build(S)
    S.setframe(0, c['x'], c['v0'], c['C0'], c['F0'], np.ones(N, np.int32))
    agent.apply_action_p(action_p)
    for i in range(c['n_steps']):
        S.step(actions[i])
    fr = read_frame(S, S.cur_substep_local)
    val = float((c['w'] * fr['x']).sum())
    if want_aux:
        rigid = agent.effectors[0]
        return val, dict(T_final=np.asarray(rigid.mesh.T_mesh_to_voxels_np, dtype=np.float64), friction=float(rigid.mesh.friction),
                         init_state=np.asarray(rigid.init_state, dtype=np.float64))
    return val


def main():
    R = load_reference_f64()
    import make_reference_run as mr
    mr.patch_mesh_io(dict(macros=R['macros']))
    out = {}
    # C
    c = rigid_inputs()
    base = dict(x=c['x'], v=c['v'], C=c['C'], F=c['F'])
    fd = []
    for k, i in c['picks']:
        vals = []
        for sgn in (+1, -1):
            st = {q: a.copy() for q, a in base.items()}
            st[k].reshape(-1)[i] += sgn * 1e-6
            vals.append(rigid_loss(R, c, st))
        fd.append((vals[0] - vals[1]) / 2e-6)
    out.update(rb_x=c['x'], rb_v=c['v'], rb_C=c['C'], rb_F=c['F'], rb_mat=c['mat'], rb_bid=c['bid'], rb_n_grid=c['n_grid'], rb_n_sub=c['n_sub'], rb_lower=c['lower'],
               rb_upper=c['upper'], rb_pick_key=np.array([k for k, _ in c['picks']]), rb_pick_idx=np.array([i for _, i in c['picks']]), rb_fd=np.array(fd),
               rb_loss=rigid_loss(R, c, base), **{'rb_w_' + k: a for k, a in c['w'].items()})
    print('rigid bodies fd', np.array(fd)[:4])
    # D
    c = pouring_inputs()
    val, aux = pouring_loss(R, c, c['actions'], c['action_p'], want_aux=True)
    fds = []
    for eps in (1e-5, 1e-6, 1e-7):   # hit / influence thresholds make the map piecewise smooth: three step sizes are stored
        row = []
        for (i, j) in c['picks']:
            vals = []
            for sgn in (+1, -1):
                a, ap = c['actions'].copy(), c['action_p'].copy()
                if i < c['n_steps']:
                    a[i, j] += sgn * eps
                else:
                    ap[j] += sgn * eps
                vals.append(pouring_loss(R, c, a, ap))
            row.append((vals[0] - vals[1]) / (2 * eps))
        fds.append(row)
    out.update(po_x=c['x'], po_mat=c['mat'], po_vox=c['vox'], po_v0=c['v0'], po_C0=c['C0'], po_F0=c['F0'], po_w=c['w'], po_actions=c['actions'], po_action_p=c['action_p'],
               po_n_grid=c['n_grid'], po_n_steps=c['n_steps'], po_T=c['T'], po_lower=c['lower'], po_upper=c['upper'], po_e_lower=c['e_lower'], po_e_upper=c['e_upper'],
               po_picks=np.array(c['picks']), po_fd=np.array(fds), po_loss=val, po_T_final=aux['T_final'], po_friction=aux['friction'], po_init_state=aux['init_state'])
    print('pouring loss', val, 'fd', np.array(fds))
    # A
    c = cloud_inputs()
    base = dict(x=c['x'], v=c['v'], C=c['C'], F=c['F'])
    fd = []
    eps = 1e-6
    for k, i in c['picks']:
        vals = []
        for sgn in (+1, -1):
            st = {q: a.copy() for q, a in base.items()}
            st[k].reshape(-1)[i] += sgn * eps
            vals.append(cloud_loss(R, c, st))
        fd.append((vals[0] - vals[1]) / (2 * eps))
    out.update(cloud_x=c['x'], cloud_v=c['v'], cloud_C=c['C'], cloud_F=c['F'], cloud_mat=c['mat'], cloud_n_grid=c['n_grid'], cloud_n_sub=c['n_sub'],
               cloud_lower=c['lower'], cloud_upper=c['upper'], cloud_pick_key=np.array([k for k, _ in c['picks']]), cloud_pick_idx=np.array([i for _, i in c['picks']]),
               cloud_fd=np.array(fd), cloud_loss=cloud_loss(R, c, base), **{'cloud_w_' + k: a for k, a in c['w'].items()})
    print('cloud fd', np.array(fd)[:4])
    # B
    c = injector_inputs()
    val, aux = injector_loss(R, c, c['actions'], c['action_p'], want_aux=True)
    fd = []
    eps = 1e-6
    for (i, j) in c['picks']:
        vals = []
        for sgn in (+1, -1):
            a, ap = c['actions'].copy(), c['action_p'].copy()
            if i < c['n_steps']:
                a[i, j] += sgn * eps
            else:
                ap[j] += sgn * eps
            vals.append(injector_loss(R, c, a, ap))
        fd.append((vals[0] - vals[1]) / (2 * eps))
    out.update(inj_x=c['x'], inj_used=c['used'], inj_mat=c['mat'], inj_actions=c['actions'], inj_action_p=c['action_p'], inj_tgt=c['tgt'], inj_n_grid=c['n_grid'],
               inj_flux=c['flux'], inj_n_steps=c['n_steps'], inj_T=c['T'], inj_lower=c['lower'], inj_upper=c['upper'], inj_e_lower=c['e_lower'], inj_e_upper=c['e_upper'],
               inj_picks=np.array(c['picks']), inj_fd=np.array(fd), inj_loss=val, inj_random_vector=aux['random_vector'], inj_init_state=aux['init_state'])
    print('injector loss', val, 'fd', np.array(fd))
    # E: the reference's manual SVD adjoint MPM:272-292 (backward_svd + clamp), called directly on random inputs incl. near-degenerate sigmas
    S0 = R['sim'].MPMSimulator(dim=3, quality=0.25, gravity=(0.0, -10.0, 0.0), horizon=10, max_substeps_local=10, max_substeps_global=1000, ckpt_dest='cpu')
    import taichi as ti
    rng = np.random.RandomState(305)
    cases = []
    for it in range(24):
        A = np.eye(3) + rng.randn(3, 3) * (0.3 if it % 3 else 1e-5)       # every third case: nearly equal singular values (the 1e-8 clamp acts)
        U, Sg, Vh = np.linalg.svd(A); V = Vh.T
        gU, gV, gS = rng.randn(3, 3), rng.randn(3, 3), np.diag(rng.randn(3))
        r = S0.backward_svd(ti.Matrix(gU.tolist()), ti.Matrix(gS.tolist()), ti.Matrix(gV.tolist()), ti.Matrix(U.tolist()), ti.Matrix(np.diag(Sg).tolist()), ti.Matrix(V.tolist()))
        cases.append(np.concatenate([gU.ravel(), gS.ravel(), gV.ravel(), U.ravel(), Sg, V.ravel(), np.asarray(r, dtype=np.float64).ravel()]))
    out['svd_grad_cases'] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, 'reference_fd.npz'), **out)
    print('wrote', os.path.getsize(os.path.join(HERE, 'reference_fd.npz')), 'bytes')


if __name__ == '__main__':
    main()
