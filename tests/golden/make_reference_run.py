"""Fixtures from FORWARD RUNS OF THE REAL REFERENCE KERNELS (build container only).

`tests/golden/taichi_emu.py` lets the UNMODIFIED source of zhouxian/FluidLab's `MPMSimulator`, boundaries, effectors, agents and SDF
meshes execute eagerly on NumPy float32 (read its header for what is emulated: `ti.svd`, the fp32 summation order inside 3x3 products,
and no autodiff).  This script builds small scenes through the reference's own API, steps them with the reference's own
`substep` / `step`, and stores inputs + resulting particle state in `tests/golden/reference_run_<scene>.npz`.
`tests/test_reference_run.py` then requires the CPU oracle to reproduce those states — the pin of the oracle's FORWARD restatement
against the reference itself that DESIGN.md §2 asks for (the adjoints are pinned by finite differences and torch.autograd instead).

    python tests/golden/make_reference_run.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('FLUIDLAB_REFERENCE', '/root/reference')
sys.path.insert(0, HERE); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))


def load_reference():
    import taichi_emu
    taichi_emu.install()
    taichi_emu.stub_optional_dependencies()
    taichi_emu.install_value_semantics(REF)
    sys.path.insert(0, REF)
    mods = dict(sim=importlib.import_module('fluidlab.fluidengine.simulators.mpm_simulator'),
                macros=importlib.import_module('fluidlab.configs.macros'),
                agents=importlib.import_module('fluidlab.fluidengine.agents'),
                effectors=importlib.import_module('fluidlab.fluidengine.effectors'),
                meshes=importlib.import_module('fluidlab.fluidengine.meshes'))
    return mods


SDF_REGISTRY = {}   # file name -> {'voxels', 'T_mesh_to_voxels'}: synthetic baked volumes standing in for assets/meshes/processed/*.sdf


def patch_mesh_io(R):
    """replace ONLY the file I/O of Mesh.load_file (trimesh + pickle, meshes/mesh.py:41-66): geometry for rendering becomes one dummy vertex,
    the SDF volume comes from SDF_REGISTRY; init_transform / sdf_ / normal_ / collide stay the reference's own code"""
    mesh_mod = importlib.import_module('fluidlab.fluidengine.meshes.mesh')
    FRICTION, DT = R['macros'].FRICTION, R['macros'].DTYPE_NP

    def load_file(self):
        self.raw_vertices = np.zeros((1, 3), np.float32); self.raw_vertex_normals_np = np.zeros((1, 3), np.float32)
        self.faces_np = np.zeros(3, np.int32); self.n_vertices, self.n_faces = 1, 3
        self.colors_np = np.zeros((1, 4), np.float32)
        if self.has_dynamics:
            self.friction = FRICTION[self.material]
            sdf = SDF_REGISTRY[self.raw_file]
            self.sdf_voxels_np = sdf['voxels'].astype(DT); self.sdf_voxels_res = self.sdf_voxels_np.shape[0]
            self.T_mesh_to_voxels_np = sdf['T_mesh_to_voxels'].astype(DT)
    mesh_mod.Mesh.load_file = load_file


def read_frame(S, f):
    N = S.n_particles
    x, v = np.zeros((N, 3), np.float32), np.zeros((N, 3), np.float32)
    C, F, u = np.zeros((N, 3, 3), np.float32), np.zeros((N, 3, 3), np.float32), np.zeros((N,), np.int32)
    S.readframe(f, x, v, C, F, u)
    return dict(x=x, v=v, C=C, F=F, used=u)


# ---------------------------------------------------------------------------------------------------------------- scene 1
def scene_multimat(R):
    """every material class, random (v, C, F), cube walls that some particles hit, 12 substeps"""
    from fluidlab_b200 import macros as M
    rng = np.random.RandomState(201)
    n_grid, n_sub = 16, 12
    mats = [M.WATER, M.ELASTIC, M.ICECREAM, M.MILK_VIS, M.PLASTIC_DEMO, M.RIGID]   # RIGID here = one body with a single rigid material
    N = 96
    x = rng.uniform(0.33, 0.67, size=(N, 3)).astype(np.float32)
    mat = np.array([mats[i % 5] for i in range(N)], dtype=np.int32)
    used = (rng.rand(N) > 0.12).astype(np.int32)
    v = (rng.randn(N, 3) * 0.8).astype(np.float32); C = (rng.randn(N, 3, 3) * 4.0).astype(np.float32)
    F = (np.eye(3)[None] + rng.randn(N, 3, 3) * 0.03).astype(np.float32)
    bnd = dict(type='cube', lower=(0.3, 0.3, 0.3), upper=(0.7, 0.7, 0.7))
    S = R['sim'].MPMSimulator(dim=3, quality=n_grid / 64, gravity=(0.5, -10.0, 0.2), horizon=10, max_substeps_local=20, max_substeps_global=1000, ckpt_dest='cpu')
    S.setup_boundary(**bnd)
    rho = np.array([R['macros'].RHO[int(m)] for m in mat], dtype=np.float32)
    S.build(None, None, [], dict(x=x, used=used, mat=mat, rho=rho, body_id=np.zeros(N), bodies={'n': 1}))
    S.setframe(0, x, v, C, F, used)
    for f in range(n_sub):
        S.substep(f, True)
    out = read_frame(S, n_sub)
    return dict(n_grid=n_grid, n_sub=n_sub, gravity=(0.5, -10.0, 0.2), b_lower=bnd['lower'], b_upper=bnd['upper'], x0=x, v0=v, C0=C, F0=F, used0=used, mat=mat,
                **{'ref_' + k: a for k, a in out.items()})


# ---------------------------------------------------------------------------------------------------------------- scene 2
def scene_rigid_bodies(R):
    """two MAT_RIGID bodies (shape matching in advect, MPM:428-505) + water + an elastic blob inside a cylinder boundary, 10 substeps"""
    from fluidlab_b200 import macros as M
    rng = np.random.RandomState(202)
    n_grid, n_sub = 16, 10
    xw = rng.uniform((0.38, 0.36, 0.38), (0.62, 0.46, 0.62), size=(70, 3))
    xa = rng.uniform((0.40, 0.50, 0.40), (0.50, 0.56, 0.47), size=(30, 3))
    xe = rng.uniform((0.52, 0.50, 0.40), (0.58, 0.56, 0.48), size=(16, 3))
    xb = rng.uniform((0.50, 0.48, 0.50), (0.58, 0.60, 0.56), size=(24, 3))
    x = np.concatenate([xw, xa, xe, xb]).astype(np.float32)
    mat = np.concatenate([np.full(70, M.WATER), np.full(30, M.RIGID), np.full(16, M.ELASTIC), np.full(24, M.RIGID_HEAVY)]).astype(np.int32)
    bid = np.concatenate([np.zeros(70), np.ones(30), np.full(16, 2), np.full(24, 3)]).astype(np.int32)
    N = len(x)
    used = np.ones(N, np.int32)
    v = (rng.randn(N, 3) * 0.6).astype(np.float32); C = (rng.randn(N, 3, 3) * 3.0).astype(np.float32)
    F = (np.eye(3)[None] + rng.randn(N, 3, 3) * 0.02).astype(np.float32)
    bnd = dict(type='cylinder', xz_radius=0.17, xz_center=(0.5, 0.5), y_range=(0.34, 0.7))
    S = R['sim'].MPMSimulator(dim=3, quality=n_grid / 64, gravity=(0.0, -10.0, 0.0), horizon=10, max_substeps_local=20, max_substeps_global=1000, ckpt_dest='cpu')
    S.setup_boundary(**bnd)
    rho = np.array([R['macros'].RHO[int(m)] for m in mat], dtype=np.float32)
    S.build(None, None, [], dict(x=x, used=used, mat=mat, rho=rho, body_id=bid, bodies={'n': 4}))
    S.setframe(0, x, v, C, F, used)
    for f in range(n_sub):
        S.substep(f, True)
    out = read_frame(S, n_sub)
    return dict(n_grid=n_grid, n_sub=n_sub, gravity=(0.0, -10.0, 0.0), xz_radius=0.17, xz_center=(0.5, 0.5), y_range=(0.34, 0.7), x0=x, v0=v, C0=C, F0=F, used0=used,
                mat=mat, body_id=bid, **{'ref_' + k: a for k, a in out.items()})


# ---------------------------------------------------------------------------------------------------------------- scene 3
def scene_jetbot(R, randomize_inject_v=False):
    """the real AgentJetBot (agents/agent_jetbot.py): a 6-DOF Injector whose pose is driven through Agent.set_action -> set_velocity ->
    move_kernel (quaternion chain), injecting WATER into a pool, plus the collector; 3 steps through MPMSimulator.step"""
    from fluidlab_b200 import macros as M
    rng = np.random.RandomState(203)
    n_grid, n_pool, n_parked, flux, n_steps, T = 16, 90, 70, 2, 3, 20
    x = np.concatenate([np.tile(M.NOWHERE, (n_parked, 1)), rng.uniform((0.36, 0.36, 0.36), (0.64, 0.44, 0.64), size=(n_pool, 3))]).astype(np.float32)
    used = np.concatenate([np.zeros(n_parked), np.ones(n_pool)]).astype(np.int32)
    mat = np.full(len(x), M.WATER, dtype=np.int32)
    N = len(x)
    bnd = dict(type='cube', lower=(0.3, 0.3, 0.3), upper=(0.7, 0.7, 0.7))
    ebnd = dict(type='cube', lower=(0.1, 0.1, 0.1), upper=(0.9, 0.9, 0.9))
    cbnd = dict(type='cube', lower=(0.0, 0.0, 0.0), upper=(1.0, 1.0, 0.60))
    common = dict(max_substeps_local=T, max_substeps_global=1000, max_action_steps_global=20, ckpt_dest='cpu')
    np.random.seed(31)
    agent = R['agents'].AgentJetBot(collector_boundary=cbnd, **common)
    agent.add_effector(type='Injector', params=dict(radius=0.015, flux=flux, init_pos=(0.58, 0.55, 0.5), init_euler=(20.0, 35.0, -10.0), inject_v=(-3.0, 0.0, 0.0),
                                                     inject_p=(-0.07, 0.0, 0.0), randomize_inject_v=randomize_inject_v, action_dim=6, action_scale_p=(1.0,) * 6, action_scale_v=(1.0, 1.0, 1.0, 5.0, 5.0, 5.0)),
                       mesh_cfg=None, boundary_cfg=ebnd)
    S = R['sim'].MPMSimulator(dim=3, quality=n_grid / 64, gravity=(0.0, -10.0, 0.0), horizon=10, max_substeps_local=T, max_substeps_global=1000, ckpt_dest='cpu')
    S.setup_boundary(**bnd)
    S.build(agent, None, [], dict(x=x, used=used, mat=mat, rho=np.ones(N, np.float32), body_id=np.zeros(N), bodies={'n': 1}))
    agent.build(S)
    inj = agent.effectors[0]
    actions = np.array([[0.003, -0.002, 0.001, 0.02, 0.03, -0.02], [-0.002, 0.001, 0.002, -0.01, 0.02, 0.03], [0.001, 0.0, -0.002, 0.03, -0.02, 0.01]], dtype=np.float32)
    action_p = np.array([0.58, 0.55, 0.5, 0, 0, 0], dtype=np.float32)
    agent.apply_action_p(action_p)
    # the reference's own index-matched loss (losses/shapematching_loss.py:80-93) accumulated the way TaichiEnv.step does (taichi_env.py:171-172)
    sm = importlib.import_module('fluidlab.fluidengine.losses.shapematching_loss')
    import taichi as ti
    tgt = rng.uniform(0.4, 0.6, size=(n_steps, N, 3)).astype(np.float32)
    loss = sm.ShapeMatchingLoss(matching_mat=M.WATER, temporal_range_type='all', max_loss_steps=n_steps, weights={'chamfer': 1.0}, target_file=None)
    loss.build(S)
    loss.target = {'x': [tgt[i] for i in range(n_steps)]}
    loss.tgt_particles_x = ti.Vector.field(3, dtype=R['macros'].DTYPE_TI, shape=N)
    for i in range(n_steps):
        S.step(actions[i])
        loss.step()
    ref_loss = float(loss.get_final_loss()['loss'])
    out = read_frame(S, S.cur_substep_local)
    pose = np.asarray(inj.get_state(S.cur_substep_local), dtype=np.float64)
    return dict(n_grid=n_grid, n_steps=n_steps, T=T, flux=flux, x0=x, used0=used, mat=mat, b_lower=bnd['lower'], b_upper=bnd['upper'], e_lower=ebnd['lower'],
                e_upper=ebnd['upper'], c_lower=cbnd['lower'], c_upper=cbnd['upper'], actions=actions, action_p=action_p, init_state=np.asarray(inj.init_state, dtype=np.float64),
                random_vector=inj.random_vector.to_numpy(), ref_pose=pose, tgt=tgt, ref_loss=ref_loss, **{'ref_' + k: a for k, a in out.items()})


# ---------------------------------------------------------------------------------------------------------------- scene 2b
def scene_locked(R):
    """the boundary options of envs/transporting_env.py:80-87: a cube boundary with lock_dims=[2] (z velocity zeroed on every grid node) and a
    non-zero restitution (walls reflect), water + a MAT_RIGID body thrown against the walls; 10 substeps"""
    from fluidlab_b200 import macros as M
    rng = np.random.RandomState(206)
    n_grid, n_sub = 16, 10
    xw = rng.uniform((0.34, 0.33, 0.40), (0.50, 0.45, 0.60), size=(60, 3))
    xr = rng.uniform((0.55, 0.34, 0.42), (0.65, 0.44, 0.52), size=(30, 3))
    x = np.concatenate([xw, xr]).astype(np.float32)
    mat = np.concatenate([np.full(60, M.WATER), np.full(30, M.RIGID_HEAVY)]).astype(np.int32)
    bid = np.concatenate([np.zeros(60), np.ones(30)]).astype(np.int32)
    N = len(x)
    v = (rng.randn(N, 3) * 0.3 + np.array([1.5, -2.0, 0.8])).astype(np.float32)     # towards the +x wall and the floor
    C = (rng.randn(N, 3, 3) * 2.0).astype(np.float32); F = (np.eye(3)[None] + rng.randn(N, 3, 3) * 0.02).astype(np.float32)
    bnd = dict(type='cube', lower=(0.32, 0.32, 0.32), upper=(0.68, 0.68, 0.68), restitution=0.25, lock_dims=[2])
    S = R['sim'].MPMSimulator(dim=3, quality=n_grid / 64, gravity=(0.0, -10.0, 0.0), horizon=10, max_substeps_local=20, max_substeps_global=1000, ckpt_dest='cpu')
    S.setup_boundary(**bnd)
    rho = np.array([R['macros'].RHO[int(m)] for m in mat], dtype=np.float32)
    S.build(None, None, [], dict(x=x, used=np.ones(N), mat=mat, rho=rho, body_id=bid, bodies={'n': 2}))
    S.setframe(0, x, v, C, F, np.ones(N, np.int32))
    for f in range(n_sub):
        S.substep(f, True)
    out = read_frame(S, n_sub)
    return dict(n_grid=n_grid, n_sub=n_sub, x0=x, v0=v, C0=C, F0=F, mat=mat, body_id=bid, b_lower=bnd['lower'], b_upper=bnd['upper'], restitution=0.25,
                lock_dims=np.array([2]), **{'ref_' + k: a for k, a in out.items()})


# ---------------------------------------------------------------------------------------------------------------- scene 3b
def scene_latteart(R):
    """the LatteArt configuration (envs/latteart_env.py, agent_latteart.yaml) in miniature: AgentInjector with a locally-random Injector whose
    own boundary is a cylinder with y pinned (the radial clamp of CylinderBoundary.impose_x is hit by the actions), parked MILK injected into a
    COFFEE pool inside a cylinder boundary, gravity -20; 3 steps"""
    from fluidlab_b200 import macros as M
    rng = np.random.RandomState(205)
    n_grid, n_coffee, n_milk, flux, n_steps, T = 16, 110, 70, 2, 3, 20
    x = np.concatenate([np.tile(M.NOWHERE, (n_milk, 1)), rng.uniform((0.37, 0.36, 0.37), (0.63, 0.45, 0.63), size=(n_coffee, 3))]).astype(np.float32)
    used = np.concatenate([np.zeros(n_milk), np.ones(n_coffee)]).astype(np.int32)
    mat = np.concatenate([np.full(n_milk, M.MILK), np.full(n_coffee, M.COFFEE)]).astype(np.int32)
    N = len(x)
    bnd = dict(type='cylinder', xz_radius=0.2, xz_center=(0.5, 0.5), y_range=(0.34, 0.9))
    ebnd = dict(type='cylinder', xz_radius=0.12, xz_center=(0.5, 0.5), y_range=(0.55, 0.55))
    common = dict(max_substeps_local=T, max_substeps_global=1000, max_action_steps_global=20, ckpt_dest='cpu')
    np.random.seed(32)
    agent = R['agents'].AgentInjector(**common)
    agent.add_effector(type='Injector', params=dict(radius=0.0075, flux=flux, init_pos=(0.5, 0.5, 0.5), action_dim=3, inject_v=(0.0, -3.0, 0.0), action_scale_p=(1.0, 1.0, 1.0),
                                                     action_scale_v=(1.0, 1.0, 1.0), locally_random=True), mesh_cfg=None, boundary_cfg=ebnd)
    S = R['sim'].MPMSimulator(dim=3, quality=n_grid / 64, gravity=(0.0, -20.0, 0.0), horizon=10, max_substeps_local=T, max_substeps_global=1000, ckpt_dest='cpu')
    S.setup_boundary(**bnd)
    rho = np.array([R['macros'].RHO[int(m)] for m in mat], dtype=np.float32)
    S.build(agent, None, [], dict(x=x, used=used, mat=mat, rho=rho, body_id=np.zeros(N), bodies={'n': 1}))
    agent.build(S)
    inj = agent.effectors[0]
    actions = np.array([[0.05, 0.02, 0.03], [0.06, -0.01, 0.05], [-0.02, 0.0, 0.04]], dtype=np.float32)   # leaves the r = 0.12 cylinder: radial clamp + pinned y
    action_p = np.array([0.47, 0.7, 0.52], dtype=np.float32)
    agent.apply_action_p(action_p)
    for i in range(n_steps):
        S.step(actions[i])
    out = read_frame(S, S.cur_substep_local)
    return dict(n_grid=n_grid, n_steps=n_steps, T=T, flux=flux, x0=x, used0=used, mat=mat, actions=actions, action_p=action_p,
                init_state=np.asarray(inj.init_state, dtype=np.float64), random_vector=inj.random_vector.to_numpy(),
                ref_pose=np.asarray(inj.get_state(S.cur_substep_local), dtype=np.float64), **{'ref_' + k: a for k, a in out.items()})


# ---------------------------------------------------------------------------------------------------------------- scene 4
def scene_pouring(R):
    """the real AgentPouring (agents/agent_pouring.py): 6-DOF Rigid whose Dynamic mesh (meshes/dynamic.py) collides at grid AND particle
    level with friction + soft influence, plus the collector; elastic blob; 2 steps"""
    from conftest import box_sdf
    from fluidlab_b200 import macros as M
    rng = np.random.RandomState(204)
    n_grid, N, n_steps, T = 16, 110, 2, 20
    x = rng.uniform((0.40, 0.42, 0.40), (0.60, 0.58, 0.60), size=(N, 3)).astype(np.float32)
    mat = np.full(N, M.ELASTIC, dtype=np.int32)
    # box whose centre is NOT on a sample plane of the volume: a box sampled symmetrically has a zero finite-difference gradient on its
    # medial planes, where the reference's normal = g / |g|_eps is pure fp32 round-off (any two fp32 implementations disagree there)
    res, he, half, ctr = 32, 0.2, np.array([0.12, 0.05, 0.08]), np.array([0.0043, 0.0031, -0.0052])
    ax = np.linspace(-he, he, res)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing='ij')
    q = np.stack([np.abs(X - ctr[0]) - half[0], np.abs(Y - ctr[1]) - half[1], np.abs(Z - ctr[2]) - half[2]], -1)
    vox = (np.linalg.norm(np.maximum(q, 0), axis=-1) + np.minimum(q.max(-1), 0)).astype(np.float32)
    sc = (res - 1) / (2 * he)
    Tm = np.eye(4); Tm[0, 0] = Tm[1, 1] = Tm[2, 2] = sc; Tm[:3, 3] = sc * he
    SDF_REGISTRY['box.obj'] = dict(voxels=vox, T_mesh_to_voxels=Tm)
    bnd = dict(type='cube', lower=(0.25, 0.25, 0.25), upper=(0.75, 0.75, 0.75))
    ebnd = dict(type='cube', lower=(0.05, 0.05, 0.05), upper=(0.95, 0.95, 0.95))
    cbnd = dict(type='cube', lower=(0.0, 0.0, 0.0), upper=(1.0, 1.0, 0.585))
    common = dict(max_substeps_local=T, max_substeps_global=1000, max_action_steps_global=20, ckpt_dest='cpu')
    agent = R['agents'].AgentPouring(collector_boundary=cbnd, **common)
    agent.add_effector(type='Rigid', params=dict(init_pos=(0.5, 0.64, 0.5), init_euler=(0.0, 23.0, 5.0), action_dim=6, action_scale_p=(1.0,) * 6, action_scale_v=(1.0,) * 6),
                       mesh_cfg=dict(file='box.obj', material=M.STIRRER, softness=100.0, scale=(1.0, 0.9, 1.1), euler=(0.0, 10.0, 0.0)), boundary_cfg=ebnd)
    S = R['sim'].MPMSimulator(dim=3, quality=n_grid / 64, gravity=(0.0, -10.0, 0.0), horizon=10, max_substeps_local=T, max_substeps_global=1000, ckpt_dest='cpu')
    S.setup_boundary(**bnd)
    S.build(agent, None, [], dict(x=x, used=np.ones(N), mat=mat, rho=np.ones(N, np.float32), body_id=np.zeros(N), bodies={'n': 1}))
    agent.build(S)
    rigid = agent.effectors[0]
    actions = np.array([[0.004, -0.03, 0.002, 0.02, -0.03, 0.05], [-0.003, -0.03, 0.004, -0.04, 0.02, 0.03]], dtype=np.float32)
    action_p = np.array([0.5, 0.64, 0.5, 0, 0, 0], dtype=np.float32)
    agent.apply_action_p(action_p)
    for i in range(n_steps):
        S.step(actions[i])
    out = read_frame(S, S.cur_substep_local)
    return dict(n_grid=n_grid, n_steps=n_steps, T=T, x0=x, mat=mat, b_lower=bnd['lower'], b_upper=bnd['upper'], e_lower=ebnd['lower'], e_upper=ebnd['upper'],
                c_lower=cbnd['lower'], c_upper=cbnd['upper'], actions=actions, action_p=action_p, init_state=np.asarray(rigid.init_state, dtype=np.float64),
                vox=vox.astype(np.float32), T_final=np.asarray(rigid.mesh.T_mesh_to_voxels_np, dtype=np.float64), friction=float(rigid.mesh.friction), softness=100.0,
                ref_pose=np.asarray(rigid.get_state(S.cur_substep_local), dtype=np.float64), **{'ref_' + k: a for k, a in out.items()})


# ---------------------------------------------------------------------------------------------------------------- scene 5
def scene_icecream(R):
    """the real AgentIceCreamDynamic (BallInjector of plasto-elastic ICECREAM that stops at inject_till + Rigid sphere collider acting only
    above y = 0.25) and a Static mesh with dynamics colliding in grid_op (meshes/static.py); 3 steps"""
    from conftest import sphere_sdf, box_sdf
    from fluidlab_b200 import macros as M
    n_grid, N, n_steps, T, flux, inject_till = 16, 140, 3, 20, 4, 17
    x = np.tile(np.array(M.NOWHERE), (N, 1)).astype(np.float32)
    mat = np.full(N, M.ICECREAM, dtype=np.int32)
    vox, Tm = sphere_sdf(0.10, 0.2)
    SDF_REGISTRY['cone.obj'] = dict(voxels=vox, T_mesh_to_voxels=Tm)
    bv, bT = box_sdf((0.3, 0.05, 0.3), 0.4)
    SDF_REGISTRY['plate.obj'] = dict(voxels=bv, T_mesh_to_voxels=bT)
    cube = dict(type='cube', lower=(0.05, 0.05, 0.05), upper=(0.95, 0.95, 0.95))
    common = dict(max_substeps_local=T, max_substeps_global=1000, max_action_steps_global=20, ckpt_dest='cpu')
    np.random.seed(41)
    agent = R['agents'].AgentIceCreamDynamic(inject_till=inject_till, **common)
    agent.add_effector(type='BallInjector', params=dict(locally_random=True, radius=0.035, flux=flux, init_pos=(0.5, 0.62, 0.5), inject_v=(0.0, -0.4, 0.0), action_dim=3),
                       mesh_cfg=None, boundary_cfg=cube)
    agent.add_effector(type='Rigid', params=dict(init_pos=(0.5, 0.46, 0.5), action_dim=3), mesh_cfg=dict(file='cone.obj', material=M.CONE, softness=100.0), boundary_cfg=cube)
    statics = R['meshes'].Statics()
    statics.add_static(file='plate.obj', material=M.CUP, has_dynamics=True, pos=(0.5, 0.36, 0.5))
    S = R['sim'].MPMSimulator(dim=3, quality=n_grid / 64, gravity=(0.0, -10.0, 0.0), horizon=10, max_substeps_local=T, max_substeps_global=1000, ckpt_dest='cpu')
    S.setup_boundary(**cube)
    S.build(agent, None, statics, dict(x=x, used=np.zeros(N), mat=mat, rho=np.full(N, R['macros'].RHO[M.ICECREAM], np.float32), body_id=np.zeros(N), bodies={'n': 1}))
    agent.build(S)
    inj, rigid = agent.effectors
    actions = (np.array([[0.3, 0.2, -0.2], [-0.2, 0.4, 0.3], [0.1, -0.3, 0.2]]) * 0.02).astype(np.float32)
    action_p = np.array([0.5, 0.46, 0.5], dtype=np.float32)
    agent.apply_action_p(action_p)
    for i in range(n_steps):
        S.step(actions[i])
    out = read_frame(S, S.cur_substep_local)
    return dict(n_grid=n_grid, n_steps=n_steps, T=T, flux=flux, inject_till=inject_till, x0=x, mat=mat, lower=cube['lower'], upper=cube['upper'], actions=actions,
                action_p=action_p, random_vector=np.asarray(inj.random_vector_np, dtype=np.float64), vox=vox.astype(np.float32),
                T_rigid=np.asarray(rigid.mesh.T_mesh_to_voxels_np, dtype=np.float64), friction_rigid=float(rigid.mesh.friction),
                svox=bv.astype(np.float32), T_static=np.asarray(statics[0].T_mesh_to_voxels_np, dtype=np.float64), friction_static=float(statics[0].friction),
                ref_pose=np.asarray(rigid.get_state(S.cur_substep_local), dtype=np.float64), **{'ref_' + k: a for k, a in out.items()})


def main():
    R = load_reference()
    patch_mesh_io(R)
    for name, fn in (('multimat', scene_multimat), ('rigid_bodies', scene_rigid_bodies), ('locked', scene_locked), ('jetbot', scene_jetbot), ('jetbot_randv', lambda R_: scene_jetbot(R_, randomize_inject_v=True)), ('latteart', scene_latteart), ('pouring', scene_pouring),
                     ('icecream', scene_icecream)):
        d = fn(R)
        np.savez_compressed(os.path.join(HERE, f'reference_run_{name}.npz'), **d)
        print(name, os.path.getsize(os.path.join(HERE, f'reference_run_{name}.npz')), 'bytes')


if __name__ == '__main__':
    main()
