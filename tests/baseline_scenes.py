"""BASELINE.json configs[2] (C3) and configs[3] (C4) as TaichiEnv scenes for `bench.py --config C3|C4` — the same scenes the parity tests
test_c3_latteart_two_material_fwd_bwd_full_size / c4_case (tests/test_gpu_parity.py) compare with the oracle, without the oracle legs.
Scene sources: envs/latteart_env.py:28-74 + agent_latteart.yaml (C3), envs/icecreamdynamic_env.py + agent_icecreamdynamic.yaml:26-37 (C4)."""
import numpy as np


def latteart_demo_actions(horizon_action=250):
    """scripted pour of envs/latteart_env.py:113-140 (demo_policy): returns (actions_v [T,3], action_p [3])."""
    init_p = np.array([0.15, 0.65, 0.5]); x_range = 0.7
    cur = init_p.copy(); amp = np.array([0.15, 0.25]); acts = np.zeros((horizon_action, 3))
    for i in range(horizon_action):
        t = i + 1
        tx = init_p[0] + t / horizon_action * x_range
        rad = t / horizon_action * (np.pi * 2) * 3
        a = amp[1] - np.abs((t * 2 / horizon_action) - 1) * (amp[1] - amp[0])
        tp = np.array([tx, init_p[1], np.sin(rad) * a + 0.5])
        acts[i] = tp - cur; cur += acts[i]
    return acts, init_p


def c3_env(n_steps, T=50, device_kw=None, scale=1.0):
    """C3: 262,144 slots = 62,144 parked MILK + 200,000 COFFEE in the cylinder r = 0.42, 128^3, gravity -20, Injector with flux 8, LatteArtLoss.
    Returns (env, actions [n_steps, 3] float32, action_p [3] float32)."""
    from conftest import make_particles
    from fluidlab_b200 import TaichiEnv, LatteArtLoss
    from fluidlab_b200 import macros as M
    n_grid, n_coffee, n_milk, flux = 128, int(200_000 * scale), int(62_144 * scale), 8   # scale < 1: script checks on the CPU execution-model shim only
    rs = np.random.RandomState(0)
    acc = []
    while sum(len(a) for a in acc) < n_coffee:
        c = rs.uniform((0.08, 0.50, 0.08), (0.92, 0.60, 0.92), size=(100_000, 3))
        acc.append(c[(c[:, 0] - 0.5) ** 2 + (c[:, 2] - 0.5) ** 2 <= 0.42 ** 2])
    xc = np.concatenate(acc)[:n_coffee]
    x = np.concatenate([np.tile(M.NOWHERE, (n_milk, 1)), xc])
    mat = np.concatenate([np.full(n_milk, M.MILK), np.full(n_coffee, M.COFFEE)])
    used = np.concatenate([np.zeros(n_milk), np.ones(n_coffee)]).astype(np.int32)
    P = make_particles(x, mat, n_grid, used=used)
    ebnd = dict(type='cylinder', xz_radius=0.42, xz_center=(0.5, 0.5), y_range=(0.65, 0.65))
    bnd = dict(type='cylinder', xz_radius=0.42, xz_center=(0.5, 0.5), y_range=(0.5, 0.95))
    env = TaichiEnv(quality=2, max_substeps_local=T, gravity=(0.0, -20.0, 0.0), horizon=n_steps, **(device_kw or {}))
    np.random.seed(0)
    env.setup_agent(dict(type='AgentInjector', effectors=[dict(type='Injector', params=dict(
        radius=0.0075, flux=flux, init_pos=(0.5, 0.5, 0.5), action_dim=3, inject_v=(0.0, -3.0, 0.0), locally_random=True), boundary=ebnd)]))
    env.particle_bodies.get = lambda: P
    env.setup_boundary(**bnd)
    tgt = [rs.uniform(0.3, 0.7, size=x.shape).astype(np.float32) for _ in range(n_steps)]
    env.setup_loss(loss_cls=LatteArtLoss, type='diff', target=tgt, weights={'chamfer': 1.0})
    env.build()
    acts, init_p = latteart_demo_actions()
    return env, acts[:n_steps].astype(np.float32), init_p.astype(np.float32)


def c4_env(n_steps=1, n_grid=192, n_each=1_000_000, T=10, cone_y=0.75, device_kw=None):
    """C4: 1M ELASTIC + 1M ICECREAM (plasto-elastic) particles on a 192^3 grid with the soft cone collider of agent_icecreamdynamic.yaml:26-37 acting at particle
    level, IceCreamDynamicLoss.  `cone_y`: height of the cone's origin — at 0.75 it hovers above the blocks (every particle still evaluates the collider's
    SDF; with the reference's fixed dt a contact perturbation at 192^3 grows 5x per substep, see test_c4_cone_collider_fwd_bwd_at_the_full_particle_count),
    at 0.56 its tip dips into the ice cream as in the parity test.  Returns (env, actions, action_p)."""
    from conftest import make_particles, cone_sdf
    from fluidlab_b200 import TaichiEnv, IceCreamDynamicLoss
    from fluidlab_b200 import macros as M
    rs = np.random.RandomState(0)
    xa = rs.uniform((0.20, 0.30, 0.30), (0.45, 0.55, 0.70), size=(n_each, 3))
    xb = rs.uniform((0.55, 0.30, 0.30), (0.80, 0.55, 0.70), size=(n_each, 3))
    x = np.concatenate([xa, xb]); mat = np.concatenate([np.full(len(xa), M.ELASTIC), np.full(len(xb), M.ICECREAM)])
    N = len(x)
    P = make_particles(x, mat, n_grid)
    vox, Tm = cone_sdf(0.10, 0.22, 0.2)
    cube = dict(type='cube', lower=(0.05, 0.05, 0.05), upper=(0.95, 0.95, 0.95))
    env = TaichiEnv(quality=n_grid / 64, max_substeps_local=T, gravity=(0.0, -10.0, 0.0), horizon=n_steps, **(device_kw or {}))
    init_pos = (0.66, cone_y, 0.5)
    env.setup_agent(dict(type='AgentRigid', params=dict(collide_type='particle'), effectors=[dict(
        type='Rigid', params=dict(init_pos=init_pos, init_euler=(0.0, 0.0, 0.0), action_dim=3, action_scale_p=(1.0, 1.0, 1.0), action_scale_v=(1.0, 1.0, 1.0)),
        mesh=dict(file='cone_tip.obj', scale=(0.726, 0.726, 0.726), euler=(-90.0, 0.0, 30.0), material='CONE', softness=100.0, sdf=dict(voxels=vox, T_mesh_to_voxels=Tm)),
        boundary=cube)]))
    env.setup_boundary(**cube)
    env.particle_bodies.get = lambda: P
    tgt = [(x + rs.randn(N, 3) * 0.01).astype(np.float32) for _ in range(n_steps)]
    env.setup_loss(loss_cls=IceCreamDynamicLoss, type='default', target=tgt, weights={'chamfer': 1.0})
    env.build()
    actions = np.tile(np.array([[0.2, -0.9, 0.1]], dtype=np.float32) * 0.002, (n_steps, 1))
    return env, actions, np.array(init_pos, dtype=np.float32)
