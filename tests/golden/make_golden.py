"""Generates tests/golden/*.npz with the CPU oracle (oracle/, fp32 arithmetic, one thread -> deterministic summation order;
gradients from the fp64 oracle).  The reference itself cannot run here (Taichi is not installable: PARITY UNPINNED), so these
vectors freeze the oracle's restatement of mpm_simulator.py at the commit that generated them.

    python tests/golden/make_golden.py
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import make_particles  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from fluidlab_b200 import macros as M  # noqa: E402


def multimat():
    """four cubes (water, elastic, ice-cream, viscous coffee) dropping on the floor of a cube boundary; 40 substeps."""
    rng = np.random.RandomState(101)
    n_grid, n_sub = 32, 40
    boxes = [((0.25, 0.22, 0.25), (0.45, 0.40, 0.45), M.WATER), ((0.55, 0.22, 0.25), (0.75, 0.40, 0.45), M.ELASTIC),
             ((0.25, 0.22, 0.55), (0.45, 0.40, 0.75), M.ICECREAM), ((0.55, 0.22, 0.55), (0.75, 0.40, 0.75), M.COFFEE_VIS)]
    xs, mats = [], []
    for lo, hi, m in boxes:
        xs.append(rng.uniform(lo, hi, size=(1500, 3))); mats.append(np.full(1500, m))
    x, mat = np.concatenate(xs).astype(np.float32), np.concatenate(mats).astype(np.int32)
    bnd = dict(type='cube', lower=(0.2, 0.2, 0.2), upper=(0.8, 0.8, 0.8))
    P = make_particles(x, mat, n_grid)
    orc.lib().orc_set_threads(1)
    o = orc.OracleSim(n_grid, P, gravity=(0, -10, 0), boundary=bnd, max_substeps_local=n_sub, precision=32)
    for f in range(n_sub):
        o.substep(f)
    fr = o.get_frame(n_sub)
    o64 = orc.OracleSim(n_grid, P, gravity=(0, -10, 0), boundary=bnd, max_substeps_local=n_sub, precision=64)
    for f in range(n_sub):
        o64.substep(f)
    f64 = o64.get_frame(n_sub)
    np.savez_compressed(os.path.join(HERE, 'multimat_n32_40sub.npz'), n_grid=n_grid, n_sub=n_sub, x0=x, mat=mat,
                        b_lower=bnd['lower'], b_upper=bnd['upper'], gravity=(0, -10, 0),
                        x=fr['x'].astype(np.float32), v=fr['v'].astype(np.float32), C=fr['C'].astype(np.float32), F=fr['F'].astype(np.float32),
                        x64=f64['x'], v64=f64['v'], F64=f64['F'])


def latte_mini():
    """LatteArt-like (envs/latteart_env.py): parked MILK + COFFEE pool + Injector; 3 steps over a T=20 ring, fwd + dLoss/dAction."""
    rng = np.random.RandomState(102)
    n_grid, n_steps, T, flux = 32, 3, 20, 2
    n_coffee, n_milk = 3000, 200
    x = np.concatenate([np.tile(M.NOWHERE, (n_milk, 1)), rng.uniform((0.35, 0.36, 0.35), (0.65, 0.45, 0.65), size=(n_coffee, 3))]).astype(np.float32)
    mat = np.concatenate([np.full(n_milk, M.MILK), np.full(n_coffee, M.COFFEE)]).astype(np.int32)
    used = np.concatenate([np.zeros(n_milk), np.ones(n_coffee)]).astype(np.int32)
    bnd = dict(type='cylinder', xz_radius=0.2, xz_center=(0.5, 0.5), y_range=(0.34, 0.9))
    ebnd = dict(type='cylinder', xz_radius=0.2, xz_center=(0.5, 0.5), y_range=(0.55, 0.55))
    rv = rng.uniform(size=(T, flux, 3)).astype(np.float32)
    actions = rng.uniform(-0.004, 0.004, size=(n_steps, 3)).astype(np.float32)
    action_p = np.array([0.47, 0.55, 0.52], dtype=np.float32)
    tgt = rng.uniform(0.4, 0.6, size=(n_steps,) + x.shape).astype(np.float32)
    P = make_particles(x, mat, n_grid, used=used)
    out = {}
    for prec in (32, 64):
        orc.lib().orc_set_threads(1)
        o = orc.OracleSim(n_grid, P, gravity=(0, -20, 0), boundary=bnd, max_substeps_local=T, precision=prec)
        o.add_effector(type=1, action_dim=3, boundary=ebnd, radius=0.0075, flux=flux, inject_v=(0, -3, 0), inject_p=(0, 0, 0), locally_random=True,
                       random_vector=rv, act_range=np.where(used == 0)[0], max_action_steps=n_steps + 1)
        o.enable_grad()
        o.set_effector_state(0, 0, np.array([0.5, 0.5, 0.5, 1, 0, 0, 0, 0.0]))
        o.apply_action_p(action_p)
        loss = 0.0
        for i in range(n_steps):
            o.step(actions[i]); loss += o.loss_value(o.cur_substep_local, M.MILK, 1.0, tgt[i])
        fr = o.get_frame(o.cur_substep_local)
        o.reset_grad()
        for i in range(n_steps - 1, -1, -1):
            o.loss_seed(o.cur_substep_local, M.MILK, 1.0, tgt[i]); o.step_grad(actions[i])
        o.apply_action_p_grad()
        out[prec] = (fr, loss, o.get_action_grad(n_steps))
    fr, loss32, grad32 = out[32]
    _, loss64, grad64 = out[64]
    np.savez_compressed(os.path.join(HERE, 'latte_mini_n32.npz'), n_grid=n_grid, n_steps=n_steps, T=T, flux=flux, x0=x, mat=mat, used0=used,
                        random_vector=rv, actions=actions, action_p=action_p, tgt=tgt,
                        x=fr['x'].astype(np.float32), v=fr['v'].astype(np.float32), F=fr['F'].astype(np.float32), used=fr['used'],
                        loss32=loss32, loss64=loss64, grad32=grad32, grad64=grad64)


def rigid_pour_scene():
    """everything a Pouring / Gathering scene exercises at once: WATER pool + one MAT_RIGID cuboid (shape matching, MPM:449-505) + an
    elastic blob, a 6-DOF Rigid box collider at grid AND particle level (AgentPouring, 'both') and the collector."""
    from conftest import box_sdf
    rng = np.random.RandomState(103)
    n_grid = 32
    xw = rng.uniform((0.38, 0.30, 0.38), (0.62, 0.40, 0.62), size=(2500, 3))
    xr = rng.uniform((0.42, 0.42, 0.44), (0.52, 0.48, 0.52), size=(700, 3))
    xe = rng.uniform((0.53, 0.42, 0.50), (0.60, 0.50, 0.58), size=(500, 3))
    x = np.concatenate([xw, xr, xe]).astype(np.float32)
    mat = np.concatenate([np.full(len(xw), M.WATER), np.full(len(xr), M.RIGID), np.full(len(xe), M.ELASTIC)]).astype(np.int32)
    bid = np.concatenate([np.zeros(len(xw)), np.ones(len(xr)), np.full(len(xe), 2)]).astype(np.int32)
    vox, Tm = box_sdf(np.array([0.10, 0.04, 0.07]), 0.2)
    cfg = dict(n_grid=n_grid, n_steps=2, T=20, x0=x, mat=mat, body_id=bid,
               bnd=dict(type='cube', lower=(0.25, 0.25, 0.25), upper=(0.75, 0.75, 0.75)),
               ebnd=dict(type='cube', lower=(0.05, 0.05, 0.05), upper=(0.95, 0.95, 0.95)),
               cbnd=dict(type='cube', lower=(0.0, 0.0, 0.0), upper=(1.0, 1.0, 0.605)),
               vox=vox.astype(np.float32), Tm=Tm.astype(np.float64), friction=float(M.FRICTION[M.STIRRER]), softness=100.0,
               init=np.array([0.5, 0.56, 0.5, np.cos(0.15), 0.0, np.sin(0.15), 0.0, 0.0]),
               actions=np.array([[0.004, -0.035, 0.002, 0.02, -0.03, 0.05], [-0.003, -0.035, 0.004, -0.04, 0.02, 0.03]], dtype=np.float32),
               action_p=np.array([0.5, 0.56, 0.5, 0, 0, 0], dtype=np.float32),
               tgt=rng.uniform(0.4, 0.6, size=(2,) + x.shape).astype(np.float32))
    return cfg


def run_rigid_pour(cfg, prec, threads=1):
    P = make_particles(cfg['x0'], cfg['mat'], int(cfg['n_grid']))
    N = len(cfg['x0'])
    orc.lib().orc_set_threads(threads)
    o = orc.OracleSim(int(cfg['n_grid']), P, gravity=(0, -10, 0), boundary=cfg['bnd'], max_substeps_local=int(cfg['T']), precision=prec)
    o.set_bodies(cfg['body_id'], 3)
    o.add_effector(type=0, action_dim=6, scale_v=(1,) * 6, boundary=cfg['ebnd'], max_action_steps=int(cfg['n_steps']) + 1)
    o.set_rigid_mesh(cfg['vox'], cfg['Tm'], friction=cfg['friction'], softness=cfg['softness'], collide_type='both')
    o.set_collector(cfg['cbnd'], mat=-1)
    o.enable_grad()
    o.set_frame(0, P['x'], np.zeros((N, 3)), np.zeros((N, 3, 3)), np.tile(np.eye(3), (N, 1, 1)), P['used'])
    o.set_effector_state(0, 0, cfg['init']); o.apply_action_p(cfg['action_p'])
    loss = 0.0
    n_steps = int(cfg['n_steps'])
    for i in range(n_steps):
        o.step(cfg['actions'][i]); loss += o.loss_value(o.cur_substep_local, M.WATER, 1.0, cfg['tgt'][i])
    fr = o.get_frame(o.cur_substep_local)
    o.reset_grad()
    for i in range(n_steps - 1, -1, -1):
        o.loss_seed(o.cur_substep_local, M.WATER, 1.0, cfg['tgt'][i]); o.step_grad(cfg['actions'][i])
    o.apply_action_p_grad()
    return fr, loss, o.get_action_grad(n_steps)


def rigid_pour():
    cfg = rigid_pour_scene()
    fr, loss32, grad32 = run_rigid_pour(cfg, 32)
    f64, loss64, grad64 = run_rigid_pour(cfg, 64)
    flat = {k: v for k, v in cfg.items() if not isinstance(v, dict)}
    for name in ('bnd', 'ebnd', 'cbnd'):
        flat[name + '_lower'] = cfg[name]['lower']; flat[name + '_upper'] = cfg[name]['upper']
    np.savez_compressed(os.path.join(HERE, 'rigid_pour_n32.npz'), **flat,
                        x=fr['x'].astype(np.float32), v=fr['v'].astype(np.float32), F=fr['F'].astype(np.float32), used=fr['used'],
                        x64=f64['x'], v64=f64['v'], F64=f64['F'], used64=f64['used'],
                        loss32=loss32, loss64=loss64, grad32=grad32, grad64=grad64)


def load_rigid_pour(path):
    d = dict(np.load(path))
    for name in ('bnd', 'ebnd', 'cbnd'):
        d[name] = dict(type='cube', lower=tuple(d.pop(name + '_lower')), upper=tuple(d.pop(name + '_upper')))
    d['friction'] = float(d['friction']); d['softness'] = float(d['softness'])
    return d


def c1_scene():
    """BASELINE.json configs[0] / SURVEY.md §8d C1: LatteArt-v0 default scene (envs/latteart_env.py:28-74, agent_latteart.yaml) with the
    host classes of fluidlab_b200 (no CUDA needed): 60,000 parked MILK + 55,480 COFFEE, cylinder boundary, Injector with flux 2, the
    scripted pour.  The injector's random table is the first draw after np.random.seed(0), exactly as in the GPU test."""
    from fluidlab_b200.bodies import Bodies
    from test_gpu_parity import latteart_demo_actions
    b = Bodies(dim=3, particle_density=1e6)
    b.add_body(type='nowhere', n_particles=60000, material=M.MILK)
    b.add_body(type='cylinder', center=(0.5, 0.55, 0.5), height=0.1, radius=0.42, material=M.COFFEE)
    Pb = b.get()
    rv = np.random.RandomState(0).uniform(size=(50, 2, 3)).astype(np.float32)
    acts, init_p = latteart_demo_actions()
    return Pb, rv, acts, init_p


def run_c1(prec, n_steps=10):
    Pb, rv, acts, init_p = c1_scene()
    P = make_particles(Pb['x'], Pb['mat'], 64, used=Pb['used'].astype(np.int32))
    bnd = dict(type='cylinder', xz_radius=0.42, xz_center=(0.5, 0.5), y_range=(0.5, 0.95))
    orc.lib().orc_set_threads(orc.lib().orc_get_max_threads())
    o = orc.OracleSim(64, P, gravity=(0, -20, 0), boundary=bnd, precision=prec, max_substeps_local=50)
    o.add_effector(type=1, action_dim=3, boundary=dict(type='cylinder', xz_radius=0.42, xz_center=(0.5, 0.5), y_range=(0.65, 0.65)), radius=0.0075, flux=2,
                   inject_v=(0, -3, 0), inject_p=(0, 0, 0), locally_random=True, random_vector=rv, act_range=np.where(P['used'] == 0)[0], max_action_steps=331)
    o.set_effector_state(0, 0, np.array([0.5, 0.5, 0.5, 1, 0, 0, 0, 0.0])); o.apply_action_p(init_p)
    for i in range(n_steps):
        o.step(acts[i])
    return o.get_frame(o.cur_substep_local), P


def c1_sample_ids(used):
    """first 48 used COFFEE slots + the 16 most recently injected MILK slots (slot order [milk..., coffee...])"""
    coffee = np.where(used[60000:] != 0)[0][:48] + 60000
    milk = np.where(used[:60000] != 0)[0][-16:]
    return np.concatenate([milk, coffee])


def c1_latteart():
    """SURVEY.md §8c substitute pin 4: C1 after 100 substeps — 64 particles verbatim + checksums over all used particles."""
    fr, P = run_c1(32)
    f64, _ = run_c1(64)
    ids = c1_sample_ids(fr['used'])
    act = fr['used'] != 0
    cs = lambda a: np.array([a[act].astype(np.float64).sum(), np.abs(a[act].astype(np.float64)).sum(), (a[act].astype(np.float64) ** 2).sum()])
    np.savez_compressed(os.path.join(HERE, 'c1_latteart_100sub.npz'), ids=ids, n_used=int(act.sum()),
                        x=fr['x'][ids].astype(np.float32), v=fr['v'][ids].astype(np.float32), F=fr['F'][ids].astype(np.float32),
                        x64=f64['x'][ids], v64=f64['v'][ids], F64=f64['F'][ids],
                        cs_x=cs(f64['x']), cs_v=cs(f64['v']), cs_F=cs(f64['F']), cs_x32=cs(fr['x']), cs_v32=cs(fr['v']), cs_F32=cs(fr['F']))


if __name__ == '__main__':
    multimat(); latte_mini(); rigid_pour(); c1_latteart()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)))
