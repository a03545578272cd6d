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


if __name__ == '__main__':
    multimat(); latte_mini()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)))
