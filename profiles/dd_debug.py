import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from conftest import make_particles
from fluidlab_b200 import MPMSimulator, macros as M
def run_case(n_grid, N, sort_every, graphs, amp=0.2):
    rs = np.random.RandomState(0)
    x = rs.uniform((0.25, 0.30, 0.25), (0.75, 0.54, 0.75), size=(N, 3))
    P = make_particles(x, M.WATER, n_grid)
    s = MPMSimulator(dim=3, quality=n_grid/64, gravity=(0.0, -10.0, 0.0), horizon=100, max_substeps_local=50, max_substeps_global=100000, ckpt_dest='gpu', sort_every=sort_every)
    s.use_graphs = graphs
    s.build(None, None, [], P)
    v0 = (rs.randn(N, 3) * amp).astype(np.float32)
    u = rs.randn(N, 3).astype(np.float32)
    w = rs.randn(N, 3).astype(np.float32)
    base = s.get_state()
    def run(v):
        st = dict(base); st['v'] = v
        s.cur_substep_global = 0
        s.set_state(0, st)
        s.step(None)
        return s.get_state()['x'].astype(np.float64)
    s.enable_grad()
    xg = run(v0)
    s.reset_grad()
    z3, z9 = np.zeros((N, 3), np.float32), np.zeros((N, 3, 3), np.float32)
    s.set_grad(w, z3, z9, z9)
    s.step_grad(None)
    gv = s.get_grad(('v',))['v'].astype(np.float64)
    an = float((gv * u).sum())
    an_free = float((w.astype(np.float64) * u).sum()) * 10 * 2e-4   # ballistic part
    s.disable_grad()
    xn = run(v0)
    print(f'n_grid {n_grid} N {N} sort {sort_every} graphs {graphs}: an {an:.6f} (ballistic {an_free:.6f}) |x_grad - x_nograd| {np.abs(xg-xn).max():.2e}')
    for eps in (1e-1, 2e-2, 5e-3):
        lp = (run(v0 + eps * u) * w).sum(); lm = (run(v0 - eps * u) * w).sum()
        print(f'   eps {eps}: fd {(lp-lm)/(2*eps):.6f}')
run_case(32, 15625, 1, True)
run_case(64, 125000, 1, True)
run_case(128, 1000000, 1, True)
run_case(128, 1000000, 0, False)
