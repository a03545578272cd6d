"""Per-phase device times of the forward substep on the C2 workload (1M water particles, 128^3), CUDA events around batches of
launches.  Used for A/B runs of kernel variants: FMPM_LIB=<variant.so> python profiles/phase_times.py"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import make_particles          # noqa: E402
from fluidlab_b200 import MPMSimulator, macros as M   # noqa: E402

N = int(os.environ.get('PT_N', 1_000_000))
rs = np.random.RandomState(0)
x = rs.uniform((0.25, 0.30, 0.25), (0.75, 0.54, 0.75), size=(N, 3))
P = make_particles(x, M.WATER, 128)
s = MPMSimulator(dim=3, quality=2, gravity=(0.0, -10.0, 0.0), horizon=1000, max_substeps_local=50, max_substeps_global=10 ** 7, ckpt_dest='gpu', sort_every=2)
s.build(None, None, [], P)
for _ in range(8):   # settle into a realistic, sorted state
    s.step(None)
f = s.cur_substep_local
s.sort_frame(f)
ev = lambda: torch.cuda.Event(enable_timing=True)
out = {}
REP = 20
for name, args in (('p2g', (0,)), ('grid_op', (0,)), ('g2p', ())):
    ts = []
    for it in range(REP + 3):
        # restore the between-phases state: accumulators hold p2g(f) before grid_op / g2p runs
        s.phase('clear_grid', f); s.phase('p2g', f, 0)
        if name != 'p2g':
            s.phase('grid_op', f, 0)
        if name == 'p2g':
            s.phase('clear_grid', f)
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        s.phase(name, f, *args)
        e1.record(); torch.cuda.synchronize()
        if it >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
        if name == 'p2g':
            s.phase('grid_op', f, 1)
        else:
            s.phase('clear_grid', f)
    out[name] = float(np.median(ts))
if os.environ.get('PT_BWD'):
    # backward phases of frame f (stored-grid mode is bypassed: the phase entry points use the scratch grids)
    s.enable_grad(); s.cur_substep_global = f
    s._ensure_grad_buffers()
    s.reset_grad()
    g = torch.randn((2, 4, N, 4), device=s.device) * 0.1
    s._ga.copy_(g); s._gf.normal_(0, 0.1); s._gf8.normal_(0, 0.1)
    s.phase('clear_grid', f); s.phase('p2g', f, 0); s.phase('grid_op', f, 0)   # forward grids + block flags of frame f stay in place
    for name in ('g2p_grad_scatter', 'grid_op_grad', 'particle_grad'):
        ts = []
        for it in range(REP + 3):
            torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            s.phase(name, f, *((0,) if name == 'grid_op_grad' else ()))
            e1.record(); torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3)
        out[name] = float(np.median(ts))
    s.disable_grad()
# the fused g2p(f) + p2g(f+1) kernel in isolation (PT_FUSED=1): grid_v of frame f in place, accumulator clear
if os.environ.get('PT_FUSED'):
    ts = []
    for it in range(REP + 3):
        s.phase('clear_grid', f); s.phase('p2g', f, 0); s.phase('grid_op', f, 1)
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        s.phase('g2p2g', f)
        e1.record(); torch.cuda.synchronize()
        if it >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
        s.phase('grid_op', f + 1, 1)
    out['g2p2g'] = float(np.median(ts))
    s.fuse_g2p2g = True     # and the whole-step figure below takes the fused path
# whole steps through the CUDA-graph path
for _ in range(4):
    s.step(None)
torch.cuda.synchronize()
e0, e1 = ev(), ev(); e0.record()
K = 40
for _ in range(K):
    s.step(None)
e1.record(); torch.cuda.synchronize()
out['substep_avg'] = e0.elapsed_time(e1) * 1e3 / (K * 10)
print(os.environ.get('FMPM_LIB', 'default'), ' '.join(f'{k} {v:.1f}us' for k, v in out.items()), f"-> {1e6 / out['substep_avg']:.0f} substeps/s")
