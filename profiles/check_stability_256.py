"""Diagnostic: is the reference's fixed dt=2e-4 stable for water at 256^3 (quality=4)?  SURVEY.md §8d C5 warns c*dt/dx ~ 0.85."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fluidlab_b200 import MPMSimulator
import bench
q = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = 64 * q; dx = 1.0 / n
lo = ((32 - 0.5) * dx, 0.30, 0.36); hi = ((56 - 0.5) * dx, 0.30 + 72 * dx, 0.36 + 72 * dx)
parts = bench.workload_particles(1_000_000, seed=0, lo=lo, hi=hi)
sim = MPMSimulator(dim=3, quality=q, gravity=(0, -10, 0), horizon=10 ** 5, max_substeps_local=50, max_substeps_global=10 ** 7, ckpt_dest='gpu')
sim.build(None, None, [], parts)
for s in range(80):
    sim.step(None)
    if s % 10 == 9:
        st = sim.readframe_torch(sim.cur_substep_local, ('x', 'v'))
        print((s + 1) * 10, 'finite', bool(torch.isfinite(st['x']).all()), 'max|v|', float(st['v'].abs().max()), flush=True)
