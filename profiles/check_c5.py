"""How long does BASELINE.json configs[4] (C5: 8M water particles, 256^3, from rest) stay finite with the reference's fixed dt = 2e-4?
One GPU, the default forward path; prints the first step whose positions are not finite (c dt / dx = 0.85 at 256^3).
    python profiles/check_c5.py [n_steps] > gpurun_out/check_c5.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fluidlab_b200 import MPMSimulator  # noqa: E402

n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
q, _, parts, _ = bench.c5_shard(1, 0)
out = {}
for fuse in (True, False):
    sim = MPMSimulator(dim=3, quality=q, gravity=bench.GRAVITY, horizon=1000, max_substeps_local=20, max_substeps_global=10 ** 7, ckpt_dest='gpu', sort_every=4)
    sim.build(None, None, [], parts)
    sim.fuse_g2p2g = fuse
    first_bad, vmax = None, []
    for s in range(n_steps):
        sim.step(None)
        f = sim.cur_substep_local
        xs, alive = sim.slab_positions(f)
        v = sim._pa[f, 1, :, :3]
        ok = bool(torch.isfinite(xs[alive]).all().item())
        vmax.append(float(v[alive].abs().max().item()) if ok else float('nan'))
        if not ok:
            first_bad = s
            break
    out['fused' if fuse else 'plain'] = dict(first_non_finite_step=first_bad, steps_run=s + 1, vmax_by_step=vmax[::5])
    del sim
    torch.cuda.empty_cache()
print(json.dumps(out))
