"""A/B of the backward path on the C2 workload (1M water particles, 128^3): a forward + backward pass over a whole-trajectory ring (no chunk re-simulation),
forward and backward device time separately.   FMPM_LIB=<variant.so> python profiles/bwd_ab.py >> gpurun_out/bwd_ab.jsonl"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import fluidlab_b200  # noqa: E402
from fluidlab_b200 import MPMSimulator  # noqa: E402

nfb = int(os.environ.get('AB_STEPS', 8))
NP = int(os.environ.get('AB_N', bench.N_PARTICLES))
parts = bench.workload_particles(NP)
sim = MPMSimulator(dim=3, quality=bench.QUALITY, gravity=bench.GRAVITY, horizon=400, max_substeps_local=(nfb + 1) * 10, max_substeps_global=10 ** 7, ckpt_dest='gpu', sort_every=4)
sim.build(None, None, [], parts)
sim.fuse_g2p2g = True
init = sim.get_state()
tgt = torch.zeros((NP, 3), dtype=torch.float32, device=sim.device) + 0.5
mask = sim.material_row_mask(fluidlab_b200.macros.WATER)
ev = lambda: torch.cuda.Event(enable_timing=True)


def one_pass():
    sim.set_state(0, init); sim.enable_grad()
    a, m, b = ev(), ev(), ev()
    a.record()
    for _ in range(nfb):
        sim.step(None)
    sim.reset_grad(); sim.add_x_grad_chamfer(tgt, mask, 1.0)
    m.record()
    for _ in range(nfb):
        sim.step_grad(None)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(m), m.elapsed_time(b)


one_pass()
runs = [one_pass() for _ in range(4)]
fwd, bwd = float(np.median([r[0] for r in runs])), float(np.median([r[1] for r in runs]))
g = sim.get_grad(('x',))['x']
print(json.dumps(dict(lib=os.path.basename(os.environ.get('FMPM_LIB', 'default')), fwd_us_per_substep=fwd * 1e3 / (nfb * 10), bwd_us_per_substep=bwd * 1e3 / (nfb * 10),
                      pairs_per_s=nfb * 10 / ((fwd + bwd) * 1e-3), grad_checksum=float(np.abs(g.astype(np.float64)).sum()), finite=bool(np.isfinite(g).all()))))
