"""A/B of the forward paths of fmpm_substeps_fused on the C2 workload (1M water particles, 128^3): whole-step throughput through MPMSimulator.step
(CUDA-graph replay), one line per FMPM_FWD_MASK value (0 round-1 grid_op + k_g2p2g, 1 k_fwd, 3 + liquid, 5 + inlined grid_op, 7 all).
    python profiles/fwd_ab.py [masks...] > gpurun_out/fwd_ab.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fluidlab_b200 import MPMSimulator  # noqa: E402

masks = [int(a) for a in sys.argv[1:]] or [-1, 0, 1, 3, 5, 7]
K, W = int(os.environ.get('AB_STEPS', 60)), 8
out = []
parts = bench.workload_particles(bench.N_PARTICLES)
for mask in masks:
    sim = MPMSimulator(dim=3, quality=bench.QUALITY, gravity=bench.GRAVITY, horizon=4000, max_substeps_local=50, max_substeps_global=10 ** 7, ckpt_dest='gpu',
                       sort_every=int(os.environ.get('AB_SORT', 4)))
    sim.build(None, None, [], parts)
    sim.fuse_g2p2g = mask >= 0           # -1: the plain p2g / grid_op / g2p substeps
    if mask >= 0:
        sim._ck(sim._lib.fmpm_set_fwd_mask(sim._h, mask), 'mask')
    for _ in range(W):
        sim.step(None)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        sim.step(None)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    x = sim.get_x()
    out.append(dict(mask=mask, path=int(sim._lib.fmpm_fwd_path(sim._h)) if mask >= 0 else None, substeps_per_s=K * 10 / (ms * 1e-3), us_per_substep=ms * 1e3 / (K * 10),
                    x_checksum=float(np.abs(x.astype(np.float64)).sum()), finite=bool(np.isfinite(x).all())))
    print(json.dumps(out[-1]), flush=True)
    del sim
    torch.cuda.empty_cache()
