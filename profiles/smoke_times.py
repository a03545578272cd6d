"""Per-phase timings of the smoke solver (csrc/fsmk_smoke.cu) on one GPU: CUDA events around each phase-level entry point of
include/fluidsmoke.h, forward and adjoint, at the circulation env's size (128^3, band 60 < j < 68) for 50 and 500 Jacobi sweeps.
NOT YET RUN ON A B200 (written after round 1's GPU budget was spent); first thing to run next round:

    gpurun --timeout 600 -- 'python profiles/smoke_times.py > gpurun_out/smoke_times.json'

Prints one JSON object: ms per phase, per step and the algorithmic HBM bytes of the dense passes."""
import ctypes as C
import json
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build(res, iters, dev, lower_y=60, higher_y=68):
    from fluidlab_b200 import smoke as smoke_mod, meshes
    T = 100
    z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
    air = types.SimpleNamespace(pos=z(T + 1, 3), quat=z(T + 1, 4), s=z(T + 1), r=z(T + 1), gpos=z(T + 1, 3), gquat=z(T + 1, 4), gs=z(T + 1), gr=z(T + 1),
                                inject_v=np.array([-0.3, 0.0, 1.0]))
    air.quat[:, 0] = 1; air.pos[:] = torch.tensor([0.55, 0.5, 0.27], device=dev); air.s[:] = 2000.0; air.r[:] = 2.0
    agent = types.SimpleNamespace(aircon=air)
    stream = (lambda: None) if dev.type != 'cuda' else (lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    sim = types.SimpleNamespace(max_steps_local=10, agent=agent, device=dev, statics=meshes.Statics(), _stream=stream)
    sf = smoke_mod.SmokeField(dim=3, ckpt_dest='gpu', res=res, dt=0.03, solver_iters=iters, q_dim=1)
    sf.lower_y, sf.higher_y = lower_y, higher_y
    sf.build(sim, agent)
    return sf


def timed(fn, reps, event_cls, sync):
    fn(); sync()
    a, b = event_cls(enable_timing=True), event_cls(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); sync()
    return a.elapsed_time(b) / reps


def run(res=128, dev=None, reps=20, band=(60, 68), iters_list=(50, 500), event_cls=None, sync=None):
    dev = torch.device('cuda', 0) if dev is None else dev
    event_cls = event_cls or torch.cuda.Event
    sync = sync or torch.cuda.synchronize
    out = {'res': res, 'band_cells': (band[1] - band[0] - 1) * res * res, 'dense_bytes_per_step': res ** 3 * (16 + 16 + 4 + 4 + 1 + 16), 'runs': []}
    for iters in iters_list:
        sf = build(res, iters, dev, *band)
        L, h, st = sf._lib, sf._h, sf._stream
        for s in range(3):          # a few steps so that the state is not trivial
            sf.step(s, 10 * s)
        sf._ensure_grad_buffers()
        sf._gv[4].normal_(); sf._gq[4].normal_(); sf._gp[4].normal_()
        ph = {}
        ph['free_space'] = timed(lambda: sf._ck(L.fsmk_free_space(h, 3, st()), 'free_space'), reps, event_cls, sync)
        ph['advect'] = timed(lambda: sf._ck(L.fsmk_advect(h, 3, 30, st()), 'advect'), reps, event_cls, sync)
        ph['divergence'] = timed(lambda: sf._ck(L.fsmk_divergence(h, 3, st()), 'divergence'), reps, event_cls, sync)
        ph['pressure'] = timed(lambda: sf._ck(L.fsmk_pressure(h, 3, st()), 'pressure'), reps, event_cls, sync)
        ph['project'] = timed(lambda: sf._ck(L.fsmk_project(h, 3, st()), 'project'), reps, event_cls, sync)
        ph['step'] = timed(lambda: sf.step(3, 30), reps, event_cls, sync)
        ph['project_grad'] = timed(lambda: sf._ck(L.fsmk_project_grad(h, 3, st()), 'project_grad'), reps, event_cls, sync)
        ph['pressure_grad'] = timed(lambda: sf._ck(L.fsmk_pressure_grad(h, 3, st()), 'pressure_grad'), reps, event_cls, sync)
        ph['divergence_grad'] = timed(lambda: sf._ck(L.fsmk_divergence_grad(h, 3, st()), 'divergence_grad'), reps, event_cls, sync)
        ph['advect_grad'] = timed(lambda: sf._ck(L.fsmk_advect_grad(h, 3, 30, st()), 'advect_grad'), reps, event_cls, sync)
        ph['step_grad'] = timed(lambda: sf.step_grad(3, 30), reps, event_cls, sync)
        out['runs'].append({'solver_iters': iters, 'ms': ph, 'jacobi_launches': (iters + 7) // 8, 'us_per_sweep': 1e3 * ph['pressure'] / max(iters, 1)})
    return out


if __name__ == '__main__':
    print(json.dumps(run()))
