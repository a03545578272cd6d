"""LatteArt-v0 (BASELINE.json configs[0], C1: 115,480 slots, 64^3, injector agent, demo policy) rolled out through TaichiEnv.step, forward only and
forward + backward, with MPMSimulator.fuse_g2p2g off and on: steps/s and substeps/s.  C1 is the reference's own CPU-runnable case; at this size
the substep is launch- and latency-bound rather than HBM-bound, so the fused path's fewer launches and bytes are what is being measured.
NOT YET RUN ON A B200 (written after round 1's GPU budget was spent):

    gpurun --timeout 600 -- 'python profiles/c1_rollout_times.py > gpurun_out/c1_rollout.json'
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))


def build_env(fuse, device=None, n_milk=60000, quality=1, T=50):
    from fluidlab_b200 import TaichiEnv, macros as M
    env = TaichiEnv(dim=3, quality=quality, particle_density=1e6, max_substeps_local=T, gravity=(0.0, -20.0, 0.0), horizon=330, device=device,
                    ckpt_dest='gpu' if device is None else 'cpu')
    np.random.seed(0)
    env.setup_agent(dict(type='AgentInjector', effectors=[dict(type='Injector', params=dict(
        radius=0.0075, flux=2, init_pos=(0.5, 0.5, 0.5), action_dim=3, inject_v=(0.0, -3.0, 0.0), action_scale_p=(1.0, 1.0, 1.0),
        action_scale_v=(1.0, 1.0, 1.0), locally_random=True), boundary=dict(type='cylinder', xz_radius=0.42, xz_center=(0.5, 0.5), y_range=(0.65, 0.65)))]))
    env.add_body(type='nowhere', n_particles=n_milk, material=M.MILK)
    env.add_body(type='cylinder', center=(0.5, 0.55, 0.5), height=0.1, radius=0.42, material=M.COFFEE)
    env.setup_boundary(type='cylinder', xz_radius=0.42, xz_center=(0.5, 0.5), y_range=(0.5, 0.95))
    env.build()
    env.simulator.fuse_g2p2g = fuse
    if device is not None:
        env.simulator.use_graphs = False
    return env


def run(device=None, n_steps=30, reps=3, sync=None, **kw):
    from test_gpu_parity import latteart_demo_actions
    sync = sync or torch.cuda.synchronize
    acts, init_p = latteart_demo_actions()
    out = {}
    for fuse in (False, True):
        env = build_env(fuse, device, **kw)
        st0 = env.get_state()['state']
        ts = []
        for _ in range(reps + 1):
            env.set_state(st0, grad_enabled=False)
            env.apply_agent_action_p(init_p)
            sync(); t0 = time.perf_counter()
            for i in range(n_steps):
                env.step(acts[i])
            sync(); ts.append(time.perf_counter() - t0)
        t = float(np.median(ts[1:]))
        out['fused' if fuse else 'plain'] = {'steps_per_s': n_steps / t, 'substeps_per_s': 10 * n_steps / t, 'n_particles': env.n_particles, 'runs_s': ts[1:]}
    return out


if __name__ == '__main__':
    print(json.dumps(run()))
