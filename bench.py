#!/usr/bin/env python
"""bench.py — MPM substeps/s of the B200-native FluidEngine substep (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--bwd 1]

A bench "step" is one simulator step = 10 MLS-MPM substeps (MPM:30,749-751) of the workload BASELINE.json's metric
is quoted on: configs[1], single-material water block free fall, 1M particles, 128^3 grid, fp32, forward only
(SURVEY.md §8d C2).  `value` = substeps/s with the state resident in HBM; `e2e` = the same through the public
MPMSimulator API with host buffers (set_state from pinned host memory at every episode start, get_state_RL D2H every
step, as envs/fluid_env.py:131-150 does).  `roofline` is for the dominant kernel (p2g); `cpu_baseline` times the CPU
oracle (restatement of the reference algorithm; Taichi is not installable) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

N_PARTICLES = 1_000_000
QUALITY = 2  # 128^3
GRAVITY = (0.0, -10.0, 0.0)
LO, HI = (0.25, 0.30, 0.25), (0.75, 0.54, 0.75)
SUBSTEPS_PER_STEP = 10


def workload_particles(n=N_PARTICLES, seed=0, lo=LO, hi=HI):
    from fluidlab_b200 import macros as M
    x = np.random.RandomState(seed).uniform(lo, hi, size=(n, 3))
    return dict(x=x, mat=np.full(n, M.WATER, dtype=np.int32), used=np.ones(n, dtype=np.int32), rho=np.full(n, M.RHO[M.WATER]),
                body_id=np.zeros(n, dtype=np.int32), bodies={'n': 1})


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe): ONE nvidia-smi process in
    loop mode, started before and stopped after the timed region (forking a sampler per reading perturbs the launch loop)."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0, period_ms=50):
        self.index, self.period_ms, self.samples, self.proc = index, period_ms, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i', str(self.index),
                                          '-lms', str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.3)  # let it start sampling before the timed region
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.1)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill(); out = ''
            for line in out.strip().splitlines():
                self.samples.append([t.strip() for t in line.split(',')])

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace('.', '').isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace('.', '').isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), s[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(self.samples)}


def host_cores():
    """cores this process may run on (cgroup / affinity aware)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def oracle_lib_all_cores():
    """the CPU oracle, free to use ALL host cores: torchrun exports OMP_NUM_THREADS=1 to its workers, which round 1's reference arm inherited at
    N > 1 (cpu_baseline.cores = 1).  The thread count actually used is tuned per host by `tune_threads` (more threads are not always faster: the
    scatter's atomics and NUMA placement make the 128-thread run of a 128-core box 3x slower than its 32-thread run)."""
    os.environ['OMP_NUM_THREADS'] = str(host_cores())
    from oracle import oracle as orc
    L = orc.lib()
    L.orc_set_threads(host_cores())
    return orc, L


def tune_threads(L, o, probe=2):
    """fastest OpenMP thread count for this host among {all, 1/2, 1/4, 1/8 of the cores, 16, 8}: `probe` substeps each (the reference arm gets the
    CPU's best configuration; SURVEY.md 8d: the speed-up target is against the faster CPU number)"""
    n = host_cores()
    best = (None, 0.0)
    for t in sorted({n, max(1, n // 2), max(1, n // 4), max(1, n // 8), min(n, 16), min(n, 8)}, reverse=True):
        L.orc_set_threads(t)
        o.substep(0)
        t0 = time.perf_counter()
        for i in range(probe):
            o.substep(i % 2)
        r = probe / (time.perf_counter() - t0)
        if r > best[1]:
            best = (t, r)
    L.orc_set_threads(best[0])
    return best[0]


def cpu_baseline_run(n_particles, max_seconds=15.0, max_substeps=60, threads=None):
    """Time the CPU oracle (restatement of mpm_simulator.py:515-533, fp32, OpenMP) on the same workload."""
    from conftest import make_particles
    from fluidlab_b200 import macros as M
    wp = workload_particles(n_particles)
    P = make_particles(wp['x'], M.WATER, 64 * QUALITY)
    orc, L = oracle_lib_all_cores()
    o = orc.OracleSim(64 * QUALITY, P, gravity=GRAVITY, max_substeps_local=2, precision=32)
    o.substep(0)  # warm-up (page faults, thread pool)
    cores = int(threads) if threads else tune_threads(L, o)
    L.orc_set_threads(cores)
    t0, n = time.perf_counter(), 0
    while n < max_substeps and (time.perf_counter() - t0) < max_seconds:
        o.substep(n % 2); n += 1
    dt = time.perf_counter() - t0
    return dict(value=n / dt, unit='substeps/s', cores=int(cores), kind='port', host_threads_available=host_cores(),
                sample=f'{n} forward substeps of the full workload ({n_particles} particles, {64 * QUALITY}^3 grid) in {dt:.1f}s with {cores} threads (fastest of 6 counts probed); '
                       'C++/OpenMP restatement of the reference algorithm (Taichi ti.cpu is not installable)')


def slab_layout(world, rank, N):
    """the multi-GPU arm's workload: one water body on a 256^3 grid cut into `world` x-slabs of N particles each (weak scaling)"""
    from fluidlab_b200.slab import slab_bounds
    q = 4; n = 64 * q; dx = 1.0 / n; slab_w = 24
    bounds = slab_bounds(32, 32 + slab_w * world, world)
    lo = ((bounds[rank] - 0.5) * dx, 0.30, 0.36); hi = ((bounds[rank + 1] - 0.5) * dx, 0.30 + 72 * dx, 0.36 + 72 * dx)
    return q, bounds, lo, hi


C5_N, C5_LO, C5_HI = 8_000_000, (0.10, 0.20, 0.10), (0.90, 0.42, 0.90)   # WATER, ~8.4 particles per cell at 256^3


def c5_shard(world, rank, n=C5_N):
    """BASELINE.json configs[4]: the 8M-particle water body on the 256^3 grid; rank r of `world` gets the particles whose stencil-centre plane lies
    in its x-slab (slabs of equal width over the body's 206 planes, boundaries on multiples of 8).  Returns (quality, bounds, particle dict, global ids)."""
    from fluidlab_b200 import macros as M
    from fluidlab_b200.slab import slab_bounds
    q, ng = 4, 256
    x = np.random.RandomState(0).uniform(C5_LO, C5_HI, size=(n, 3))
    bounds = slab_bounds(24, 232, world) if world > 1 else [24, 232]
    cp = (x[:, 0] * ng - 0.5).astype(np.int32) + 1
    lo = bounds[rank] if rank > 0 else -10 ** 6
    hi = bounds[rank + 1] if rank < world - 1 else 10 ** 6
    mine = np.where((cp >= lo) & (cp < hi))[0]
    parts = dict(x=x[mine], mat=np.full(len(mine), M.WATER, dtype=np.int32), used=np.ones(len(mine), dtype=np.int32), rho=np.full(len(mine), M.RHO[M.WATER]),
                 body_id=np.zeros(len(mine), dtype=np.int32), bodies={'n': 1})
    return q, bounds, parts, mine


def run_reference(args):
    """--impl reference: the reference's CPU path = the oracle port on all host cores, on the workload the repo arm runs at this N: C2 at N = 1, the
    N-slab water body on the 256^3 grid at N > 1 (rank 0 alone computes it; the other ranks exit).  Each bench step is a bounded sample of ONE
    substep (a 10-substep step of 1M particles is seconds of CPU).  `value` uses the repo arm's unit: 1M-particle substeps/s = N x global substeps/s."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from conftest import make_particles
    from fluidlab_b200 import macros as M
    orc, L = oracle_lib_all_cores()
    world = max(1, args.gpus)
    if world == 1:
        wp = workload_particles(N_PARTICLES)
        n_grid, x = 64 * QUALITY, wp['x']
        workload = 'C2 water block free fall, 1M particles, 128^3 grid, forward'
    else:
        xs = []
        for r in range(world):
            q, _, lo, hi = slab_layout(world, r, N_PARTICLES)
            xs.append(workload_particles(N_PARTICLES, seed=r, lo=lo, hi=hi)['x'])
        n_grid, x = 64 * q, np.concatenate(xs)
        workload = f'C2-weak: {world} x {N_PARTICLES} water particles as x-slabs of one body, 256^3 grid, forward (the repo arm\'s workload at {world} GPUs)'
    P = make_particles(x, M.WATER, n_grid)
    o = orc.OracleSim(n_grid, P, gravity=GRAVITY, max_substeps_local=2, precision=32)
    for _ in range(max(1, min(args.warmup, 2))):
        o.substep(0)
    cores = tune_threads(L, o)
    K = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for i in range(K):
        o.substep(i % 2)
    dt = time.perf_counter() - t0
    val = world * K / dt
    line = {'metric': 'mpm_substeps_per_s_fwd', 'value': val, 'unit': 'substeps/s', 'n_gpus': args.gpus, 'steps': K, 'warmup': args.warmup,
            'ms_per_step': 1e3 * dt / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'impl': 'reference',
            'config': {'workload': workload, 'step': '1 substep (bounded sample)', 'same_config_as_repo_arm': True},
            'cpu_baseline': {'value': val, 'unit': 'substeps/s', 'cores': int(cores), 'kind': 'port',
                             'sample': f'{K} forward substeps of the full workload ({len(x)} particles, {n_grid}^3 grid), {cores} of {host_cores()} host threads (the fastest of 6 thread counts probed); '
                                       'C++/OpenMP restatement of the reference (Taichi not installable)'},
            'e2e': {'value': val, 'unit': 'substeps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def run_config(args):
    """--config C3 | C4: BASELINE.json configs[2] / configs[3] through the reference-facing TaichiEnv API (tests/baseline_scenes.py builds the same scenes the
    parity tests compare with the oracle): forward substeps/s, forward + backward substep pairs/s (loss + dLoss/dAction, chunk re-simulation included) and
    an end-to-end figure with host actions and a get_state_RL read-back per step.  An extra to the C2 line: the driver's line stays the default config."""
    import torch
    from baseline_scenes import c3_env, c4_env
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    if args.config == 'C3':
        n_steps = max(1, min(args.steps, 50)) if args.steps != 100 else 10   # 50-step horizon; default 10 steps = 100 substeps = two T = 50 chunks
        env, actions, action_p = c3_env(n_steps, T=50, device_kw=dict(device=dev), scale=1.0 if args.particles == N_PARTICLES else args.particles / 262_144)
        workload = (f'C3 LatteArt two-material ({env.simulator.n_particles} slots: COFFEE in the cup + parked MILK, Injector flux 8, cylinder boundary), 128^3 grid, fp32, {n_steps} steps x 10 substeps, '
                    'T = 50 ring, LatteArtLoss, dLoss/dAction (BASELINE.json configs[2])')
    else:
        n_steps = 1
        small = args.particles != N_PARTICLES   # (script checks on the CPU execution-model shim: fewer particles on a coarser grid)
        env, actions, action_p = c4_env(n_steps=n_steps, n_grid=32 if small else 192, n_each=args.particles // 2 if small else 1_000_000, T=10, device_kw=dict(device=dev))
        workload = (f'C4 IceCream elastic + plasto-elastic, 2 x {env.simulator.n_particles // 2} particles, {env.simulator.n_grid}^3 grid, fp32, soft cone collider '
                    '(agent_icecreamdynamic.yaml:26-37) evaluated per particle but hovering above the blocks (no contact: dLoss/dAction is ~0; with the reference\'s fixed dt a '
                    'contact perturbation grows 5x per substep at 192^3, see tests/test_gpu_parity.py c4_case), 1 step x 10 substeps from rest per repetition, '
                    'IceCreamDynamicLoss, forward + backward (BASELINE.json configs[3])')
    sim = env.simulator
    st0 = env.get_state()['state']
    n_sub = n_steps * SUBSTEPS_PER_STEP

    def fwd(grad):
        env.set_state(st0, grad_enabled=grad)
        env.apply_agent_action_p(action_p)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n_steps):
            env.step(actions[i])
        b.record()
        return a, b

    def fwd_bwd():
        a, _ = fwd(True)
        env.get_final_loss()
        env.reset_grad(); env.get_final_loss_grad()
        for i in range(n_steps - 1, -1, -1):
            env.step_grad(actions[i])
        env.apply_agent_action_p_grad(action_p)
        b = torch.cuda.Event(enable_timing=True); b.record()
        return a, b

    def timed(fn, min_ms):
        for _ in range(max(args.warmup, 3) if n_steps == 1 else 2):
            fn()
        torch.cuda.synchronize()
        tot, reps = 0.0, 0
        while reps == 0 or (tot < min_ms and reps < 2000):
            a, b = fn(); torch.cuda.synchronize()
            tot += a.elapsed_time(b); reps += 1
        return tot / reps, reps
    with ClockSampler(0) as cs:
        ms_f, reps_f = timed(lambda: fwd(False), args.min_seconds * 1e3)
    clocks = cs.summary()
    ms_fb, reps_fb = timed(fwd_bwd, args.min_seconds * 1e3)
    grad = env.agent.get_grad(n_steps)
    x_ = sim.get_x()
    if not (np.isfinite(x_).all() and np.isfinite(grad).all()):
        raise RuntimeError(f'{args.config}: non-finite state / gradient after the timed region: the measurement is void')
    # end to end: host actions in, x / v / used of every step out (get_state_RL), set_state from host at every episode start
    d2h = sim.n_particles * 28
    h2d = sum(np.asarray(v).nbytes for k, v in st0.items() if k in ('x', 'v', 'C', 'F', 'used')) / n_steps

    def episode():
        env.set_state(st0, grad_enabled=False)
        env.apply_agent_action_p(action_p)
        for i in range(n_steps):
            env.step(actions[i])
            out = env.get_state_RL()
        return out
    episode(); torch.cuda.synchronize()
    t0, n_ep = time.perf_counter(), 0
    while n_ep == 0 or time.perf_counter() - t0 < args.min_seconds:
        episode(); n_ep += 1
    torch.cuda.synchronize()
    e2e = n_ep * n_sub / (time.perf_counter() - t0)
    line = {'metric': 'mpm_substeps_per_s_fwd', 'value': n_sub / (ms_f * 1e-3), 'unit': 'substeps/s', 'n_gpus': 1, 'steps': n_steps, 'warmup': max(args.warmup, 3),
            'timed_steps': n_steps * reps_f, 'ms_per_step': ms_f / n_steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload, 'substeps_per_step': SUBSTEPS_PER_STEP, 'dt': 2e-4, 'n_particle_slots': int(sim.n_particles),
                       'api': 'TaichiEnv.set_state / apply_agent_action_p / step(action) [/ get_final_loss / reset_grad / get_final_loss_grad / step_grad / apply_agent_action_p_grad]; '
                              'device time of the step loop (CUDA events), set_state outside the timed region',
                       'l2_policy': 'inputs larger than L2 (100 B x particle slots per frame, walked frame by frame)', 'parallelism': 'single GPU'},
            'clocks': clocks,
            'fwd_bwd': {'value': n_sub / (ms_fb * 1e-3), 'unit': 'substeps/s (each = 1 forward + 1 backward substep; loss, loss gradient and chunk re-simulation included)',
                        'ms_per_pass': ms_fb, 'passes_timed': reps_fb, 'dloss_daction_absmax': float(np.abs(grad).max())},
            'e2e': {'value': e2e, 'unit': 'substeps/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'api': 'TaichiEnv.set_state(host) / step(host action) / get_state_RL (blocking D2H of x, v, used every step)'},
            'gpu_launches': None}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='C2', choices=['C2', 'C3', 'C4'], help='C2 (default): BASELINE.json configs[1], the line the metric is quoted on; C3 / C4: configs[2] / configs[3] '
                    'through TaichiEnv (forward, forward + backward, e2e) as an extra line; C5 is --scaling strong')
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--particles', type=int, default=N_PARTICLES)
    ap.add_argument('--bwd', type=int, default=1, help='also time forward+backward (extra keys)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'], help="weak (default): 1M particles per GPU (C2 at N = 1; N x-slabs of one body on a 256^3 grid at N > 1); "
                    "strong: BASELINE.json configs[4] — C5, 8M water particles on a 256^3 grid, sharded into N x-slabs (N = 1: one GPU holds it all)")
    ap.add_argument('--slab-shape', type=int, default=0, help='diagnostic (N = 1 only): run ONE slab of the `--gpus SLAB_SHAPE` weak-scaling workload (1M particles as a 24-plane slab on '
                    'the 256^3 grid, 10 % spare slots) alone on one GPU through the single-GPU path — the per-GPU cost of that workload without any exchange')
    ap.add_argument('--min-substeps', type=int, default=1000, help='minimum number of substeps in the timed region (SURVEY.md 8d: >= 1,000)')
    ap.add_argument('--min-seconds', type=float, default=1.0, help='minimum length of the timed region (the K steps are repeated)')
    ap.add_argument('--fuse-g2p2g', type=int, default=1, help='1 (default, measured faster: profiles/README.md): forward steps use fmpm_substeps_fused (the gather of substep f and the '
                    'scatter of f+1 in one kernel, k_fwd); 0: the plain p2g / grid_op / g2p substeps')
    ap.add_argument('--sort-every', type=int, default=4, help='cell-sort period in steps (measured with the warp-local key ranking: 1 -> 7.14k, 2 -> 7.43k, 4 -> 7.60k, 8 -> 7.44k substeps/s)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.config != 'C2':
        return run_config(args)

    import torch
    import torch.distributed as dist
    from fluidlab_b200 import MPMSimulator

    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    W = max(args.warmup, 3)
    K = args.steps
    T = 50
    N = args.particles
    slab = None
    strong = args.scaling == 'strong'
    if strong:
        T = 20   # 8M particles x 21 frames x 100 B = 17 GB per GPU at N = 1
        args.bwd, args.no_cpu = 0, True   # the strong-scaling arm is an extra: forward throughput + e2e only
    if world == 1:
        if strong:
            q5, _, parts, _ = c5_shard(1, 0, args.particles if args.particles != N_PARTICLES else C5_N)
            N = len(parts['x'])
        sim = MPMSimulator(dim=3, quality=q5 if strong else (4 if args.slab_shape else QUALITY), gravity=GRAVITY, horizon=max(K + W + 4, 100) * 4, max_substeps_local=T, max_substeps_global=10 ** 7,
                           ckpt_dest='gpu', device=dev, sort_every=args.sort_every)
        if args.slab_shape:
            from fluidlab_b200 import macros as M_
            _, _, lo_, hi_ = slab_layout(args.slab_shape, 0, N)
            parts = workload_particles(N, seed=0, lo=lo_, hi=hi_)
            pad = int(N * 0.1) + 1024
            parts = dict(x=np.concatenate([parts['x'], np.tile(np.array(M_.NOWHERE, dtype=np.float64), (pad, 1))]), mat=np.concatenate([parts['mat'], np.full(pad, M_.WATER, np.int32)]),
                         used=np.concatenate([parts['used'], np.zeros(pad, np.int32)]), rho=np.concatenate([parts['rho'], np.full(pad, M_.RHO[M_.WATER])]),
                         body_id=np.zeros(N + pad, dtype=np.int32), bodies={'n': 1})
            args.bwd, args.no_cpu = 0, True
        elif not strong:
            parts = workload_particles(N, seed=rank)
        sim.build(None, None, [], parts)
        sim.fuse_g2p2g = bool(args.fuse_g2p2g)
        # The block falls 0.25 of the domain: free fall lasts ~110 steps (SURVEY.md 8d times 1,000 substeps after 100 warm-up).  Longer timed regions
        # replay that episode: every EPISODE steps the initial state is restored from a device-side copy INSIDE the timed region (~0.1 % of the time).
        EPISODE = 8 if strong else (30 if args.slab_shape else 100)    # C5 from rest turns non-finite after ~140 substeps with the reference's fixed dt (profiles/check_c5.py: c dt / dx = 0.85 at 256^3)
        _cnt1 = [0]
        _init1 = [None]

        def step_fn():
            if _init1[0] is not None and _cnt1[0] and _cnt1[0] % EPISODE == 0:
                sim.cur_substep_global = 0
                sim.set_state(0, _init1[0])
            _cnt1[0] += 1
            sim.step(None)
        workload = (f'C5 water body, {N} particles, 256^3 grid, fp32, forward (BASELINE.json configs[4], strong scaling: the whole body on one GPU)' if strong else
                    f'C2 water block free fall, {N} particles, 128^3 grid, fp32, forward (BASELINE.json configs[1])')
        if args.slab_shape:
            workload = (f'DIAGNOSTIC: one slab of the {args.slab_shape}-GPU weak-scaling workload alone ({N} water particles as a 24-plane slab on the 256^3 grid, 10 % spare slots), '
                        'single-GPU path, no exchange')
        parallelism = 'single GPU'
    else:
        # weak scaling: the C2 block (same particle count and ~8 particles/cell per GPU) laid out as x-slabs of one global
        # water body on a 256^3 grid; ghost planes of the (momentum, mass) grid are summed between neighbours every substep.
        from fluidlab_b200.slab import SlabMPMSimulator
        if strong:
            q, bounds, parts, gid5 = c5_shard(world, rank, args.particles if args.particles != N_PARTICLES else C5_N)
            N = len(parts['x'])
        else:
            q, bounds, lo, hi = slab_layout(world, rank, N)
            parts = workload_particles(N, seed=rank, lo=lo, hi=hi)
        slab = SlabMPMSimulator(q, GRAVITY, parts, gid=gid5 if strong else np.arange(N) + rank * N, bounds=bounds, capacity=int(N * 1.1) + 1024, max_substeps_local=T, device=dev,
                                exchange=os.environ.get('SLAB_EXCHANGE', 'peer'), sync=os.environ.get('SLAB_SYNC', 'signal'), sort_every=args.sort_every,
                                halo=int(os.environ.get('SLAB_HALO', '4')), migrate_every=int(os.environ.get('SLAB_MIGRATE_EVERY', '4')))
        sim = slab.sim
        sim.fuse_g2p2g = bool(args.fuse_g2p2g)   # x-slab mode: the fused kernel's scatter half reduces frame f+1's ghost planes into the neighbour
        # the reference's fixed dt = 2e-4 is unstable for water at 256^3 beyond ~700 substeps (profiles/check_stability_256.py,
        # SURVEY.md §8d C5): restore the initial state (device-side copy, ~0.1% of the time) every 30 steps
        _cnt = [0]
        _init_dev = {k: v.clone() for k, v in sim.readframe_torch(0).items()}
        _gid0 = slab.gid.clone()

        def step_fn():
            if _cnt[0] and _cnt[0] % (8 if strong else 30) == 0:   # (C5 is only stable for ~140 substeps from rest, profiles/check_c5.py)
                sim.cur_substep_global = 0
                sim.set_state(0, _init_dev); slab.gid.copy_(_gid0)
            _cnt[0] += 1
            slab.step()
        workload = ((f'C5 water body, {C5_N if args.particles == N_PARTICLES else args.particles} particles, 256^3 grid, fp32, forward, sharded into {world} x-slabs (BASELINE.json '
                     f'configs[4], strong scaling); value = substeps/s of the WHOLE body') if strong else
                    (f'C2-weak: {N} water particles per GPU (~8/cell) as {world} x-slabs of one body, 256^3 grid, fp32, forward; value counts '
                     f'1M-particle substeps (global substeps/s = value / n_gpus)'))
        if slab.exchange == 'peer':
            sync_txt = ('one neighbour handshake per substep inside the library, the whole step in one call' if slab.sync == 'signal'
                        else 'one device-side signal-pad barrier per substep')
            ghost_txt = ('ghost-plane reduction PULLED by grid_op (it adds the neighbours\' partial sums of the 2 x halo ghost planes, read over NVLink peer memory after the '
                         'handshake; the scatter stays local; ghost blocks cleared one handshake later)' if slab.pull else
                         'ghost-plane reduction fused into p2g (vector REDs into the neighbour grid over NVLink peer memory)')
            parallelism = (f'{world} x-slabs, halo {slab.halo}; {ghost_txt}, accumulators double-buffered by substep parity, {sync_txt}, no data-path collective; '
                           f'census + migration every {slab.migrate_every} step(s)')
        else:
            parallelism = (f'{world} x-slabs; NCCL pair all-reduce of {slab.ghost.bytes_per_exchange()} B of ghost planes per rank per substep; census + migration every {slab.migrate_every} step(s)')
    init = sim.get_state()
    if slab is None:
        _init1[0] = {k: v.clone() for k, v in sim.readframe_torch(0).items()}
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ------------------------------------------------------------------ device-resident throughput
    for _ in range(W):
        step_fn()
    barrier()
    # the timed region is `reps` x K steps, reps chosen (from one untimed probe of K steps, agreed across ranks) so that it lasts >= --min-seconds
    # and covers >= 1,000 substeps whatever --steps is (round 1's 20-step region was 29 ms)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(K):
        step_fn()
    p1.record(); barrier()
    probe_ms = max_over_ranks(p0.elapsed_time(p1))
    reps = max(1, int(np.ceil(args.min_seconds * 1e3 / max(probe_ms, 1e-3))), int(np.ceil(args.min_substeps / (K * SUBSTEPS_PER_STEP))))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as cs:
        e0.record()
        for _ in range(K * reps):
            step_fn()
        e1.record()
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms = ms_total / reps            # per K steps
    # a run that blew up (NaN positions freeze the particles: less work per substep) must not produce a number
    fcur = sim.cur_substep_local
    xs_, alive_ = sim.slab_positions(fcur)
    if not bool(torch.isfinite(xs_[alive_]).all().item()):
        raise RuntimeError(f'rank {rank}: non-finite particle positions after the timed region: the workload is unstable, the measurement is void')
    clocks = cs.summary()
    value = (1 if strong else world) * K * SUBSTEPS_PER_STEP / (ms * 1e-3)

    # ------------------------------------------------------------------ per-kernel timing (roofline)
    f = sim.cur_substep_local
    if f >= T:
        f = 0
    used = int(sim.readframe_torch(f, ('used',))['used'].sum().item())

    def time_phase(fn, n=20, pre=None):
        """average device time of fn(): n back-to-back launches between two events (no per-launch event overhead);
        a `pre` callable (e.g. the grid clear) runs before every fn() and its own batched time is subtracted."""
        def batch(body):
            for _ in range(3):
                body()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record()
            for _ in range(n):
                body()
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) / n
        if pre is None:
            return batch(fn)
        return batch(lambda: (pre(), fn())) - batch(pre)

    sort_age = None
    if slab is None:
        # Average launch duration of every forward kernel OVER THE TIMED TRAJECTORY: the same W + K steps are replayed from the initial
        # state on the per-substep path with CUDA events between the three launches of each substep (the stream stays saturated:
        # ~40 us of host work per substep against ~130 us of device work), so stale-sort states weigh in exactly as they do in `value`.
        L_, h_ = sim._lib, sim._h
        evs, gts, rec = [], [], [False]
        orig_substep, orig_graphs, orig_fuse = L_.fmpm_substep, sim.use_graphs, sim.fuse_g2p2g
        sim.fuse_g2p2g = False   # the per-kernel replay times the plain p2g / grid_op / g2p kernels (they are what SURVEY 8d's byte counts describe)

        def timed_substep(h, fr, stream):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record(); rc = L_.fmpm_p2g(h, fr, 1, stream)
            e[1].record(); rc |= L_.fmpm_grid_op(h, fr, 1, stream)
            e[2].record(); rc |= L_.fmpm_g2p(h, fr, stream)
            e[3].record()
            if rec[0]:
                evs.append(e)
            return rc
        sim.cur_substep_global = 0
        sim.set_state(0, init)
        sim.use_graphs = False
        L_.fmpm_substep = timed_substep
        try:
            for i in range(W + K):
                rec[0] = i >= W
                if rec[0] and (i - W) % max(1, K // 4) == 0:   # touched-node census at 4 points of the run (p2g of the current frame, accumulator only)
                    fc = sim.cur_substep_local
                    sim.phase('clear_grid', fc); sim.phase('p2g', fc, 0)
                    gts.append(int((sim._grid_pm.reshape(-1, 4)[:, 3] > 0).sum().item()))
                    sim.phase('grid_op', fc, 1)
                sim.step(None)
            torch.cuda.synchronize()
        finally:
            L_.fmpm_substep = orig_substep
            sim.use_graphs = orig_graphs
            sim.fuse_g2p2g = orig_fuse
        t_p2g = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
        t_gop = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
        t_g2p = float(np.mean([e[2].elapsed_time(e[3]) for e in evs]))
        g_t = int(np.mean(gts))
        if args.sort_every > 1:   # how much the scatter slows down as the cell sort ages (steps since the last sort: 0, 1, ...)
            per_age = [[[], []] for _ in range(args.sort_every)]
            for j, e in enumerate(evs):
                age = (W + j // SUBSTEPS_PER_STEP) % args.sort_every
                per_age[age][0].append(e[0].elapsed_time(e[1])); per_age[age][1].append(e[2].elapsed_time(e[3]))
            sort_age = {'p2g': [float(np.mean(a[0])) for a in per_age if a[0]], 'g2p': [float(np.mean(a[1])) for a in per_age if a[1]]}
        timing_note = f'mean over the {len(evs)} substeps of a replay of the timed trajectory (CUDA events between the launches)'
    else:
        t_p2g = time_phase(lambda: sim.phase('p2g', f, 0), pre=lambda: sim.phase('clear_grid', f))
        sim.phase('clear_grid', f); sim.phase('p2g', f, 0)
        g_t = int((sim._grid_pm.reshape(-1, 4)[:, 3] > 0).sum().item())
        def _refill():
            sim.phase('clear_grid', f); sim.phase('p2g', f, 0)
        t_gop = time_phase(lambda: sim.phase('grid_op', f, 0), pre=_refill)
        # g2p writes frame f+1: time it on a scratch frame pair (f -> f+1 is rewritten by the next step anyway)
        t_g2p = time_phase(lambda: sim._ck(sim._lib.fmpm_g2p(sim._h, f, sim._stream()), 'g2p'))
        # In peer mode every p2g above also reduced into the NEIGHBOURS' accumulators, unsynchronised: all ranks must have finished their
        # isolated launches before anyone clears, or a late reduction lands in an accumulator that the next step assumes clear (found by
        # running this arm on the CPU execution-model shim with skewed ranks, tests/cuda_emu/run_bench_emu.py).
        barrier()
        sim.phase('clear_grid', f)
        barrier()
        timing_note = 'batched launches on the final state'
    peak, peak_src = peaks()
    p2g_bytes = 136 * used + 16 * g_t            # p2g particle bytes + accumulated grid write-back
    g2p_bytes = 76 * used + 12 * g_t             # g2p(+advect)
    def traffic_of(kernel):
        """dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed `ncu --set full` captures of this same workload"""
        for name in ('r02_traffic.json', 'r01_traffic.json'):
            tp = os.path.join(ROOT, 'profiles', name)
            if os.path.exists(tp) and world == 1 and N == N_PARTICLES:
                v = json.load(open(tp)).get(kernel)
                if v is not None:
                    return v
        return None
    traffic = traffic_of('k_p2g')
    roof = {'bound': 'hbm', 'kernel': 'k_p2g', 'achieved': p2g_bytes / (t_p2g * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
            'frac': p2g_bytes / (t_p2g * 1e-3) / 1e9 / peak, 'traffic': traffic, 'peak_source': peak_src,
            'algorithmic_bytes_per_launch': p2g_bytes, 'launch_ms': t_p2g, 'n_used': used, 'touched_nodes': g_t, 'timing': timing_note}
    roof_pair = {'kernels': 'k_p2g+k_g2p', 'achieved': (p2g_bytes + g2p_bytes) / ((t_p2g + t_g2p) * 1e-3) / 1e9,
                 'frac': (p2g_bytes + g2p_bytes) / ((t_p2g + t_g2p) * 1e-3) / 1e9 / peak, 'p2g_ms': t_p2g, 'g2p_ms': t_g2p, 'grid_op_ms': t_gop,
                 'bytes': p2g_bytes + g2p_bytes, 'ms_by_steps_since_sort': sort_age}

    # The fused path's own figures: the timed trajectory replayed launch by launch (fmpm_p2g, then [fmpm_grid_op, fmpm_fwd_step] x 9, fmpm_grid_op,
    # fmpm_g2p: exactly what fmpm_substeps_fused enqueues) with CUDA events around every launch.  One fused launch does the work SURVEY 8(d)
    # counts for p2g + g2p of a substep (212 B x N_u + 28 B x G_t), so `achieved` uses that figure and the fraction is comparable with
    # roofline_p2g_g2p; the bytes the kernel itself moves are fewer and stated.  Event-bracketed launches read a few per cent longer than the
    # same launches inside the CUDA graph of the timed region (round 1's verdict: 153.9 vs 145.3 us per substep), so every launch time is
    # also given scaled by graph substep time / replayed substep time (`launch_ms`, used for `frac`; the raw mean is `launch_ms_events`).
    roof_fused = None
    if slab is None and args.fuse_g2p2g:
        try:
            L_, h_, st_ = sim._lib, sim._h, sim._stream
            path = int(L_.fmpm_fwd_path(h_))
            ev = lambda: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            e_fused, e_gop, e_p2g, e_g2p, e_step = [], [], [], [], []
            sim.cur_substep_global = 0
            sim.set_state(0, init)
            for i in range(W + K):
                es = ev(); es[0].record()
                if sim.sort_every > 0 and sim.cur_step_global % sim.sort_every == 0:
                    sim.sort_frame(sim.cur_substep_local)
                f0 = sim.cur_substep_local
                e = ev(); e[0].record(); sim.phase('p2g', f0, 1); e[1].record(); e_p2g.append(e)
                for j in range(SUBSTEPS_PER_STEP):
                    fj = f0 + j
                    e = ev(); e[0].record(); sim.phase('grid_op', fj, 1); e[1].record(); e_gop.append(e)
                    e = ev(); e[0].record()
                    if j + 1 < SUBSTEPS_PER_STEP:
                        sim._ck(L_.fmpm_fwd_step(h_, fj, int(j + 2 == SUBSTEPS_PER_STEP), st_()), 'fmpm_fwd_step'); e[1].record(); e_fused.append(e)
                    else:
                        sim.phase('g2p', fj); e[1].record(); e_g2p.append(e)
                    sim._frame_ord[fj + 1] = sim._frame_ord[fj]
                    sim.cur_substep_global += 1
                es[1].record(); e_step.append(es)
                if sim.cur_substep_local == 0:
                    sim.memory_to_cache()
            torch.cuda.synchronize()
            mean = lambda evs: float(np.mean([a.elapsed_time(b) for a, b in evs[W * (len(evs) // (W + K)):]]))
            t_fused, t_gop_f, t_p2g_f, t_g2p_f, t_step = mean(e_fused), mean(e_gop), mean(e_p2g), mean(e_g2p), mean(e_step)
            scale = min(1.0, (ms / K) / t_step)     # CUDA-graph step of the timed region / event-bracketed replay of the same step
            pair_bytes = p2g_bytes + g2p_bytes
            liquid = bool(path & 2)
            moved = (40 if liquid else 104) * used + 28 * g_t
            kname = ('k_fwd<all-liquid>' if liquid else 'k_fwd') if path & 1 else 'k_g2p2g'
            roof_fused = {'bound': 'hbm', 'kernel': kname, 'fwd_path': path, 'achieved': pair_bytes / (t_fused * scale * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                          'frac': pair_bytes / (t_fused * scale * 1e-3) / 1e9 / peak, 'launch_ms': t_fused * scale, 'launch_ms_events': t_fused,
                          'event_to_graph_scale': scale, 'algorithmic_bytes_per_launch': pair_bytes, 'bytes_the_fused_kernel_moves': moved,
                          'traffic': traffic_of('k_fwd' if path & 1 else 'k_g2p2g'),
                          'peak_source': peak_src, 'n_used': used, 'touched_nodes': g_t, 'launches_timed': len(e_fused) * K // (W + K),
                          'other_launches_ms': {'first_p2g': t_p2g_f * scale, 'grid_op': t_gop_f * scale, 'last_g2p': t_g2p_f * scale},
                          'substep_ms_graph': ms / K / SUBSTEPS_PER_STEP, 'substep_ms_events': t_step / SUBSTEPS_PER_STEP,
                          'timing': 'mean over the fused launches of a launch-by-launch replay of the timed trajectory, scaled to the CUDA-graph step time'}
        except Exception as ex:
            roof_fused = {'error': f'{type(ex).__name__}: {ex}'}

    # ------------------------------------------------------------------ forward+backward (BASELINE metric, second half)
    fb = None
    if args.bwd and world == 1:
        import fluidlab_b200  # noqa
        nfb = max(2, min(K, 10))
        tgt = torch.zeros((N, 3), dtype=torch.float32, device=dev) + 0.5
        mask = sim.material_row_mask(fluidlab_b200.macros.WATER)

        def fwd_bwd():
            sim.set_state(0, init); sim.enable_grad()
            for _ in range(nfb):
                sim.step(None)
            sim.reset_grad()
            sim.add_x_grad_chamfer(tgt, mask, 1.0)
            for _ in range(nfb):
                sim.step_grad(None)
        fwd_bwd(); barrier()
        fb_runs = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fwd_bwd(); b.record(); barrier()
            fb_runs.append(max_over_ranks(a.elapsed_time(b)))
        fb_ms = float(np.median(fb_runs))
        # backward kernels alone (frame 0 of the ring holds valid state, adjoint buffer 0 the final adjoint)
        L_, h_, st_ = sim._lib, sim._h, sim._stream
        t_pg = time_phase(lambda: sim._ck(L_.fmpm_particle_grad(h_, 0, 0, 1, st_()), 'particle_grad'))
        t_sc = time_phase(lambda: sim._ck(L_.fmpm_g2p_grad_scatter(h_, 0, 0, st_()), 'g2p_grad_scatter'))
        sim.phase('clear_grid', 0)
        fb = {'value': world * nfb * SUBSTEPS_PER_STEP / (fb_ms * 1e-3), 'unit': 'substeps/s (each = 1 forward + 1 backward substep, incl. chunk re-simulation and set_state)',
              'steps': nfb, 'runs_ms': fb_runs, 'particle_grad_ms': t_pg, 'g2p_grad_scatter_ms_incl_dense_clear': t_sc}
        sim.disable_grad()
        # The same pass with a ring that holds the WHOLE trajectory (max_substeps_local = (steps + 1) * 10 instead of the reference's 50):
        # (steps * 10 + 1) frames x 100 B x N + the per-frame grids are ~18 GB at 1M particles / 100 substeps — nothing on a 180 GB B200 — and
        # the backward pass no longer re-simulates every chunk (MPM:856-912).  Extra key; the headline fwd_bwd keeps the reference's scheme.
        try:
            T2 = (nfb + 1) * SUBSTEPS_PER_STEP
            sim2 = MPMSimulator(dim=3, quality=QUALITY, gravity=GRAVITY, horizon=max(K + W + 4, 100) * 4, max_substeps_local=T2, max_substeps_global=10 ** 7,
                                ckpt_dest='gpu', device=dev, sort_every=args.sort_every)
            sim2.build(None, None, [], parts)
            sim2.fuse_g2p2g = bool(args.fuse_g2p2g)
            tgt2, mask2 = tgt, sim2.material_row_mask(fluidlab_b200.macros.WATER)

            mid = [None]

            def fwd_bwd2():
                sim2.set_state(0, init); sim2.enable_grad()
                for _ in range(nfb):
                    sim2.step(None)
                sim2.reset_grad()
                sim2.add_x_grad_chamfer(tgt2, mask2, 1.0)
                mid[0] = torch.cuda.Event(enable_timing=True); mid[0].record()
                for _ in range(nfb):
                    sim2.step_grad(None)
            fwd_bwd2(); barrier()
            runs2, bwd2 = [], []
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fwd_bwd2(); b.record(); barrier()
                runs2.append(a.elapsed_time(b)); bwd2.append(mid[0].elapsed_time(b))
            fb['whole_trajectory_ring'] = {'value': nfb * SUBSTEPS_PER_STEP / (float(np.median(runs2)) * 1e-3), 'unit': 'substeps/s (1 forward + 1 backward substep each, no re-simulation)',
                                           'max_substeps_local': T2, 'runs_ms': runs2, 'stored_grids': sim2._pm_ring is not None}
            # the backward substep against SURVEY 8(d)'s byte count for it (432 B x N_u + 156 B x G_t: k_g2p_grad_scatter + k_grid_op_grad +
            # k_particle_grad, forward grids kept in the ring): mean device time of a backward substep over the nfb x 10 of the pass above
            t_bs = float(np.median(bwd2)) / (nfb * SUBSTEPS_PER_STEP)
            bwd_bytes = 432 * used + 156 * g_t
            fb['roofline_bwd'] = {'bound': 'hbm', 'kernels': 'k_g2p_grad_scatter + k_grid_op_grad + k_particle_grad (+ the sparse clear of the adjoint grid)',
                                  'achieved': bwd_bytes / (t_bs * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s', 'frac': bwd_bytes / (t_bs * 1e-3) / 1e9 / peak,
                                  'backward_substep_ms': t_bs, 'algorithmic_bytes_per_substep': bwd_bytes,
                                  'timing': f'device time of the {nfb * SUBSTEPS_PER_STEP} backward substeps of the whole-trajectory pass (CUDA events), median of 3'}
            del sim2
            torch.cuda.empty_cache()
        except Exception as ex:   # an extra: never let it take the bench line down
            fb['whole_trajectory_ring'] = {'error': f'{type(ex).__name__}: {ex}'}
        # the optimiser step that closes one solver iteration (optimizer/solver.py:62-67): Adam on LatteArt's 251 x 3 action table, on the device
        try:
            import types as _types
            from fluidlab_b200.optimizer import Adam as _Adam
            _ad = _Adam((251, 3), _types.SimpleNamespace(type='Adam', lr=0.05, beta_1=0.9, beta_2=0.999, epsilon=1e-8)); _ad.bind(sim)
            _tab = torch.zeros((251, 3), dtype=torch.float64, device=dev); _gr = torch.randn((251, 3), dtype=torch.float32, device=dev)
            fb['adam_step_ms'] = time_phase(lambda: _ad.step_device(_tab, _gr, clip=(-1.0, 1.0)))
        except Exception as ex:
            fb['adam_step_ms'] = {'error': f'{type(ex).__name__}: {ex}'}

    # ------------------------------------------------------------------ end to end through the public API with host buffers
    pin = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in init.items() if k in ('x', 'v', 'C', 'F', 'used')}
    EP = 10  # steps per episode
    n_ep = max(1, K // EP)
    h2d = sum(t.numel() * t.element_size() for t in pin.values()) / EP
    d2h = sim.n_particles * (12 + 12 + 4)
    gid0 = slab.gid.clone() if slab is not None else None

    staged = [None]

    def episode(pipelined):
        sim.cur_substep_global = 0
        if pipelined:                              # H2D of the episode's initial state (pinned host) on the copy stream: this episode's state was uploaded while
            st_ = staged[0] if staged[0] is not None else sim.stage_state_async(pin)   # the previous one was stepping, the next one's upload starts now
            sim.set_state(0, st_)
            staged[0] = sim.stage_state_async(pin)
        else:
            sim.set_state(0, pin)                  # the reference's call sequence: a blocking upload on the compute stream
        if slab is not None:
            slab.gid.copy_(gid0)
        out, pend = None, None
        for _ in range(EP):
            step_plain()
            if pipelined:                          # D2H of x, v, used of EVERY step (FluidEnv._get_obs), consumed one step later: the copy overlaps the next step
                nxt = sim.get_state_RL_async()
                if pend is not None:
                    out = pend.result()
                pend = nxt
            else:
                out = sim.get_state_RL()           # the reference's blocking call sequence (MPM:683-696)
        if pend is not None:
            out = pend.result()
        return out

    def time_episodes(pipelined):
        episode(pipelined); barrier()
        t0 = time.perf_counter()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n_ep):
            episode(pipelined)
        b.record(); barrier()
        ms_ = max_over_ranks(max(a.elapsed_time(b), (time.perf_counter() - t0) * 1e3))
        return (1 if strong else world) * n_ep * EP * SUBSTEPS_PER_STEP / (ms_ * 1e-3)
    step_plain = (lambda: slab.step()) if slab is not None else (lambda: sim.step(None))
    e2e_blocking = time_episodes(False)
    e2e_val = time_episodes(True)

    # the same episodes with the observation assembled on the device (MPMSimulator.get_obs_RL): FluidEnv._get_obs keeps
    # ~200 particles per body, so the per-step D2H shrinks from 28 B x N to a few KB.  Reported as an EXTRA key; `e2e` stays the full-state API.
    e2e_obs = None
    if slab is None:
        try:
            def episode_obs():
                sim.cur_substep_global = 0
                sim.set_state(0, pin)
                out = None
                for _ in range(EP):
                    step_fn()
                    out = sim.get_obs_RL(200)
                return out
            o = episode_obs(); barrier()
            t0 = time.perf_counter()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n_ep):
                episode_obs()
            b.record(); barrier()
            ms_o = max(a.elapsed_time(b), (time.perf_counter() - t0) * 1e3)
            e2e_obs = {'value': n_ep * EP * SUBSTEPS_PER_STEP / (ms_o * 1e-3), 'unit': 'substeps/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(o.nbytes),
                       'api': 'MPMSimulator.set_state(pinned host)/step/get_obs_RL(200): FluidEnv._get_obs assembled on the device'}
        except Exception as ex:   # never let the extra measurement take the bench line down
            e2e_obs = {'error': f'{type(ex).__name__}: {ex}'}

    if rank == 0:
        cpu = None if args.no_cpu else cpu_baseline_run(N)
        line = {
            'metric': 'mpm_substeps_per_s_fwd', 'value': value, 'unit': 'substeps/s', 'n_gpus': world, 'steps': K, 'warmup': W,
            'timed_steps': K * reps, 'timed_region_ms': ms_total,
            'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload,
                       'substeps_per_step': SUBSTEPS_PER_STEP, 'dt': 2e-4, 'gravity': GRAVITY, 'max_substeps_local': T,
                       'cell_sort_every_steps': args.sort_every,
                       'g2p2g_fused': bool(args.fuse_g2p2g),
                       'l2_policy': 'inputs larger than L2 (one substep touches >= 212 B x 1M particles = 212 MB > 126 MB L2)',
                       'parallelism': parallelism},
            'clocks': clocks,
            'e2e': {'value': e2e_val, 'unit': 'substeps/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'api': 'MPMSimulator.stage_state_async(pinned host) + set_state (every episode\'s initial state uploaded on a copy stream while the previous episode steps) / step / '
                           'get_state_RL_async (x, v, used of every step read back on the copy stream and consumed one step later), 10-step episodes; one 100 B x N upload and ten '
                           '28 B x N read-backs per episode, all inside the timed region',
                    'blocking_api_value': e2e_blocking,
                    'blocking_api': 'the same with the reference\'s blocking get_state_RL after every step (MPM:683-696): the D2H is serialised with the steps'},
            'gpu_launches': K * ((2 * SUBSTEPS_PER_STEP + 1) if args.fuse_g2p2g else SUBSTEPS_PER_STEP * 3) + (2 * ((K + args.sort_every - 1) // args.sort_every) if args.sort_every else 0),   # p2g, grid_op, g2p per substep + k_sort_keys, k_reorder per cell sort (CUB's own kernels not counted)
            'e2e_obs_bridge': e2e_obs,
            'roofline': roof_fused if (roof_fused and 'frac' in roof_fused) else roof,
            'roofline_unfused_p2g': roof, 'roofline_p2g_g2p': roof_pair, 'roofline_fused': roof_fused,
            'fwd_bwd': fb,
            'cpu_baseline': cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
