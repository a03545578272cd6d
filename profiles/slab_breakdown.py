"""Diagnostic (2+ ranks): where does the per-substep time of the slab path go?  Variants of the substep loop, device-timed."""
import os, sys, time, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fluidlab_b200.slab import SlabMPMSimulator, slab_bounds
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local); dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
q = 4; n = 64 * q; dx = 1.0 / n
bounds = slab_bounds(32, 32 + 24 * world, world)
lo = ((bounds[rank] - 0.5) * dx, 0.30, 0.36); hi = ((bounds[rank + 1] - 0.5) * dx, 0.30 + 72 * dx, 0.36 + 72 * dx)
parts = bench.workload_particles(1_000_000, seed=rank, lo=lo, hi=hi)
MODE = os.environ.get('SLAB_EXCHANGE', 'peer')
slab = SlabMPMSimulator(q, (0, -10, 0), parts, gid=np.arange(1_000_000) + rank * 1_000_000, bounds=bounds, capacity=1_100_000, device=dev, exchange=MODE)
sim = slab.sim
for _ in range(3): slab.step()
init = {k: v.clone() for k, v in sim.readframe_torch(sim.cur_substep_local).items()}

def loop(n, exch, flags, mig=False):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
    for i in range(n):
        if mig and i % 10 == 0: slab._migrate(); sim.sort_frame(sim.cur_substep_local)
        f = sim.cur_substep_local
        sim.phase('p2g', f, 1)
        if exch and slab.exchange == 'nccl': slab.ghost.exchange_sum(sim._grid_pm)
        if exch and slab.exchange == 'peer': slab._symm.barrier(channel=0)
        if flags: slab.ghost.flag_ghost_blocks(sim._blk_flags)
        sim.phase('grid_op', f, 1); sim.phase('g2p', f)
        sim.cur_substep_global += 1
        if sim.cur_substep_local == 0: sim.memory_to_cache()
    cpu = (time.perf_counter() - t0) / n * 1e6
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3, cpu
for name, kw in (('no exchange', dict(exch=False, flags=False)), ('flags only', dict(exch=False, flags=True)), ('exchange only', dict(exch=True, flags=False)),
                 ('exchange+flags', dict(exch=True, flags=True)), ('full (+migrate+sort per 10)', dict(exch=True, flags=True, mig=True))):
    sim.cur_substep_global = 0; sim.set_state(0, init); sim.phase('clear_grid', 0); sim._blk_flags.zero_()
    loop(20, **kw)
    g, c = loop(100, **kw)
    if rank == 0: print(f'[{slab.exchange}] {name:32s} gpu {g:7.1f} us/substep   cpu-issue {c:7.1f} us/substep', flush=True)
dist.destroy_process_group()
