"""Diagnostic (N ranks): device time of each phase of the slab substep (events around every phase, so launch gaps are included)."""
import os, sys, numpy as np, torch, torch.distributed as dist
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
slab = SlabMPMSimulator(q, (0, -10, 0), parts, gid=np.arange(1_000_000) + rank * 1_000_000, bounds=bounds, capacity=1_100_000, device=dev,
                        exchange=os.environ.get('SLAB_EXCHANGE', 'peer'))
sim = slab.sim
for _ in range(5): slab.step()
names = ['p2g', 'barrier', 'flags', 'grid_op', 'g2p']
tot = {k: 0.0 for k in names}; nsub = 0
for step in range(10):
    slab._migrate(); sim.sort_frame(sim.cur_substep_local)
    for _ in range(10):
        f = sim.cur_substep_local
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record(); sim.phase('p2g', f, 1)
        ev[1].record(); slab._symm.barrier(channel=0) if slab.exchange == 'peer' else slab.ghost.exchange_sum(sim._grid_pm)
        ev[2].record()
        if slab.exchange != 'peer': slab.ghost.flag_ghost_blocks(sim._blk_flags)
        ev[3].record(); sim.phase('grid_op', f, 1)
        ev[4].record(); sim.phase('g2p', f)
        ev[5].record(); torch.cuda.synchronize()
        for i, k in enumerate(names): tot[k] += ev[i].elapsed_time(ev[i + 1]) * 1e3
        nsub += 1
        sim.cur_substep_global += 1
        if sim.cur_substep_local == 0: sim.memory_to_cache()
print(f'rank {rank} [{slab.exchange}] ' + '  '.join(f'{k} {tot[k]/nsub:6.1f}us' for k in names), flush=True)
dist.destroy_process_group()
