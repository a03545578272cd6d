"""Microbenchmark: 2-rank in-place all-reduce / send-recv of a ghost-plane sized buffer (what fluidlab_b200/slab.py does per substep)."""
import os, sys, torch, torch.distributed as dist
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local); dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
pg = dist.new_group([0, 1])
for mb in (1, 4, 8, 16):
    buf = torch.ones(mb * 1024 * 1024 // 4, device=dev)
    for _ in range(5):
        dist.all_reduce(buf, group=pg)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        dist.all_reduce(buf, group=pg)
    b.record(); torch.cuda.synchronize()
    t_ar = a.elapsed_time(b) / 50
    peer = 1 - rank
    rbuf = torch.empty_like(buf)
    def sr():
        ops = [dist.P2POp(dist.isend, buf, peer), dist.P2POp(dist.irecv, rbuf, peer)]
        for w in dist.batch_isend_irecv(ops): w.wait()
    for _ in range(5): sr()
    torch.cuda.synchronize(); a.record()
    for _ in range(50): sr()
    b.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f'{mb} MiB: all_reduce {t_ar*1e3:.1f} us, sendrecv {a.elapsed_time(b)/50*1e3:.1f} us', flush=True)
dist.destroy_process_group()
