"""TEST INFRASTRUCTURE: run bench.py's `ours` arm on the CUDA execution-model shim with a tiny workload, to check the SCRIPT (argument
handling, replay, e2e episodes, JSON line) in a container without a GPU.  The numbers it prints are meaningless.

    python tests/cuda_emu/run_bench_emu.py --particles 3000 --steps 2 --warmup 1 --no-cpu [--fuse-g2p2g 1]
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
import torch  # noqa: E402
import harness  # noqa: E402


class _TimedEvent:
    def __init__(self, *a, **k):
        self.t = None

    def record(self, *a, **k):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


def main():
    harness.enable()
    torch.cuda.Event = _TimedEvent
    torch.cuda.set_device = lambda *a, **k: None
    real_device = torch.device
    torch.device = lambda *a, **k: real_device('cpu') if a and a[0] == 'cuda' else real_device(*a, **k)
    from fluidlab_b200 import simulator
    init = simulator.MPMSimulator.__init__

    def emu_init(self, *a, **k):
        init(self, *a, **k)
        self.use_graphs = False
    simulator.MPMSimulator.__init__ = emu_init
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:   # multi-rank arm: gloo instead of nccl, POSIX shared memory instead of NVLink peer memory
        import torch.distributed as dist
        real_init = dist.init_process_group
        dist.init_process_group = lambda backend=None, **k: real_init('gloo', **{kk: vv for kk, vv in k.items() if kk != 'device_id'})
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        import test_cuda_emu_mpm
        from fluidlab_b200 import slab
        slab.SymmetricMemoryPeers = test_cuda_emu_mpm.ShmPeers
    import bench
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        # a few thousand particles spread over the real slab box are a sparse gas that the fixed dt blows up within 40 substeps (single domain
        # too); keep the workload's ~8 particles per cell by filling only a 6 x 8 x 8-cell corner of the box at the shared slab boundary
        real_wp = bench.workload_particles

        def dense_corner(n, seed=0, lo=None, hi=None):
            if lo is None:
                return real_wp(n, seed=seed)
            dx = 1.0 / 256
            rank = int(os.environ['RANK'])
            x0 = hi[0] - 6 * dx if rank % 2 == 0 else lo[0]
            return real_wp(n, seed=seed, lo=(x0, lo[1], lo[2]), hi=(x0 + 6 * dx, lo[1] + 8 * dx * max(1, n // 3000), lo[2] + 8 * dx))
        bench.workload_particles = dense_corner
    bench.main()


if __name__ == '__main__':
    main()
