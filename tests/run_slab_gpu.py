"""Multi-GPU parity driver (launched by torchrun, one rank per GPU): the x-slab sharded forward simulation must reproduce
the single-GPU simulation of the same particles (fp32 summation-order tolerance).  Particles get a lateral velocity so they
cross slab boundaries (ghost exchange AND migration are exercised).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/run_slab_gpu.py

SLAB_MODE=backward: the sharded backward pass (SlabMPMSimulator.step_grad: ghost sums of the accumulator and of the v_out adjoint,
migrate_grad) must reproduce the single-GPU dLoss/d(x0, v0, C0, F0) of L = sum |x_T - target|^2.
"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    from fluidlab_b200 import MPMSimulator, macros as M
    from fluidlab_b200.slab import SlabMPMSimulator, slab_bounds, centre_plane
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    q, n = 1, 64
    rng = np.random.RandomState(5)
    Ntot = 150000
    x = rng.uniform((0.2, 0.3, 0.3), (0.8, 0.5, 0.7), size=(Ntot, 3)).astype(np.float32)
    v0 = np.tile(np.array([3.0, 0.0, 0.5], dtype=np.float32), (Ntot, 1))  # 3 m/s along x: ~0.4 cells per step -> migration
    mat = np.where(x[:, 2] < 0.5, M.WATER, M.ELASTIC).astype(np.int32)
    # F0 away from the identity: at F = I the SVD adjoint of the ELASTIC half is degenerate (equal singular values) and the backward test below would
    # measure amplified round-off (round 2, hardware: single-GPU gC / gF differ by 0.25 from THEMSELVES when only the summation order changes)
    F0 = (np.eye(3)[None] + rng.randn(Ntot, 3, 3) * (0.04 if os.environ.get('SLAB_MODE', 'forward') == 'backward' else 0.0)).astype(np.float32)
    bounds = slab_bounds(0, 64, world) if world > 2 else slab_bounds(8, 56, world)
    cp = centre_plane(torch.from_numpy(x), float(n)).numpy()
    lo = bounds[rank] if rank > 0 else -10 ** 6
    hi = bounds[rank + 1] if rank < world - 1 else 10 ** 6
    mine = np.where((cp >= lo) & (cp < hi))[0]

    def parts(idx):
        return dict(x=x[idx], mat=mat[idx], used=np.ones(len(idx), np.int32), rho=np.array([M.RHO[m] for m in mat[idx]]), body_id=np.zeros(len(idx), np.int32), bodies={'n': 1})
    exchange = os.environ.get('SLAB_EXCHANGE', 'peer')
    slab = SlabMPMSimulator(q, (0.0, -10.0, 0.0), parts(mine), gid=mine, bounds=bounds, capacity=int(len(mine) * 1.5) + 1000, max_substeps_local=20, device=dev,
                            exchange=exchange, sync=os.environ.get('SLAB_SYNC', 'barrier'))
    slab.sim.fuse_g2p2g = bool(int(os.environ.get('SLAB_FUSE', '0')))   # forward mode only: g2p(f) + p2g(f+1) fused
    st = slab.sim.get_state()
    st['v'][:len(mine)] = v0[mine]; st['F'][:len(mine)] = F0[mine]
    slab.sim.set_state(0, st)
    if os.environ.get('SLAB_MODE', 'forward') == 'backward':
        ok = backward_parity(slab, rank, world, dev, q, parts, x, v0, F0, Ntot, exchange)
        dist.barrier()
        dist.destroy_process_group()
        sys.exit(0 if ok else 1)
    n_steps = 6
    for _ in range(n_steps):
        slab.step()
    got = slab.gather_state()
    migrated = torch.tensor([slab.n_migrated], device=dev); dist.all_reduce(migrated)
    ok = True
    if rank == 0:
        ref = MPMSimulator(dim=3, quality=q, gravity=(0.0, -10.0, 0.0), horizon=1000, max_substeps_local=20, max_substeps_global=10 ** 6, ckpt_dest='gpu', device=dev)
        ref.build(None, None, [], parts(np.arange(Ntot)))
        s0 = ref.get_state(); s0['v'][:] = v0; s0['F'][:] = F0; ref.set_state(0, s0)
        for _ in range(n_steps):
            ref.step(None)
        r = ref.get_state()
        rel = lambda a, b: float(np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-12))
        assert len(got['gid']) == Ntot and np.array_equal(got['gid'], np.arange(Ntot)), 'particles lost or duplicated'
        ex, ev, eF = rel(got['x'], r['x']), rel(got['v'], r['v']), rel(got['F'], r['F'])
        print(f'slab world={world} exchange={exchange} sync={slab.sync} fused={slab.sim.fuse_g2p2g}: migrated={int(migrated.item())} rel err x={ex:.2e} v={ev:.2e} F={eF:.2e}')
        ok = ex < 1e-5 and eF < 1e-5 and ev < 1e-4 and int(migrated.item()) > 0
        print('SLAB_PARITY_OK' if ok else 'SLAB_PARITY_FAIL')
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


def backward_parity(slab, rank, world, dev, q, parts, x, v0, F0, Ntot, exchange):
    from fluidlab_b200 import MPMSimulator
    # 30 substeps on a 20-frame ring: one wrap (chunk checkpoint + re-run in the backward pass), a migration at step 2.  (Round 2, first hardware run: with
    # 50 substeps of this water + ELASTIC cloud at 3 m/s the SINGLE-GPU gradient differs from itself by gC, gF ~ 0.3 when only the summation order changes —
    # amplified round-off, not an exchange error; the yardstick below is therefore that single-GPU spread, measured in the same run.)
    n_steps = int(os.environ.get('SLAB_BWD_STEPS', '3'))
    tgt = torch.from_numpy((x + np.random.RandomState(9).randn(Ntot, 3).astype(np.float32) * 0.05).astype(np.float32)).to(dev)
    slab.enable_grad()
    for _ in range(n_steps):
        slab.step()
    ls = slab.local_state()
    used = ls['used'] != 0
    gx = 2.0 * (ls['x'] - tgt[ls['gid'].long().clamp(min=0)]) * used[:, None]
    slab.set_final_grad(gx)
    for _ in range(n_steps):
        slab.step_grad()
    got = slab.gather_grad()
    migrated = torch.tensor([slab.n_migrated, len(slab._records)], device=dev); dist.all_reduce(migrated)
    ok = True
    if rank == 0:
        ref = MPMSimulator(dim=3, quality=q, gravity=(0.0, -10.0, 0.0), horizon=1000, max_substeps_local=20, max_substeps_global=10 ** 6, ckpt_dest='gpu', device=dev)
        ref.build(None, None, [], parts(np.arange(Ntot)))
        s0 = ref.get_state(); s0['v'][:] = v0; s0['F'][:] = F0; ref.set_state(0, s0)
        ref.enable_grad()
        for _ in range(n_steps):
            ref.step(None)
        xT = ref.get_state()['x']
        ref.reset_grad()
        z9 = np.zeros((Ntot, 3, 3), np.float32)
        ref.set_grad(2.0 * (xT - tgt.cpu().numpy()), np.zeros((Ntot, 3), np.float32), z9, z9)
        for _ in range(n_steps):
            ref.step_grad(None)
        g = ref.get_grad()
        rel = lambda a, b: float(np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-12))
        self_err = None
        if os.environ.get('SLAB_SELFCHECK', '1') == '1':
            # conditioning of the test itself: the same single-GPU gradient again with another summation order (no cell sort, recompute path)
            ref2 = MPMSimulator(dim=3, quality=q, gravity=(0.0, -10.0, 0.0), horizon=1000, max_substeps_local=20, max_substeps_global=10 ** 6, ckpt_dest='gpu', device=dev, sort_every=0)
            ref2.store_grids = False
            ref2.build(None, None, [], parts(np.arange(Ntot)))
            s0 = ref2.get_state(); s0['v'][:] = v0; s0['F'][:] = F0; ref2.set_state(0, s0)
            ref2.enable_grad()
            for _ in range(n_steps):
                ref2.step(None)
            xT2 = ref2.get_state()['x']
            ref2.reset_grad(); ref2.set_grad(2.0 * (xT2 - tgt.cpu().numpy()), np.zeros((Ntot, 3), np.float32), z9, z9)
            for _ in range(n_steps):
                ref2.step_grad(None)
            g2 = ref2.get_grad()
            self_err = {k: rel(g2[k], g[k]) for k in ('x', 'v', 'C', 'F')}
            print('selfcheck single-GPU vs single-GPU (unsorted, recompute path): ' + ' '.join(f'g{k}={self_err[k]:.2e}' for k in ('x', 'v', 'C', 'F')) + f' xT={rel(xT2, xT):.2e}')
        assert len(got['gid']) == Ntot and np.array_equal(got['gid'], np.arange(Ntot)), 'particles lost or duplicated'
        errs = {k: rel(got[k], g[k]) for k in ('x', 'v', 'C', 'F')}
        print(f'slab backward world={world} exchange={exchange}: migrated={int(migrated[0].item())} at {int(migrated[1].item())} (rank, step) pairs; rel err ' +
              ' '.join(f'g{k}={e:.2e}' for k, e in errs.items()))
        # fp32 summation order differs between the sharded and the single-GPU scatter: same bars as the single-GPU adjoint tests
        bars = dict(x=1e-4, v=1e-4, C=2e-3, F=2e-3)
        if self_err is not None:   # never stricter than what the single-GPU path reproduces of itself; and the test must stay discriminating
            bars = {k: max(b, 3.0 * self_err[k]) for k, b in bars.items()}
            assert max(self_err.values()) < 5e-2, f'the scene is too ill-conditioned to test anything: {self_err}'
        ok = all(errs[k] < bars[k] for k in bars) and int(migrated[0].item()) > 0
        print('SLAB_GRAD_OK' if ok else 'SLAB_GRAD_FAIL')
    return ok


if __name__ == '__main__':
    main()
