"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: slab partition, ghost-plane sum exchange, particle migration."""
import os
import socket
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fluidlab_b200.slab import slab_bounds, GhostExchange, migrate, migrate_grad, centre_plane


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, fn, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return dict(ret)


def test_slab_bounds_are_block_aligned():
    b = slab_bounds(32, 224, 8)
    assert b[0] == 32 and b[-1] == 224 and all(x % 8 == 0 for x in b) and all(b[i + 1] - b[i] >= 16 for i in range(8))
    assert slab_bounds(32, 96, 2) == [32, 64, 96]


def _ghost_job(rank, world):
    n = 32
    bounds = slab_bounds(0, 32, world)  # [0, 16, 32]
    g = torch.Generator().manual_seed(100 + rank)
    # each rank's local p2g result: non-zero only on its own slab +- 3 planes
    grid = torch.zeros((n ** 3, 4))
    lo, hi = max(bounds[rank] - 3, 0), min(bounds[rank + 1] + 3, n)
    grid[lo * n * n:hi * n * n] = torch.rand(((hi - lo) * n * n, 4), generator=g)
    local = grid.clone()
    ex = GhostExchange(n, bounds, rank, world, halo=8)
    ex.exchange_sum(grid)
    flags = torch.zeros((n // 8) ** 3, dtype=torch.int32)
    ex.flag_ghost_blocks(flags)
    return local.numpy(), grid.numpy(), flags.numpy(), ex.bytes_per_exchange()


def test_ghost_exchange_gives_both_ranks_the_global_sum_on_the_ghost_region():
    out = _run(_ghost_job)
    n = 32
    total = out[0][0] + out[1][0]
    reg = slice(8 * n * n, 24 * n * n)  # planes [16-8, 16+8)
    for r in (0, 1):
        assert np.allclose(out[r][1][reg], total[reg])
        keep = np.ones(n ** 3, bool); keep[reg] = False
        assert np.array_equal(out[r][1][keep], out[r][0][keep])  # nothing else touched
        f = out[r][2].reshape(4, 4, 4)
        assert f[1:3].all() and not f[0].any() and not f[3].any()
        assert out[r][3] == 16 * n * n * 16
    # every plane a rank's particles can read (own slab +- 3) now holds the global sum
    assert np.allclose(out[0][1][:19 * n * n], total[:19 * n * n])
    assert np.allclose(out[1][1][13 * n * n:], total[13 * n * n:])


def _migrate_job(rank, world):
    n = 32
    inv_dx = float(n)
    bounds = [0, 16, 32]
    rng = np.random.RandomState(7 + rank)
    N = 64
    x = rng.uniform(0.1, 0.9, size=(N, 3)).astype(np.float32)
    used = np.ones(N, np.int32); used[48:] = 0; x[48:] = -100.0
    st = dict(x=torch.from_numpy(x.copy()), v=torch.from_numpy(rng.randn(N, 3).astype(np.float32)), C=torch.from_numpy(rng.randn(N, 3, 3).astype(np.float32)),
              F=torch.from_numpy(rng.randn(N, 3, 3).astype(np.float32)), used=torch.from_numpy(used.copy()),
              mrow=torch.from_numpy(np.arange(N, dtype=np.int32) % 3), gid=torch.from_numpy(np.arange(N, dtype=np.int32) + 1000 * rank))
    before = {k: v.clone().numpy() for k, v in st.items()}
    n_out, n_in = migrate(st, bounds[rank], bounds[rank + 1], rank, world, inv_dx)
    return before, {k: v.numpy() for k, v in st.items()}, n_out, n_in


def test_migration_moves_leavers_and_conserves_particles():
    out = _run(_migrate_job)
    n, inv_dx = 32, 32.0
    allb = {}
    for r in (0, 1):
        b = out[r][0]
        for i in np.where(b['used'] != 0)[0]:
            allb[int(b['gid'][i])] = (b['x'][i], b['v'][i], b['F'][i], int(b['mrow'][i]))
    seen = {}
    for r in (0, 1):
        a = out[r][1]
        cp = centre_plane(torch.from_numpy(a['x']), inv_dx).numpy()
        for i in np.where(a['used'] != 0)[0]:
            g = int(a['gid'][i])
            assert g not in seen, 'duplicated particle'
            seen[g] = r
            lo, hi = (0, 16) if r == 0 else (16, 32)
            assert (cp[i] >= lo or r == 0) and (cp[i] < hi or r == 1), 'particle on the wrong rank after migration'
            x0, v0, F0, m0 = allb[g]
            assert np.array_equal(a['x'][i], x0) and np.array_equal(a['v'][i], v0) and np.array_equal(a['F'][i], F0) and int(a['mrow'][i]) == m0
    assert set(seen) == set(allb)
    assert out[0][2] == out[1][3] and out[1][2] == out[0][3] and out[0][2] + out[1][2] > 0


def _migrate_adjoint_job(rank, world):
    """<M u, w> on the used slots after the migration == <u, M^T w> on the used slots before it (summed over ranks)"""
    n, inv_dx, bounds = 32, 32.0, [0, 16, 32]
    rng = np.random.RandomState(17 + rank)
    N = 64
    x = rng.uniform(0.1, 0.9, size=(N, 3)).astype(np.float32)
    used = np.ones(N, np.int32); used[48:] = 0; x[48:] = -100.0
    f32 = lambda a: torch.from_numpy(a.astype(np.float32))
    st = dict(x=torch.from_numpy(x.copy()), v=f32(rng.randn(N, 3)), C=f32(rng.randn(N, 3, 3)), F=f32(rng.randn(N, 3, 3)), used=torch.from_numpy(used.copy()),
              mrow=torch.from_numpy(np.arange(N, dtype=np.int32) % 3), gid=torch.from_numpy(np.arange(N, dtype=np.int32) + 1000 * rank))
    u = {k: st[k].clone().double() for k in ('x', 'v', 'C', 'F')}
    used_before = st['used'].clone()
    rec = {}
    n_out, n_in = migrate(st, bounds[rank], bounds[rank + 1], rank, world, inv_dx, record=rec)
    used_after = st['used'] != 0
    w = {k: f32(rng.randn(*st[k].shape)) for k in ('x', 'v', 'C', 'F')}
    for k in w:
        w[k][~used_after] = 0
    lhs = sum((st[k].double()[used_after] * w[k].double()[used_after]).sum() for k in w)
    g = {k: w[k].clone() for k in w}
    migrate_grad(g, rec)
    ub = used_before != 0
    assert all(float(g[k][~ub].abs().max()) == 0.0 for k in g), 'adjoint leaked into a slot that was parked before the migration'
    rhs = sum((u[k][ub] * g[k].double()[ub]).sum() for k in g)
    t = torch.tensor([float(lhs), float(rhs)], dtype=torch.float64)
    dist.all_reduce(t)
    return float(t[0]), float(t[1]), n_out, n_in


def test_migrate_grad_is_the_adjoint_of_migrate():
    out = _run(_migrate_adjoint_job)
    lhs, rhs = out[0][0], out[0][1]
    assert out[0][2] + out[1][2] > 0, 'nothing migrated'
    assert abs(lhs - rhs) < 1e-9 * max(1.0, abs(lhs)), (lhs, rhs)
