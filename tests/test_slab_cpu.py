"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: slab partition, ghost-plane sum exchange, particle migration."""
import os
import socket
import numpy as np
import pytest
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


# ----------------------------------------------------------------------------------------------------------------------------------------
# The slab ORCHESTRATION of the product (fluidlab_b200/slab.py: SlabMPMSimulator.step / step_grad, ghost sums, migration and its adjoint)
# driven on CPU: the local simulator is replaced by a stand-in with MPMSimulator's step-level methods and slab_* hooks whose kernels are
# the oracle's phases.  The sharded result (2 ranks, gloo) must equal the single-domain oracle, forward and backward.
# ----------------------------------------------------------------------------------------------------------------------------------------
class OracleLocalSim:
    """stand-in for fluidlab_b200.simulator.MPMSimulator inside SlabMPMSimulator (test infrastructure: wraps oracle.OracleSim)"""

    def __init__(self, quality, gravity, particles, boundary, max_substeps_local):
        from oracle import oracle as orc
        from conftest import make_particles
        self.n_grid = int(round(64 * quality)); self.inv_dx = float(self.n_grid)
        self.n_substeps = 10; self.max_substeps_local = max_substeps_local
        self.device = torch.device('cpu')
        self.N = len(particles['x'])
        self._make = lambda mat: make_particles(np.zeros((len(mat), 3)), mat, self.n_grid)
        P = make_particles(particles['x'], particles['mat'], self.n_grid, used=particles['used'])
        self.o = orc.OracleSim(self.n_grid, P, gravity=gravity, boundary=boundary, precision=32, max_substeps_local=max_substeps_local)
        self._mrow = torch.from_numpy(np.asarray(particles['mat'], dtype=np.int32).copy())   # "material row" = the material id itself
        self._layouts = [(0, self._mrow.clone())]   # (first frame, material per slot): migrations change what a slot holds
        self.cur_substep_global = 0
        self.grad_enabled = False
        self._acc = None; self._adj = None; self._gframe = 0

    # indices
    @property
    def cur_substep_local(self): return self.cur_substep_global % self.max_substeps_local
    @property
    def cur_step_global(self): return self.cur_substep_global // self.n_substeps
    def enable_grad(self): self.grad_enabled = True; self.cur_substep_global = 0
    def memory_to_cache(self): self.o.L.orc_copy_frame(self.o.h, self.max_substeps_local, 0)
    def copy_frame(self, a, b): self.o.L.orc_copy_frame(self.o.h, a, b)
    def slab_snapshot_frame(self, f):
        return dict(frame=self.o.get_frame(f), mrow=self._mrow.clone(), layouts=[(f0, m.clone()) for f0, m in self._layouts])
    def slab_restore_frame(self, f, snap):
        fr = snap['frame']
        self.o.set_frame(f, fr['x'], fr['v'], fr['C'], fr['F'], fr['used'])
        self._mrow.copy_(snap['mrow']); self._layouts = [(0, snap['mrow'].clone())]
        self._use_layout(f)

    def _use_layout(self, f):
        mrow = [m for f0, m in self._layouts if f0 <= f][-1]
        P = self._make(mrow.numpy())
        from oracle.oracle import _p, _d
        self.o.L.orc_set_particle_info(self.o.h, _p(P['mat']), _p(P['cls']), _p(_d(P['mu'])), _p(_d(P['lam'])), _p(_d(P['mass'])))

    # frames
    def sort_frame(self, f): pass
    def readframe_torch(self, f, want=('x', 'v', 'C', 'F', 'used')):
        fr = self.o.get_frame(f)
        out = {k: torch.from_numpy(fr[k].astype(np.float32)) for k in ('x', 'v', 'C', 'F')}
        out['used'] = torch.from_numpy(fr['used'].astype(np.int32))
        return {k: out[k] for k in want}
    def setframe(self, f, x, v, Cm, F, used):
        self.o.set_frame(f, x.numpy(), v.numpy(), Cm.numpy(), F.numpy(), used.numpy())
        self._layouts = [(f0, m) for f0, m in self._layouts if f0 < f] + [(f, self._mrow.clone())]
        self._use_layout(f)
    def slab_positions(self, f):
        fr = self.o.get_frame(f)
        return torch.from_numpy(fr['x'][:, 0].astype(np.float32)), torch.from_numpy(fr['used'] != 0)

    # forward phases
    def phase(self, name, f, *a):
        L, h = self.o.L, self.o.h
        if name == 'p2g':
            L.orc_phase_reset_grid(h); L.orc_phase_p2g(h, f, int(a[0]) if a else 1)
            vin, m, _ = self.o.get_grid()
            self._acc = torch.from_numpy(np.concatenate([vin, m[:, None]], 1).astype(np.float32))
        elif name == 'grid_op': L.orc_phase_grid_op(h, f)
        elif name == 'g2p': L.orc_phase_g2p(h, f)
        else: raise KeyError(name)
    def slab_grid_acc(self, f): return self._acc
    def slab_grid_acc_commit(self, f, acc): self.o.set_grid(acc[:, :3].numpy(), acc[:, 3].numpy())
    def slab_flag_blocks(self, f, flagger): flagger(torch.zeros((self.n_grid // 8) ** 3, dtype=torch.int32))

    # backward
    def reset_grad(self):
        self.o.L.orc_reset_grad(self.o.h); self._gframe = self.cur_substep_local
    def write_grad_torch(self, g):
        self.o.set_grad_frame(self._gframe, g['x'].numpy(), g['v'].numpy(), g['C'].numpy(), g['F'].numpy())
    def read_grad_torch(self):
        g = self.o.get_grad_frame(self._gframe)
        return {k: torch.from_numpy(g[k].astype(np.float32)) for k in ('x', 'v', 'C', 'F')}
    def slab_adjoint_moves_to_frame(self, f):   # the oracle keeps one adjoint per ring frame (like the reference): MPM:858-860
        self.o.L.orc_copy_grad(self.o.h, 0, f); self.o.L.orc_reset_grad_till(self.o.h, f); self._gframe = f
    def slab_substep_grad_p2g(self, f):
        assert self._gframe == f + 1
        self._use_layout(f)
        z3, z9 = np.zeros((self.N, 3)), np.zeros((self.N, 3, 3))
        self.o.set_grad_frame(f, z3, z3, z9, z9)
        self.phase('p2g', f, 0)
    def slab_substep_grad_scatter(self, f):
        L, h = self.o.L, self.o.h
        L.orc_phase_grid_op(h, f); L.orc_phase_g2p_grad(h, f)
        _, _, gvout = self.o.get_grid_grad()
        self._adj = torch.from_numpy(gvout.astype(np.float32))
    def slab_grid_adj(self, f): return self._adj
    def slab_grid_adj_commit(self, f, adj):
        G = self.n_grid ** 3
        self.o.set_grid_grad(np.zeros((G, 3)), np.zeros(G), adj.numpy())
    def slab_substep_grad_finish(self, f):
        L, h = self.o.L, self.o.h
        L.orc_phase_grid_op_grad(h, f); L.orc_phase_p2g_grad(h, f); L.orc_phase_unused_grad(h, f)
        self._gframe = f


_SLAB_STEPS = 5


def _slab_scene():
    from fluidlab_b200 import macros as M
    rng = np.random.RandomState(11)
    n, Ntot = 32, 1500
    x = rng.uniform((0.30, 0.35, 0.35), (0.70, 0.55, 0.65), size=(Ntot, 3)).astype(np.float32)
    # a fast stream along +x in the lower half and along -x in the upper half: particles cross the slab boundary in both directions
    v0 = np.where(x[:, 1:2] < 0.45, np.array([[6.0, 0.0, 0.5]]), np.array([[-6.0, 0.3, 0.0]])).astype(np.float32)
    v0 += (rng.randn(Ntot, 3) * 0.2).astype(np.float32)
    mat = np.where(x[:, 2] < 0.5, M.WATER, M.ELASTIC).astype(np.int32)
    tgt = (x + rng.randn(Ntot, 3) * 0.05).astype(np.float32)
    return n, x, v0, mat, tgt


def _slab_fwd_bwd_job(rank, world, T=60):
    from fluidlab_b200 import macros as M
    from fluidlab_b200.slab import SlabMPMSimulator
    torch.set_num_threads(1)
    n, x, v0, mat, tgt = _slab_scene()
    bounds = slab_bounds(0, 32, world)
    cp = centre_plane(torch.from_numpy(x), float(n)).numpy()
    mine = np.where((cp >= bounds[rank]) & (cp < bounds[rank + 1]))[0]
    parts = dict(x=x[mine], mat=mat[mine], used=np.ones(len(mine), np.int32), rho=np.array([M.RHO[m] for m in mat[mine]]), body_id=np.zeros(len(mine), np.int32),
                 bodies={'n': 1})
    slab = SlabMPMSimulator(0.5, (0.0, -10.0, 0.0), parts, gid=mine, bounds=bounds, capacity=len(mine) + 400, max_substeps_local=T, halo=4,
                            exchange='nccl', sim_factory=OracleLocalSim)
    st = slab.sim.readframe_torch(0)
    st['v'][:len(mine)] = torch.from_numpy(v0[mine])
    slab.sim.setframe(0, st['x'], st['v'], st['C'], st['F'], st['used'])
    slab.enable_grad()
    n_steps = _SLAB_STEPS
    for _ in range(n_steps):
        slab.step()
    fwd = slab.gather_state()
    ls = slab.local_state()
    used = (ls['used'] != 0)
    gid = ls['gid'].long().clamp(min=0)
    gx = 2.0 * (ls['x'] - torch.from_numpy(tgt)[gid]) * used[:, None]
    slab.set_final_grad(gx.float())
    for _ in range(n_steps):
        slab.step_grad()
    grad = slab.gather_grad()
    return fwd, grad, slab.n_migrated, sorted(slab._records)


def _slab_fwd_bwd_job_short_ring(rank, world):
    return _slab_fwd_bwd_job(rank, world, T=20)


@pytest.mark.parametrize('job', [_slab_fwd_bwd_job, _slab_fwd_bwd_job_short_ring], ids=['one_chunk', 'ring_of_2_steps'])
def test_slab_orchestration_forward_and_backward_match_the_single_domain_oracle(job):
    """2 ranks over gloo: SlabMPMSimulator.step x5 (ghost sums, migration in both directions at two step boundaries) then step_grad x5
    (ghost sums of the accumulator AND of the v_out adjoint, migrate_grad) == the single-domain oracle's states and
    dLoss/d(x0, v0, C0, F0).  `ring_of_2_steps`: the ring holds 2 steps, so the 5-step trajectory wraps twice — every chunk start is
    checkpointed and the backward pass re-runs each chunk (ghost sums and migrations included) before walking it.  The sharded run computes in fp32; it must be as close to the fp64 single-domain result as the fp32
    single-domain run is (x3 slack), i.e. sharding adds nothing beyond summation-order noise."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle import oracle as orc
    from conftest import make_particles
    out = _run(job)
    n, x, v0, mat, tgt = _slab_scene()
    N, S = len(x), _SLAB_STEPS
    ref = {}
    for prec in (32, 64):
        o = orc.OracleSim(n, make_particles(x, mat, n), gravity=(0.0, -10.0, 0.0), precision=prec, max_substeps_local=60)
        o.set_frame(0, x, v0, np.zeros((N, 3, 3)), np.tile(np.eye(3), (N, 1, 1)), np.ones(N, np.int32))
        o.enable_grad()
        for _ in range(S):
            o.step(None)
        fr = o.get_frame(10 * S)
        o.reset_grad()
        z9 = np.zeros((N, 3, 3))
        o.set_grad_frame(10 * S, 2.0 * (fr['x'] - tgt), np.zeros((N, 3)), z9, z9)
        for _ in range(S):
            o.step_grad(None)
        ref[prec] = (fr, o.get_grad_frame(0))
    rel = lambda a, b: float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-12))
    fr64, g64 = ref[64]
    fr32, g32 = ref[32]
    for r in (0, 1):
        fwd, grad, n_mig, rec_steps = out[r]
        assert np.array_equal(fwd['gid'], np.arange(N)) and np.array_equal(grad['gid'], np.arange(N)), 'particles lost or duplicated'
        for k in ('x', 'v', 'F'):
            assert rel(fwd[k], fr64[k]) < max(3 * rel(fr32[k], fr64[k]), 1e-6), (k, rel(fwd[k], fr64[k]), rel(fr32[k], fr64[k]))
        for k in ('x', 'v', 'C', 'F'):
            assert np.abs(g64[k]).max() > 0
            assert rel(grad[k], g64[k]) < max(3 * rel(g32[k], g64[k]), 1e-6), (k, rel(grad[k], g64[k]), rel(g32[k], g64[k]))
        assert rel(grad['x'], g64['x']) < 1e-4 and rel(grad['v'], g64['v']) < 1e-4
    assert out[0][2] > 0 and out[1][2] > 0, 'the scene must migrate particles in both directions'
    assert len(out[0][3]) >= 2, 'migrations must happen at more than one step boundary'
