"""The WHOLE product on CPU: fluidlab_b200's Python host driving its real .cu translation units, compiled by g++ against a model of the
CUDA execution model (tests/cuda_emu/: one host thread per CUDA thread; __syncthreads, shared memory, warp collectives, atomics, the
time-blocked tiles ... behave as on the device), with CPU tensors in place of HBM.

Why: the build container has no GPU and GPU minutes are rationed, so this is where kernel logic and host sequencing get checked on every
`pytest -m "not gpu"` run — against the same oracle the GPU parity tests use.  It says nothing about performance or about hardware
behaviour beyond the programming model; tests/test_gpu_parity.py and tests/test_zz_smoke_gpu.py remain the parity gate on a B200."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'cuda_emu'))
import harness  # noqa: E402

from conftest import make_particles  # noqa: E402
from oracle import oracle as orc  # noqa: E402


@pytest.fixture
def emu():
    L = harness.enable()
    yield L
    harness.disable()


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-12))


def test_product_fails_loudly_without_cuda_when_the_emulation_is_off():
    from fluidlab_b200 import MPMSimulator
    assert not harness._state
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        MPMSimulator(dim=3, quality=0.25, gravity=(0, -10, 0), horizon=10, max_substeps_local=20, max_substeps_global=1000, ckpt_dest='cpu')


def test_mpm_kernels_forward_and_backward_match_the_fp64_oracle(emu):
    """k_p2g (register sliding-window scatter, warp-local key ranking), k_grid_op (sparse blocks, cube walls), k_g2p, the cell sort, the
    stored-grid backward (k_g2p_grad_scatter, k_grid_op_grad, k_particle_grad with the SVD adjoint) on water + elastic + plasto-elastic
    particles: one step (10 substeps) forward and backward through MPMSimulator.step / step_grad."""
    from fluidlab_b200 import MPMSimulator, macros as M
    rng = np.random.RandomState(3)
    n, N = 16, 300
    x = rng.uniform(0.35, 0.65, size=(N, 3)).astype(np.float32)
    mat = np.array([[M.WATER, M.ELASTIC, M.ICECREAM][i % 3] for i in range(N)], dtype=np.int32)
    v0 = (rng.randn(N, 3) * 0.5).astype(np.float32)
    # F0 away from the identity: with equal singular values the SVD adjoint sits on its 1e-8 clamp and fp32 results split into branches
    F0 = (np.eye(3)[None] + rng.randn(N, 3, 3) * 0.05).astype(np.float32); C0 = (rng.randn(N, 3, 3) * 2.0).astype(np.float32)
    P = make_particles(x, mat, n)
    bnd = dict(type='cube', lower=(0.3, 0.3, 0.3), upper=(0.7, 0.7, 0.7))
    w = rng.randn(N, 3).astype(np.float32)     # one seed direction for every configuration (the fp32 error depends on it, not on the path)
    for store, fuse in ((True, False), (False, False), (True, True)):   # last: grad-mode g2p2g fusion (fmpm_substeps_fused_store)
        s = MPMSimulator(dim=3, quality=n / 64, gravity=(0, -10, 0), horizon=50, max_substeps_local=20, max_substeps_global=1000, ckpt_dest='cpu', device='cpu')
        s.use_graphs, s.store_grids, s.fuse_g2p2g = False, store, fuse
        s.setup_boundary(**bnd)
        s.build(None, None, [], P)
        st = s.get_state(); st['v'][:] = v0; st['F'][:] = F0; st['C'][:] = C0; s.set_state(0, st)
        s.enable_grad()
        assert s._can_fuse() == fuse
        s.step(None)
        fr = s.get_state()
        z9 = np.zeros((N, 3, 3), np.float32)
        s.reset_grad(); s.set_grad(w, np.zeros((N, 3), np.float32), z9, z9)
        s.step_grad(None)
        g = s.get_grad()
        o = orc.OracleSim(n, P, gravity=(0, -10, 0), boundary=bnd, precision=64, max_substeps_local=20)
        o.set_frame(0, x, v0, C0, F0, np.ones(N, np.int32))
        o.enable_grad(); o.step(None)
        ofr = o.get_frame(10)
        o.reset_grad(); o.set_grad_frame(10, w, np.zeros((N, 3)), z9, z9); o.step_grad(None)
        og = o.get_grad_frame(0)
        assert rel(fr['x'], ofr['x']) < 1e-6 and rel(fr['F'], ofr['F']) < 1e-5 and rel(fr['v'], ofr['v']) < 1e-4, {k: rel(fr[k], ofr[k]) for k in 'xvCF'}
        for k in 'xvCF':
            assert rel(g[k], og[k]) < 1e-4, (store, fuse, k, rel(g[k], og[k]))


def test_circulation_stack_on_the_emulated_device(emu):
    """AgentCirculation + AirCon + SmokeField + CirculationLoss + parked MPM particles through TaichiEnv (reduced: 24^3 smoke grid, 10 sweeps)"""
    from circulation_case import run_circulation_stack
    dets = [[5, 16], [7, 16], [3, 16], [5, 14], [5, 18], [5, 8], [7, 8], [3, 8], [5, 6], [5, 10], [20, 12], [21, 12], [18, 12], [20, 9], [20, 16]]
    run_circulation_stack(device='cpu', res=24, iters=10, band=(8, 14), detectors=dets, detector_h=11, n_steps=3, max_substeps_local=40)
    # a ring of 2 steps for a 3-step trajectory: MPM ring and smoke ring wrap, the backward pass re-runs the first chunk (smoke steps included)
    run_circulation_stack(device='cpu', res=24, iters=10, band=(8, 14), detectors=dets, detector_h=11, n_steps=3, max_substeps_local=20, ring_wraps=True)


# ---------------------------------------------------------------------------------------------------------------- x-slabs, 2 ranks (gloo)
def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


class ShmPeers:
    """stand-in for fluidlab_b200.slab.SymmetricMemoryPeers on the emulated device: every rank's buffer is a POSIX shared-memory file that
    the other rank processes map too, so the kernels' peer pointers (vector reductions into the neighbour's grids, block flags) address
    real shared memory across processes; the barrier is a gloo barrier (the emulated kernels are synchronous)."""
    _count = 0

    def __init__(self, group, device):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.tag = os.environ['MASTER_PORT']
        self._keep = []

    def alloc(self, shape, dtype):
        idx = ShmPeers._count; ShmPeers._count += 1
        numel = int(np.prod(shape))
        path = lambda r: f'/dev/shm/fmpm_emu_{self.tag}_{idx}_{r}'
        mine = torch.from_file(path(self.rank), shared=True, size=numel, dtype=dtype)
        mine.zero_()
        self.dist.barrier(group=self.group)
        maps = [mine if r == self.rank else torch.from_file(path(r), shared=True, size=numel, dtype=dtype) for r in range(self.world)]
        self._keep += maps
        self.dist.barrier(group=self.group)
        os.unlink(path(self.rank))   # the mappings stay valid; nothing is left behind in /dev/shm
        return mine.view(*shape), [m.data_ptr() for m in maps]

    def barrier(self):
        self.dist.barrier(group=self.group)


def _slab_worker(rank, world, port, ret, exchange='nccl', fused=False, sync='barrier', pull=True, with_static=False):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        harness.enable()
        from fluidlab_b200 import MPMSimulator, macros as M
        from fluidlab_b200.slab import SlabMPMSimulator, slab_bounds, centre_plane
        os.environ['FMPM_SLAB_PULL'] = '1' if pull else '0'
        rng = np.random.RandomState(21)
        n, Ntot, n_steps = 16 * world, 700, 5          # two ranks: 32^3; three ranks (a middle slab with two neighbours): 48^3
        quality = n / 64
        x = rng.uniform((0.32, 0.36, 0.36) if world == 2 else (0.22, 0.36, 0.36), (0.68, 0.54, 0.64) if world == 2 else (0.78, 0.54, 0.64), size=(Ntot, 3)).astype(np.float32)
        v0 = np.where(x[:, 1:2] < 0.45, np.array([[6.0, 0.0, 0.5]]), np.array([[-6.0, 0.3, 0.0]])).astype(np.float32) + (rng.randn(Ntot, 3) * 0.2).astype(np.float32)
        mat = np.where(x[:, 2] < 0.5, M.WATER, M.ELASTIC).astype(np.int32)
        F0 = (np.eye(3)[None] + rng.randn(Ntot, 3, 3) * 0.04).astype(np.float32)
        tgt = torch.from_numpy((x + rng.randn(Ntot, 3) * 0.05).astype(np.float32))
        bounds = slab_bounds(0, n, world)
        cp = centre_plane(torch.from_numpy(x), float(n)).numpy()
        mine = np.where((cp >= bounds[rank]) & (cp < bounds[rank + 1]))[0]

        def parts(idx):
            return dict(x=x[idx], mat=mat[idx], used=np.ones(len(idx), np.int32), rho=np.array([M.RHO[m] for m in mat[idx]]), body_id=np.zeros(len(idx), np.int32), bodies={'n': 1})

        def statics():   # a static box collider (meshes/static.py) across the slab boundary, in the way of the falling cloud
            if not with_static:
                return []
            from conftest import box_sdf
            from fluidlab_b200 import Statics
            bv, bT = box_sdf((0.3, 0.04, 0.3), 0.4)
            s = Statics()
            s.add_static(file='box.obj', material=M.CUP, has_dynamics=True, pos=(0.5, 0.36, 0.5), sdf=dict(voxels=bv, T_mesh_to_voxels=bT))
            return s
        slab = SlabMPMSimulator(quality, (0.0, -10.0, 0.0), parts(mine), gid=mine, bounds=bounds, capacity=len(mine) + 300, max_substeps_local=20, device='cpu', halo=4,
                                exchange=exchange, peer_factory=ShmPeers, sync=sync, statics=statics())

        def static_effect(x_with):   # how far the single-domain run WITHOUT the collider ends from the one with it (rank 0, with_static only)
            if not with_static:
                return 0.0
            r0 = MPMSimulator(dim=3, quality=quality, gravity=(0.0, -10.0, 0.0), horizon=50, max_substeps_local=20, max_substeps_global=1000, ckpt_dest='cpu', device='cpu')
            r0.use_graphs = False
            r0.build(None, None, [], parts(np.arange(Ntot)))
            s0 = r0.get_state(); s0['v'][:] = v0; s0['F'][:] = F0; r0.set_state(0, s0)
            for _ in range(n_steps):
                r0.step(None)
            return float(np.abs(r0.get_state()['x'] - x_with).max())
        assert slab.exchange == exchange and (slab.sync == sync or exchange != 'peer')
        assert slab.pull == (exchange == 'peer' and sync == 'signal' and pull)   # pull form of the ghost reduction: the one-call forward steps
        slab.sim.use_graphs = False
        st = slab.sim.get_state()
        st['v'][:len(mine)] = v0[mine]; st['F'][:len(mine)] = F0[mine]
        slab.sim.set_state(0, st)
        if fused:   # forward-only: g2p(f) + p2g(f+1) fused, the scatter half reducing into the neighbour's accumulator of parity f+1
            slab.sim.fuse_g2p2g = True
            for _ in range(n_steps):
                slab.step()
            out = dict(fwd=slab.gather_state(), migrated=slab.n_migrated)
            assert not slab.sync_error()
            if exchange == 'peer':   # both parity accumulators and their block flags are clean between steps, ghost blocks included (pull form: deferred clears)
                dist.barrier()
                assert float(slab.sim._grid_pm.abs().max()) == 0.0 and int(slab.sim._blk_flags.abs().max()) == 0
            if rank == 0:
                ref = MPMSimulator(dim=3, quality=quality, gravity=(0.0, -10.0, 0.0), horizon=50, max_substeps_local=20, max_substeps_global=1000, ckpt_dest='cpu', device='cpu')
                ref.use_graphs = False
                ref.build(None, None, statics(), parts(np.arange(Ntot)))
                s0 = ref.get_state(); s0['v'][:] = v0; s0['F'][:] = F0; ref.set_state(0, s0)
                for _ in range(n_steps):
                    ref.step(None)
                r = ref.get_state()
                out.update(ref_state={k: r[k] for k in ('x', 'v', 'F')}, static_effect=static_effect(r['x']))
            ret[rank] = out
            return
        slab.enable_grad()
        for _ in range(n_steps):
            slab.step()
        fwd = slab.gather_state()
        ls = slab.local_state()
        used = ls['used'] != 0
        gx = 2.0 * (ls['x'] - tgt[ls['gid'].long().clamp(min=0)]) * used[:, None]
        slab.set_final_grad(gx.clone())
        for _ in range(n_steps):
            slab.step_grad()
        grad = slab.gather_grad()
        assert not slab.sync_error()
        out = dict(fwd=fwd, grad=grad, migrated=slab.n_migrated, rec=sorted(slab._records))
        if rank == 0:   # the single-domain reference: the same product on the same emulated device
            ref = MPMSimulator(dim=3, quality=quality, gravity=(0.0, -10.0, 0.0), horizon=50, max_substeps_local=20, max_substeps_global=1000, ckpt_dest='cpu', device='cpu')
            ref.use_graphs = False
            ref.build(None, None, statics(), parts(np.arange(Ntot)))
            s0 = ref.get_state(); s0['v'][:] = v0; s0['F'][:] = F0; ref.set_state(0, s0)
            ref.enable_grad()
            for _ in range(n_steps):
                ref.step(None)
            r = ref.get_state()
            ref.reset_grad()
            z9 = np.zeros((Ntot, 3, 3), np.float32)
            ref.set_grad(2.0 * (r['x'] - tgt.numpy()), np.zeros((Ntot, 3), np.float32), z9, z9)
            for _ in range(n_steps):
                ref.step_grad(None)
            out.update(ref_state={k: r[k] for k in ('x', 'v', 'F')}, ref_grad=ref.get_grad(), static_effect=static_effect(r['x']))
        ret[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('exchange', ['nccl', 'peer', 'peer-signal'])
def test_slab_sharded_forward_and_backward_match_the_single_domain_run_on_the_emulated_device(exchange):
    """the CUDA leg of the x-slab path that tests/run_slab_gpu.py exercises on 2 GPUs, here on 2 gloo ranks with the emulated device:
    SlabMPMSimulator.step (ghost all-reduce of the accumulator, migration) and step_grad (fmpm_p2g(write_F=0) -> ghost sum ->
    fmpm_substep_grad_scatter -> ghost sum of the v_out adjoint -> fmpm_substep_grad_finish, migrate_grad) against MPMSimulator on one
    domain: states and dL/d(x0, v0, C0, F0).  The ring holds 2 steps and the trajectory has 5: both sides wrap twice, the sharded side
    checkpoints every chunk start and re-runs each chunk (exchanges and migrations included) during the backward pass.
    `peer`: the fused path of the product — p2g's and g2p.grad's vector reductions for nodes on shared planes go straight into the
    neighbour's grids (here: POSIX shared memory across the two rank processes instead of NVLink peer memory), no all-reduce anywhere."""
    import torch.multiprocessing as mp
    mgr = mp.Manager(); ret = mgr.dict()
    sync = 'signal' if exchange.endswith('-signal') else 'barrier'   # 'signal': neighbour handshakes inside the library instead of a barrier over all ranks
    mp.spawn(_slab_worker, args=(2, _free_port(), ret, exchange.split('-')[0], False, sync), nprocs=2, join=True)
    out = dict(ret)
    N = 700
    ref_s, ref_g = out[0]['ref_state'], out[0]['ref_grad']
    for r in (0, 1):
        fwd, grad = out[r]['fwd'], out[r]['grad']
        assert np.array_equal(fwd['gid'], np.arange(N)) and np.array_equal(grad['gid'], np.arange(N)), 'particles lost or duplicated'
        assert rel(fwd['x'], ref_s['x']) < 1e-5 and rel(fwd['F'], ref_s['F']) < 1e-5 and rel(fwd['v'], ref_s['v']) < 1e-4
        errs = {k: rel(grad[k], ref_g[k].astype(np.float64)) for k in 'xvCF'}
        assert errs['x'] < 1e-4 and errs['v'] < 1e-4 and errs['C'] < 2e-3 and errs['F'] < 2e-3, errs
    assert out[0]['migrated'] > 0 and out[1]['migrated'] > 0 and len(out[0]['rec']) >= 2, (out[0]['migrated'], out[1]['migrated'], out[0]['rec'])


def test_g2p2g_fused_substeps_equal_the_unfused_path(emu):
    """fmpm_substeps_fused (p2g, [grid_op, g2p2g] x 9, grid_op, g2p: the inner gather / scatter pairs in ONE kernel, v and C never leaving
    registers) against the plain p2g / grid_op / g2p substeps and against the fp64 oracle: water + elastic + plasto-elastic particles, 12 %
    unused slots, cube walls, two steps (so the second starts from the complete frame the first one's final g2p wrote)."""
    from fluidlab_b200 import MPMSimulator, macros as M
    rng = np.random.RandomState(7)
    n, N = 16, 330
    x = rng.uniform(0.34, 0.66, size=(N, 3)).astype(np.float32)
    mat = np.array([[M.WATER, M.ELASTIC, M.ICECREAM][i % 3] for i in range(N)], dtype=np.int32)
    used = (rng.rand(N) > 0.12).astype(np.int32)
    v0 = (rng.randn(N, 3) * 0.8).astype(np.float32)
    F0 = (np.eye(3)[None] + rng.randn(N, 3, 3) * 0.05).astype(np.float32); C0 = (rng.randn(N, 3, 3) * 2.0).astype(np.float32)
    P = make_particles(x, mat, n, used=used)
    bnd = dict(type='cube', lower=(0.3, 0.3, 0.3), upper=(0.7, 0.7, 0.7))
    out = {}
    for fuse in (False, True, 'unsorted'):   # 'unsorted': no cell sort at all — the scatter's warp-local key ranking carries the whole burden
        s = MPMSimulator(dim=3, quality=n / 64, gravity=(0.3, -10, 0), horizon=50, max_substeps_local=40, max_substeps_global=1000, ckpt_dest='cpu', device='cpu',
                         sort_every=0 if fuse == 'unsorted' else 1)
        s.use_graphs, s.fuse_g2p2g = False, bool(fuse)
        s.setup_boundary(**bnd)
        s.build(None, None, [], P)
        st = s.get_state(); st['v'][:] = v0; st['F'][:] = F0; st['C'][:] = C0; s.set_state(0, st)
        assert s._can_fuse() == bool(fuse)
        s.step(None); s.step(None)
        out[fuse] = s.get_state()
    o = orc.OracleSim(n, P, gravity=(0.3, -10, 0), boundary=bnd, precision=64, max_substeps_local=40)
    o.set_frame(0, x, v0, C0, F0, used)
    o.step(None); o.step(None)
    ofr = o.get_frame(20)
    u = used != 0
    assert np.array_equal(out[True]['used'], used) and np.array_equal(out[False]['used'], used)
    for k, bar in (('x', 1e-6), ('F', 1e-5), ('v', 1e-4), ('C', 1e-3)):
        assert rel(out[True][k][u], out[False][k][u].astype(np.float64)) < bar, (k, rel(out[True][k][u], out[False][k][u].astype(np.float64)))
        assert rel(out[True][k][u], ofr[k][u]) < bar, (k, rel(out[True][k][u], ofr[k][u]))
        assert np.array_equal(out[True][k][~u], out[False][k][~u]), 'parked particles must be carried over untouched'
        assert rel(out['unsorted'][k][u], ofr[k][u]) < bar, ('unsorted', k, rel(out['unsorted'][k][u], ofr[k][u]))


def test_device_side_observation_equals_fluid_env_get_obs(emu):
    """MPMSimulator.get_obs_RL against envs/fluid_env.py:99-125 evaluated on the full get_state_RL() state: two
    bodies with different strides, an injector agent (8-vector state) — the same numbers, one small host copy."""
    from fluidlab_b200 import TaichiEnv, macros as M
    env = TaichiEnv(dim=3, quality=0.25, particle_density=3e4, max_substeps_local=40, gravity=(0.0, -10.0, 0.0), horizon=20, ckpt_dest='cpu', device='cpu')
    env.simulator.use_graphs = False
    env.setup_agent(dict(type='AgentInjector', effectors=[dict(type='Injector', params=dict(radius=0.02, flux=2, init_pos=(0.5, 0.6, 0.5), inject_v=(0.0, -2.0, 0.0), action_dim=3),
                                                               boundary=dict(type='cube', lower=(0.1, 0.1, 0.1), upper=(0.9, 0.9, 0.9)))]))
    env.setup_boundary(type='cube', lower=(0.2, 0.2, 0.2), upper=(0.8, 0.8, 0.8))
    env.add_body(type='nowhere', n_particles=60, material=M.MILK)
    env.add_body(type='cube', lower=(0.35, 0.3, 0.35), upper=(0.65, 0.42, 0.65), material=M.WATER)
    env.build()
    for _ in range(2):
        env.step(np.array([0.01, 0.0, -0.005]))
    n_obs = 25
    got = env.get_obs_RL(n_obs)
    state = env.get_state_RL()
    obs = []
    bodies = env.particles['bodies']
    assert bodies['n'] == 2
    for b in range(bodies['n']):     # fluid_env.py:104-115, verbatim logic
        ids = bodies['particle_ids'][b]
        step = max(1, bodies['n_particles'][b] // n_obs)
        obs += [state['x'][ids][::step].flatten(), state['v'][ids][::step].flatten(), state['used'][ids][::step].flatten()]
    obs += state['agent']
    want = np.concatenate(obs).astype(np.float32)
    assert got.dtype == np.float32 and got.shape == want.shape and got.size < 0.2 * state['x'].size * 3
    assert np.array_equal(got, want)
    assert state['used'][bodies['particle_ids'][0]].sum() == 40, 'the injector must have activated 2 particles in each of the 20 substeps'
    # render bridge: the same positions as device tensors, exportable through DLPack
    import torch
    r = env.simulator.get_state_render_device(env.simulator.cur_substep_local)
    x_dl = torch.utils.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(r.x))
    assert np.array_equal(x_dl.cpu().numpy(), state['x']) and np.array_equal(r.used.cpu().numpy(), state['used'])


def test_bench_script_runs_end_to_end_on_the_emulated_device():
    """bench.py's `ours` arm (warm-up, timed steps, per-kernel replay, e2e episodes with host buffers, the guarded observation-bridge
    episodes, JSON line) with a tiny workload on the shim: checks the SCRIPT, the driver's contract keys and the fused path's flag — the
    numbers are meaningless here."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(HERE, 'cuda_emu', 'run_bench_emu.py'), '--particles', '1200', '--steps', '2', '--warmup', '1', '--no-cpu', '--fuse-g2p2g', '1', '--min-seconds', '0', '--min-substeps', '40'], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'clocks', 'e2e',
              'gpu_launches', 'roofline', 'cpu_baseline'):
        assert k in line, k
    assert line['metric'] == 'mpm_substeps_per_s_fwd' and line['warmup'] >= 3 and line['value'] > 0 and line['e2e']['value'] > 0
    assert line['config']['g2p2g_fused'] is True and line['gpu_launches'] == 2 * 21 + 2
    assert set(line['roofline']) >= {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'}
    rf = line['roofline_fused']
    assert 'error' not in rf and rf['kernel'].startswith('k_fwd') and rf['launches_timed'] == 2 * 9 and line['roofline']['kernel'] == rf['kernel']
    assert line['timed_steps'] >= 2 and line['timed_steps'] % 2 == 0
    fb = line['fwd_bwd']
    assert fb['value'] > 0 and 'error' not in fb['whole_trajectory_ring'] and fb['whole_trajectory_ring']['max_substeps_local'] == 30
    assert isinstance(fb['adam_step_ms'], float), fb['adam_step_ms']
    ob = line['e2e_obs_bridge']
    assert 'error' not in ob and ob['d2h_bytes_per_step'] < line['e2e']['d2h_bytes_per_step']
    rb = fb['roofline_bwd']
    assert rb['frac'] > 0 and rb['backward_substep_ms'] > 0 and rb['algorithmic_bytes_per_substep'] == 432 * line['roofline']['n_used'] + 156 * line['roofline']['touched_nodes']


@pytest.mark.parametrize('cfg,extra', [('C3', ['--particles', '3000', '--steps', '6']), ('C4', ['--particles', '6000'])])
def test_bench_config_arms_run_on_the_emulated_device(cfg, extra):
    """`bench.py --config C3 | C4` (BASELINE configs[2] / configs[3] through TaichiEnv: forward, forward + backward with dLoss/dAction, e2e) at a few thousand
    particles on the shim: the SCRIPT and the scenes of tests/baseline_scenes.py; numbers meaningless.  C3 with 6 steps crosses the T = 50 ring boundary once."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(HERE, 'cuda_emu', 'run_bench_emu.py'), '--config', cfg, '--min-seconds', '0'] + extra, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line['metric'] == 'mpm_substeps_per_s_fwd' and line['value'] > 0 and line['fwd_bwd']['value'] > 0 and line['e2e']['value'] > 0
    assert cfg in line['config']['workload'] and line['e2e']['d2h_bytes_per_step'] == 28 * line['config']['n_particle_slots']
    if cfg == 'C3':
        assert line['fwd_bwd']['dloss_daction_absmax'] > 0


def test_circulation_stack_equals_a_run_of_the_real_reference_stack(emu):
    """the product's TaichiEnv circulation stack on the emulated device against tests/golden/reference_circulation.npz (a run of the
    reference's OWN MPMSimulator + AgentCirculation + AirCon + SmokeField); see tests/circulation_case.py"""
    from circulation_case import run_reference_stack_case
    run_reference_stack_case(device='cpu')


@pytest.mark.parametrize('exchange', ['peer', 'nccl', 'peer-signal', 'peer-signal-push', 'peer-signal-3ranks', 'peer-signal-push-3ranks'])
def test_slab_forward_with_g2p2g_fusion_on_the_emulated_device(exchange):
    """x-slabs + the fused forward kernel: 5 steps with migrations and two ring wraps against the single-domain (unfused) product.
    peer / peer-signal-push: the scatter half reduces the ghost planes of frame f+1 into the neighbour's accumulator (push form); nccl: the
    all-reduce follows it; peer-signal: the one-call step with the PULL form (local scatter, grid_op adds the neighbours' ghost planes, deferred
    ghost clears, n + 1 handshakes); -3ranks: a middle slab with two neighbours"""
    import torch.multiprocessing as mp
    mgr = mp.Manager(); ret = mgr.dict()
    sync = 'signal' if '-signal' in exchange else 'barrier'   # 'signal': the whole step is ONE library call (fmpm_substeps_slab)
    world = 3 if exchange.endswith('3ranks') else 2
    mp.spawn(_slab_worker, args=(world, _free_port(), ret, exchange.split('-')[0], True, sync, '-push' not in exchange), nprocs=world, join=True)
    out = dict(ret)
    ref_s = out[0]['ref_state']
    for r in range(world):
        fwd = out[r]['fwd']
        assert np.array_equal(fwd['gid'], np.arange(700)), 'particles lost or duplicated'
        assert rel(fwd['x'], ref_s['x']) < 1e-5 and rel(fwd['F'], ref_s['F']) < 1e-5 and rel(fwd['v'], ref_s['v']) < 1e-4
    assert all(out[r]['migrated'] > 0 for r in range(world))


@pytest.mark.parametrize('mode', ['forward-fused', 'backward'])
def test_slab_with_a_static_sdf_collider_matches_the_single_domain_run(mode):
    """x-slabs + a static SDF collider (meshes/static.py:26-104, applied in grid_op MPM:388-390) that straddles the slab boundary: every rank evaluates it on the nodes
    it converts (shared planes included, from identical ghost sums), so the sharded run — forward through the one-call pull-form steps, and forward + backward with
    dL/d(x0, v0, C0, F0) — equals the single-domain run with the same collider; and the collider must have acted (the cloud falls onto the box)."""
    import torch.multiprocessing as mp
    mgr = mp.Manager(); ret = mgr.dict()
    fused = mode == 'forward-fused'
    mp.spawn(_slab_worker, args=(2, _free_port(), ret, 'peer', fused, 'signal', True, True), nprocs=2, join=True)
    out = dict(ret)
    ref_s = out[0]['ref_state']
    assert out[0]['static_effect'] > 1e-3, 'the collider must change the run'
    for r in (0, 1):
        fwd = out[r]['fwd']
        assert np.array_equal(fwd['gid'], np.arange(700)), 'particles lost or duplicated'
        assert rel(fwd['x'], ref_s['x']) < 1e-5 and rel(fwd['F'], ref_s['F']) < 1e-5 and rel(fwd['v'], ref_s['v']) < 1e-4
        if not fused:
            errs = {k: rel(out[r]['grad'][k], out[0]['ref_grad'][k].astype(np.float64)) for k in 'xvCF'}
            assert errs['x'] < 1e-4 and errs['v'] < 1e-4 and errs['C'] < 2e-3 and errs['F'] < 2e-3, errs


@pytest.mark.parametrize('scene', ['latteart', 'jetbot', 'jetbot_randv', 'pouring', 'icecream', 'latteart_fused'])
def test_agent_scenes_equal_runs_of_the_real_reference_agents(emu, scene):
    """product (real kernels on the emulated device) vs runs of the reference's own AgentInjector / AgentJetBot / AgentPouring / AgentIceCreamDynamic scenes; tests/reference_scene_cases.py"""
    import reference_scene_cases as cases
    getattr(cases, f'run_{scene}_case')(device='cpu')


@pytest.mark.parametrize('scene', ['jetbot', 'jetbot_randv', 'pouring', 'icecream'])
def test_agent_scenes_through_the_fused_path_equal_the_reference_runs(emu, scene):
    """the same reference runs with MPMSimulator.fuse_g2p2g: 6-DOF injector + collector (JetBot), Rigid SDF collider at grid and particle level +
    collector (Pouring), BallInjector with inject_till + gated Rigid collider + Static collider (IceCream)"""
    import reference_scene_cases as cases
    cases.FUSE[0] = True
    try:
        getattr(cases, f'run_{scene}_case')(device='cpu')
    finally:
        cases.FUSE[0] = False


def test_device_adjoint_equals_finite_differences_through_the_reference_forward(emu):
    import reference_scene_cases as cases
    cases.run_cloud_adjoint_case(device='cpu')


def test_multi_rank_bench_arm_runs_on_the_emulated_device():
    """bench.py --gpus 2 (x-slabs, peer ghost reduction over shared memory standing in for NVLink, neighbour-handshake sync, g2p2g fusion, per-kernel
    timing section with its cross-rank clear, e2e episodes with migration) as two gloo ranks on the shim: the SCRIPT must finish and print the line"""
    import json
    import subprocess
    port = _free_port()
    procs = []
    for r in (0, 1):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), SLAB_SYNC='signal')
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, 'cuda_emu', 'run_bench_emu.py'), '--gpus', '2', '--particles', '3000', '--steps', '2', '--warmup', '1',
                                       '--no-cpu', '--bwd', '0', '--fuse-g2p2g', '1', '--min-seconds', '0', '--min-substeps', '40'], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
    outs = [p.communicate(timeout=900) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs[0][1][-2000:] + outs[1][1][-2000:]
    line = json.loads(outs[0][0].strip().splitlines()[-1])
    assert line['n_gpus'] == 2 and line['value'] > 0 and line['e2e']['value'] > 0 and line['config']['g2p2g_fused'] is True
    assert 'neighbour handshake' in line['config']['parallelism'] and outs[1][0].strip() == ''


def test_neighbour_handshake_gives_up_instead_of_hanging_when_a_peer_never_arrives():
    """k_slab_sync with a neighbour that never posts its epoch: the wait must end by itself and raise the error word (on hardware a spinning
    kernel that never ends would take the GPU down with it).  Built with a 0.2 s timeout instead of the product's 10 s."""
    import ctypes as C
    import subprocess
    import time
    from fluidlab_b200 import _lib
    out = os.path.join(harness.EMU_DIR, '_build', 'libfluidmpm_emu_timeout.so')
    deps = [os.path.join(harness.CSRC, f) for f in os.listdir(harness.CSRC) if f.endswith(('.cu', '.cuh'))] + [os.path.join(harness.ROOT, 'include', 'fluidmpm.h'),
                                                                                                              os.path.join(harness.EMU_DIR, 'cuda_runtime.h')]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(['/usr/bin/g++', '-std=c++20', '-O1', '-fPIC', '-shared', '-pthread', '-x', 'c++', '-I', harness.EMU_DIR, '-DFMPM_BUILD',
                               '-DFMPM_SYNC_TIMEOUT_NS=200000000ULL'] + [os.path.join(harness.CSRC, s) for s in harness.SRCS] + ['-o', out])
    L = C.CDLL(out)
    for name, (res, args) in _lib._PROTOS.items():
        fn = getattr(L, name); fn.restype = res; fn.argtypes = args
    cfg = _lib.FmpmConfig()
    cfg.n_grid, cfg.n_particles, cfg.max_substeps_local, cfg.n_substeps, cfg.n_materials = 16, 0, 10, 10, 1
    h = C.c_void_p()
    assert L.fmpm_create(C.byref(cfg), C.byref(h)) == 0
    mine, silent_peer = np.zeros(8, np.int32), np.zeros(8, np.int32)
    slab = _lib.FmpmSlab()
    slab.enabled, slab.signal, slab.peer_signal_right = 1, mine.ctypes.data, silent_peer.ctypes.data
    slab.right_lo, slab.right_hi = 4, 12
    assert L.fmpm_set_slab(h, C.byref(slab)) == 0
    t0 = time.perf_counter()
    assert L.fmpm_slab_sync(h, None) == 0
    dt = time.perf_counter() - t0
    assert 0.15 < dt < 5.0, dt
    assert mine[2] == 1 and silent_peer[0] == 1, 'the epoch must have been posted to the neighbour'
    assert mine[3] == 1, 'the error word must be raised'
    # and with a peer that has arrived the call returns at once without raising it
    mine[3] = 0; mine[1] = 2
    t0 = time.perf_counter()
    assert L.fmpm_slab_sync(h, None) == 0
    assert time.perf_counter() - t0 < 0.1 and mine[3] == 0 and mine[2] == 2
    L.fmpm_destroy(h)


def test_latteart_forward_backward_with_fused_injector_steps(emu):
    """TaichiEnv with an AgentInjector and the index-matched MILK loss (the call sequence of optimizer/solver.py:23-59), ring of 2 steps for a
    3-step horizon (chunk checkpoint + re-simulation): with MPMSimulator.fuse_g2p2g the stored-grid forward runs g2p2g kernels + the separate
    scatter of the freshly injected particles; loss and dLoss/dAction must equal the unfused run's and the fp64 oracle's"""
    from fluidlab_b200 import TaichiEnv, LatteArtLoss, macros as M
    n_grid, n_coffee, n_milk, flux, T, n_steps = 16, 400, 120, 2, 20, 3
    rng = np.random.RandomState(21)
    x = np.concatenate([np.tile(M.NOWHERE, (n_milk, 1)), rng.uniform((0.38, 0.36, 0.38), (0.62, 0.45, 0.62), size=(n_coffee, 3))])
    mat = np.concatenate([np.full(n_milk, M.MILK), np.full(n_coffee, M.COFFEE)])
    used = np.concatenate([np.zeros(n_milk), np.ones(n_coffee)]).astype(np.int32)
    P = make_particles(x, mat, n_grid, used=used)
    bnd = dict(type='cylinder', xz_radius=0.2, xz_center=(0.5, 0.5), y_range=(0.34, 0.9))
    ebnd = dict(type='cylinder', xz_radius=0.12, xz_center=(0.5, 0.5), y_range=(0.55, 0.55))
    cfg = dict(type='AgentInjector', effectors=[dict(type='Injector', params=dict(radius=0.0075, flux=flux, init_pos=(0.5, 0.5, 0.5), action_dim=3, inject_v=(0.0, -3.0, 0.0),
                                                                                 action_scale_p=(1.0, 1.0, 1.0), action_scale_v=(1.0, 1.0, 1.0), locally_random=True), boundary=ebnd)])
    tgt = [rng.uniform(0.4, 0.6, size=x.shape).astype(np.float32) for _ in range(n_steps)]
    actions = rng.uniform(-0.004, 0.004, size=(n_steps, 3)).astype(np.float32)
    action_p = np.array([0.47, 0.55, 0.52], dtype=np.float32)
    res = {}
    for fuse in (False, True):
        env = TaichiEnv(quality=n_grid / 64, max_substeps_local=T, gravity=(0.0, -20.0, 0.0), horizon=n_steps, ckpt_dest='cpu', device='cpu')
        env.simulator.use_graphs, env.simulator.fuse_g2p2g = False, fuse
        np.random.seed(5)
        env.setup_agent(cfg)
        env.particle_bodies.get = lambda: P
        env.setup_boundary(**bnd)
        env.setup_loss(loss_cls=LatteArtLoss, type='diff', target=tgt, weights={'chamfer': 1.0})
        env.build()
        rv = env.agent.effectors[0].random_vector_np
        env.set_state(env.get_state()['state'], grad_enabled=True)
        assert env.simulator._can_fuse_injector() == fuse
        env.apply_agent_action_p(action_p)
        for i in range(n_steps):
            env.step(actions[i])
        info = env.get_final_loss()
        env.reset_grad(); env.get_final_loss_grad()
        for i in range(n_steps - 1, -1, -1):
            env.step_grad(actions[i])
        env.apply_agent_action_p_grad(action_p)
        res[fuse] = (info['loss'], env.agent.get_grad(n_steps))
    o = orc.OracleSim(n_grid, P, gravity=(0, -20, 0), boundary=bnd, precision=64, max_substeps_local=T)
    o.add_effector(type=1, action_dim=3, boundary=ebnd, radius=0.0075, flux=flux, inject_v=(0, -3, 0), inject_p=(0, 0, 0), locally_random=True, random_vector=rv,
                   act_range=np.where(used == 0)[0], max_action_steps=n_steps + 1)
    N = len(x)
    o.enable_grad()
    o.set_frame(0, P['x'], np.zeros((N, 3)), np.zeros((N, 3, 3)), np.tile(np.eye(3), (N, 1, 1)), P['used'])
    o.set_effector_state(0, 0, np.array([0.5, 0.5, 0.5, 1, 0, 0, 0, 0.0]))
    o.apply_action_p(action_p)
    total = 0.0
    for i in range(n_steps):
        o.step(actions[i]); total += o.loss_value(o.cur_substep_local, M.MILK, 1.0, tgt[i])
    o.reset_grad()
    for i in range(n_steps - 1, -1, -1):
        o.loss_seed(o.cur_substep_local, M.MILK, 1.0, tgt[i]); o.step_grad(actions[i])
    o.apply_action_p_grad()
    og = o.get_action_grad(n_steps)
    for fuse in (False, True):
        loss, grad = res[fuse]
        assert abs(loss - total) <= 1e-5 * abs(total), (fuse, loss, total)
        assert np.abs(og).max() > 1e-3 and rel(grad, og) < 1e-4, (fuse, rel(grad, og))


def test_c1_rollout_timing_script_runs_on_the_emulated_device(emu):
    """profiles/c1_rollout_times.py (LatteArt rollouts with the fused path off / on, queued for the next GPU round) at a reduced size: script sanity"""
    sys.path.insert(0, os.path.join(harness.ROOT, 'profiles'))
    import c1_rollout_times
    out = c1_rollout_times.run(device='cpu', n_steps=2, reps=1, sync=lambda: None, n_milk=300, quality=0.25, T=40)
    assert out['plain']['n_particles'] == out['fused']['n_particles'] > 300 and out['fused']['substeps_per_s'] > 0


@pytest.mark.parametrize('cxxflags', ['', '-O2 -mfma -ffp-contract=fast'], ids=['default', 'fma-contracted'])
def test_gpu_marked_parity_tests_pass_on_the_shim(cxxflags):
    """the `-m gpu` parity tests THEMSELVES (tests/test_gpu_parity.py, test_golden.py, test_zz_smoke_gpu.py — the ones the B200 box runs), small
    scenes only, in a subprocess with FLUIDLAB_CUDA_EMU=1 (tests/conftest.py routes the library to the shim): whatever they assert about the
    kernels holds for the kernel code as written; what remains for the GPU is the hardware.  (The full-size ones pass on the shim too — up to
    C5's 8M particles — but take minutes and tens of GB, so they are left to manual runs.)
    'fma-contracted': the same on a shim build that contracts a*b+c into FMAs the way nvcc does (CUEMU_CXXFLAGS, tests/cuda_emu/harness.py) — a second
    rounding variant of every kernel, so a parity bar that only holds for one instruction selection fails here rather than on the GPU box."""
    import subprocess
    if cxxflags and ' fma ' not in open('/proc/cpuinfo').read().replace('\n', ' '):
        pytest.skip('host CPU has no FMA')
    sel = ('forward_phases or substep_grad_matches or state_io or dloss_daction_latteart or out_of_grid or library_error or ragged or no_used or reference_kernels '
           'or reference_agents or fused_path or finite_differences or golden or rigid_material_bodies or smoke_forward or smoke_backward or real_reference_stack')
    env = dict(os.environ, FLUIDLAB_CUDA_EMU='1', CUEMU_CXXFLAGS=cxxflags)
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(HERE, 'test_gpu_parity.py'), os.path.join(HERE, 'test_golden.py'), os.path.join(HERE, 'test_zz_smoke_gpu.py'),
                        '-m', 'gpu', '-q', '-x', '-k', sel, '-p', 'no:cacheprovider'], capture_output=True, text=True, timeout=1500, env=env, cwd=harness.ROOT)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ''
    assert r.returncode == 0 and ' passed' in tail and 'failed' not in tail, r.stdout[-3000:] + r.stderr[-1000:]
    assert int(tail.split(' passed')[0].split()[-1]) >= 56, tail


def test_graft_entry_smoke_runs_on_the_emulated_device(emu, monkeypatch):
    """__graft_entry__.smoke() — the call the driver makes on cuda:0 before the bench — as written, on the shim (device checks patched only)"""
    from fluidlab_b200 import simulator
    init = simulator.MPMSimulator.__init__

    def emu_init(self, *a, device=None, **k):
        init(self, *a, device='cpu' if device is None else device, **k)
        self.use_graphs = False
    monkeypatch.setattr(simulator.MPMSimulator, '__init__', emu_init)
    monkeypatch.setattr(torch.cuda, 'is_available', lambda: True)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda d: None)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    import __graft_entry__ as g
    g.smoke()


@pytest.mark.parametrize('liquid,boundary,sort_every', [(False, 'cube', 1), (True, 'cube', 1), (True, 'cylinder', 0), (False, 'cylinder', 2)],
                         ids=['multimat', 'liquid', 'liquid-cyl-unsorted', 'multimat-cyl-sort2'])
def test_every_forward_path_of_fmpm_substeps_fused_equals_the_plain_substeps_and_the_oracle(emu, liquid, boundary, sort_every):
    """k_fwd (g2p + [grid_op] + p2g in one kernel) with each feature switched on in turn — all-liquid specialisation (F carried as one float
    between step boundaries), grid_op inlined over the triple-buffered accumulators — against the plain substeps and the fp64 oracle
    (tests/fwd_path_case.py; the same body runs on a B200 in tests/test_gpu_parity.py)."""
    import fwd_path_case
    fwd_path_case.run('cpu', liquid, [0, 1, 3, 5, 7] if liquid else [0, 1, 5], boundary=boundary, sort_every=sort_every)


def test_c4_scene_at_reduced_size_on_the_emulated_device(emu):
    """the C4 scene of tests/test_gpu_parity.py (ELASTIC + ICECREAM blocks, soft cone collider configured the way agent_icecreamdynamic.yaml does —
    material by NAME, scale, euler, softness —, IceCreamDynamicLoss, forward + dLoss/dAction vs the oracle) with 2 x 6,000 particles on a 48^3
    grid: the same body the B200 runs with 2 x 1,000,000 on 192^3"""
    import test_gpu_parity as g
    from fluidlab_b200 import simulator
    orig = simulator.MPMSimulator.__init__

    def init_cpu(self, *a, device=None, **k):
        orig(self, *a, device='cpu', **k); self.use_graphs = False
    simulator.MPMSimulator.__init__ = init_cpu
    try:
        g.c4_case(48, 6000)
    finally:
        simulator.MPMSimulator.__init__ = orig


def test_staged_state_upload_equals_blocking_set_state_on_the_emulated_device(emu):
    """stage_state_async + set_state vs set_state (tests/test_gpu_parity.py: staged_upload_case) on the shim: the host logic of the staging sets"""
    import test_gpu_parity as g
    g.staged_upload_case(device='cpu')
