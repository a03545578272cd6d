"""Shared body of the forward-path tests (k_fwd and its specialisations, csrc/fmpm_forward.cu): run on a B200 by tests/test_gpu_parity.py
(`-m gpu`) and on the CPU execution-model shim by tests/test_cuda_emu_mpm.py.

fmpm_substeps_fused picks its kernels from what the scene allows (fmpm_fwd_path); FMPM_FWD_MASK / fmpm_set_fwd_mask switch features off:
    0  round-1 path (grid_op + k_g2p2g)        1  k_fwd                      3  k_fwd, all-liquid specialisation (F carried as one float)
    5  k_fwd + lazy in-kernel grid_op (one launch per substep, triple-buffered accumulators)       7  both      + 8: footprint tile by TMA (9, 11)
Every mask must reproduce the plain p2g / grid_op / g2p substeps and the fp64 oracle."""
import numpy as np

from conftest import make_particles


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-12))


def run(device, liquid, masks, boundary='cube', n=16, N=330, sort_every=1, steps=2, unused=0.12, seed=7):
    from fluidlab_b200 import MPMSimulator, macros as M, _lib
    from oracle import oracle as orc
    rng = np.random.RandomState(seed)
    x = rng.uniform(0.34, 0.66, size=(N, 3)).astype(np.float32)
    mats = [M.WATER, M.MILK, M.COFFEE] if liquid else [M.WATER, M.ELASTIC, M.ICECREAM]
    mat = np.array([mats[i % 3] for i in range(N)], dtype=np.int32)
    used = (rng.rand(N) > unused).astype(np.int32)
    v0 = (rng.randn(N, 3) * 0.8).astype(np.float32)
    if liquid:   # a liquid's F is s I after its first substep; start from that form with s != 1, and from a general F for a few particles
        sdiag = (1.0 + rng.randn(N) * 0.02).astype(np.float32)
        F0 = (np.eye(3)[None] * sdiag[:, None, None]).astype(np.float32)
        F0[::7] += (rng.randn(len(F0[::7]), 3, 3) * 0.02).astype(np.float32)
    else:
        F0 = (np.eye(3)[None] + rng.randn(N, 3, 3) * 0.05).astype(np.float32)
    C0 = (rng.randn(N, 3, 3) * 2.0).astype(np.float32)
    P = make_particles(x, mat, n, used=used)
    bnd = (dict(type='cube', lower=(0.3, 0.3, 0.3), upper=(0.7, 0.7, 0.7)) if boundary == 'cube'
           else dict(type='cylinder', xz_radius=0.21, xz_center=(0.5, 0.5), y_range=(0.3, 0.7)))
    T = 10 * steps + 10
    out = {}
    for mask in [None] + list(masks):
        s = MPMSimulator(dim=3, quality=n / 64, gravity=(0.3, -10, 0), horizon=50, max_substeps_local=T, max_substeps_global=1000, ckpt_dest='cpu', device=device,
                         sort_every=sort_every)
        s.use_graphs, s.fuse_g2p2g = False, mask is not None
        s.setup_boundary(**bnd)
        s.build(None, None, [], P)
        if mask is not None:
            s._ck(s._lib.fmpm_set_fwd_mask(s._h, int(mask)), 'fmpm_set_fwd_mask')
            path = int(s._lib.fmpm_fwd_path(s._h))
            assert (path & 7) == (mask & (7 if liquid else 5)) if (mask & 1) else path == 0, (mask, path)
            assert (path & 8) in (0, mask & 8)   # bit 3: the footprint tile arrives by TMA (a GPU with the driver's tensor-map encoder; never on the CPU shim)
        st = s.get_state(); st['v'][:] = v0; st['F'][:] = F0; st['C'][:] = C0; s.set_state(0, st)
        for _ in range(steps):
            s.step(None)
        out[mask] = s.get_state()
        if mask is not None and (mask & 4):   # the triple-buffered accumulators and their block flags are clear again
            assert float(s._grid_pm3.abs().max()) == 0.0 and int(s._blk_flags3.abs().max()) == 0
        assert float(s._grid_pm.abs().max()) == 0.0
    o = orc.OracleSim(n, P, gravity=(0.3, -10, 0), boundary=bnd, precision=64, max_substeps_local=T)
    o.set_frame(0, x, v0, C0, F0, used)
    for _ in range(steps):
        o.step(None)
    ofr = o.get_frame(10 * steps)
    u = used != 0
    err = {}
    for mask in masks:
        assert np.array_equal(out[mask]['used'], used)
        for k, bar in (('x', 1e-6), ('F', 1e-5), ('v', 1e-4), ('C', 1e-3)):
            e1, e2 = rel(out[mask][k][u], out[None][k][u].astype(np.float64)), rel(out[mask][k][u], ofr[k][u])
            err[(mask, k)] = (e1, e2)
            assert e1 < bar and e2 < bar, (mask, k, e1, e2)
            assert np.array_equal(out[mask][k][~u], out[None][k][~u]), 'parked particles must be carried over untouched'
    return err
