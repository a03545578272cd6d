"""fluidlab_b200/csrc/fsmk_smoke.cu — the product's smoke kernels AND their host launch logic — executed on the CPU and checked against
the oracle.

There is no GPU in the build container, so the .cu translation unit is compiled UNCHANGED by g++ against tests/cuda_emu/cuda_runtime.h, a
small model of the CUDA execution model (one host thread per CUDA thread, block by block: __syncthreads, shared memory, warp shuffles and
atomics behave as on the device).  What this covers: indexing, the time-blocked Jacobi tiles (valid-region argument, chained launches,
partial last launch), the gather-form adjoints, the block reduction of the air conditioner's adjoints, the C-ABI argument checks.  What it
cannot cover: anything about the real hardware (performance, memory-model subtleties) — that is tests/test_gpu_parity.py's job on a B200."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

from fluidlab_b200 import _lib
from oracle.smoke import SmokeOracle
from test_smoke_oracle import G as GOLDEN, rel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, 'tests', 'cuda_emu')


@pytest.fixture(scope='module')
def emu():
    """the whole product library built for the host (tests/cuda_emu/harness.py), smoke entry points attached"""
    import sys
    sys.path.insert(0, EMU_DIR)
    import harness
    return _lib.attach_smoke_protos(C.CDLL(harness.build_library()))


class EmuSmoke:
    """drives the C ABI of include/fluidsmoke.h with host (NumPy) buffers in the layouts the header documents"""

    def __init__(self, L, res, S, q_dim, iters, dt, lower_y, higher_y, inject_v, T_sub=40, grads=True):
        self.L, self.n, self.S, self.qd = L, res, S, q_dim
        n, G = res, res ** 3
        cfg = _lib.FsmkConfig()
        cfg.res, cfg.max_steps_local, cfg.q_dim, cfg.solver_iters, cfg.dt = res, S, q_dim, iters, dt
        cfg.lower_y, cfg.higher_y, cfg.low_T, cfg.device = lower_y, higher_y, 0.0, 0
        cfg.inject_v = (C.c_float * 3)(*inject_v)
        self.h = C.c_void_p()
        assert L.fsmk_create(C.byref(cfg), C.byref(self.h)) == 0
        f32 = np.float32
        self.a = dict(v=np.zeros((S + 1, G, 4), f32), v_tmp=np.zeros((S + 1, G, 4), f32), div=np.zeros((S + 1, G), f32), p=np.zeros((S + 1, G), f32),
                      q=np.zeros((S + 1, q_dim, G), f32), is_free=np.zeros((S + 1, G), np.uint8), tmp_a=np.zeros(G, f32), tmp_b=np.zeros(G, f32), acc=np.zeros(G, f32))
        if grads:
            self.a.update(gv=np.zeros((S + 1, G, 4), f32), gv_tmp=np.zeros((S + 1, G, 4), f32), gdiv=np.zeros((S + 1, G), f32), gp=np.zeros((S + 1, G), f32),
                          gq=np.zeros((S + 1, q_dim, G), f32))
        b = _lib.FsmkBuffers()
        for k, arr in self.a.items():
            setattr(b, k, arr.ctypes.data)
        self.ck(L.fsmk_bind(self.h, C.byref(b)))
        self.air = dict(pos=np.zeros((T_sub + 1, 3), f32), quat=np.zeros((T_sub + 1, 4), f32), s=np.zeros(T_sub + 1, f32), r=np.zeros(T_sub + 1, f32),
                        gpos=np.zeros((T_sub + 1, 3), f32), gquat=np.zeros((T_sub + 1, 4), f32), gs=np.zeros(T_sub + 1, f32), gr=np.zeros(T_sub + 1, f32))
        self.air['quat'][:, 0] = 1
        a = _lib.FsmkAircon()
        for k, arr in self.air.items():
            setattr(a, k, arr.ctypes.data)
        self.ck(L.fsmk_set_aircon(self.h, C.byref(a)))
        self._keep = []

    def ck(self, rc):
        assert rc == 0, self.L.fsmk_last_error(self.h).decode()

    def set_statics(self, voxs, Ts):
        arr = (_lib.FmpmSdfMesh * len(voxs))()
        for i, (vox, T) in enumerate(zip(voxs, Ts)):
            v = np.ascontiguousarray(vox, dtype=np.float32); self._keep.append(v)
            arr[i].voxels = v.ctypes.data; arr[i].res = int(round(v.size ** (1 / 3)))
            arr[i].T_mesh_to_voxels = (C.c_float * 16)(*[float(x) for x in np.asarray(T).reshape(-1)])
        self.ck(self.L.fsmk_set_statics(self.h, len(voxs), arr))

    def set_aircon(self, f, st):
        self.air['pos'][f] = st[0:3]; self.air['quat'][f] = st[3:7]; self.air['s'][f] = st[7]; self.air['r'][f] = st[8]

    def set_state(self, s, st):
        n = self.n
        self.a['v'][s, :, :3] = st['v'].reshape(-1, 3); self.a['v_tmp'][s, :, :3] = st['v_tmp'].reshape(-1, 3)
        self.a['div'][s] = st['div'].reshape(-1); self.a['p'][s] = st['p'].reshape(-1)
        self.a['q'][s] = np.moveaxis(st['q'].reshape(-1, self.qd), 1, 0)

    def get_state(self, s, grad=False):
        n, pre = self.n, ('g' if grad else '')
        return dict(v=self.a[pre + 'v'][s, :, :3].reshape(n, n, n, 3).copy(), v_tmp=self.a[pre + 'v_tmp'][s, :, :3].reshape(n, n, n, 3).copy(),
                    div=self.a[pre + 'div'][s].reshape(n, n, n).copy(), p=self.a[pre + 'p'][s].reshape(n, n, n).copy(),
                    q=np.moveaxis(self.a[pre + 'q'][s], 0, 1).reshape(n, n, n, self.qd).copy())

    def set_grad(self, s, st):
        self.a['gv'][s, :, :3] = st['v'].reshape(-1, 3); self.a['gp'][s] = st['p'].reshape(-1)
        self.a['gq'][s] = np.moveaxis(st['q'].reshape(-1, self.qd), 1, 0)

    def step(self, s, f):
        self.ck(self.L.fsmk_step(self.h, s, f, None))

    def step_grad(self, s, f):
        self.ck(self.L.fsmk_step_grad(self.h, s, f, None))


def _pair(L, d, iters, grads=True):
    kw = dict(res=int(d['res']), q_dim=int(d['q_dim']), dt=float(d['dt']), lower_y=int(d['lower_y']), higher_y=int(d['higher_y']), inject_v=tuple(float(x) for x in d['inject_v']))
    e = EmuSmoke(L, S=4, iters=iters, grads=grads, **kw)
    o = SmokeOracle(solver_iters=iters, max_steps_local=4, max_substeps_local=40, precision=32, **kw)
    e.set_statics(list(d['vox']), list(d['T_static']))
    for vox, T in zip(d['vox'], d['T_static']):
        o.add_static(vox, T)
    for f, a in zip(d['air_f'], d['air']):
        e.set_aircon(int(f), a); o.set_aircon(int(f), a)
    st0 = {k: d['st0_' + k] for k in ('v', 'v_tmp', 'div', 'p', 'q')}
    e.set_state(0, st0); o.set_state(0, st0)
    return e, o


@pytest.mark.parametrize('iters', [0, 6, 8, 21])
def test_forward_kernels_match_the_oracle_and_the_reference_run(emu, iters):
    """iters = 6: the reference-run fixture's setting (one partial launch of the time-blocked Jacobi); 8: exactly one full launch;
    21: three chained launches through both scratch buffers; 0: the copy-only chain."""
    d = np.load(os.path.join(GOLDEN, 'reference_smoke.npz'))
    e, o = _pair(emu, d, iters, grads=False)
    for s in range(3):
        e.step(s, 10 * s); o.step(s, 10 * s)
    assert np.array_equal(e.a['is_free'][0].reshape((e.n,) * 3), o.is_free(0))
    for s in (1, 2, 3):
        a, b = e.get_state(s), o.get_state(s)
        for k in ('v', 'p', 'q'):
            assert rel(a[k], b[k]) < 2e-6, (iters, s, k, rel(a[k], b[k]))
    for s in (0, 1, 2):
        a, b = e.get_state(s), o.get_state(s)
        assert rel(a['v_tmp'], b['v_tmp']) < 2e-6 and rel(a['div'], b['div']) < 1e-5
    if iters == 6:   # and directly against the run of the reference's own kernels
        a = e.get_state(3)
        for k in ('v', 'p', 'q'):
            assert rel(a[k], d['ref3_' + k]) < 2e-5, (k, rel(a[k], d['ref3_' + k]))


def test_time_blocked_jacobi_equals_sweep_by_sweep_jacobi_bit_for_bit(emu):
    """the tile kernel (up to 8 sweeps per launch in shared memory) against the oracle's plain sweeps, 19 iterations: identical float32
    results cell by cell — the reference's arithmetic in the reference's order"""
    d = np.load(os.path.join(GOLDEN, 'reference_smoke.npz'))
    e, o = _pair(emu, d, 19, grads=False)
    e.ck(emu.fsmk_free_space(e.h, 0, None)); e.ck(emu.fsmk_advect(e.h, 0, 0, None)); e.ck(emu.fsmk_divergence(e.h, 0, None))
    o.step(0, 0)
    # feed the oracle's own v_tmp / div so that the comparison isolates the solver
    ost = o.get_state(0)
    e.a['div'][0] = ost['div'].reshape(-1).astype(np.float32)
    e.ck(emu.fsmk_pressure(e.h, 0, None))
    free = o.is_free(0).astype(bool)
    assert np.array_equal(e.get_state(1)['p'][free], o.get_state(1)['p'][free].astype(np.float32))


@pytest.mark.parametrize('iters', [6, 13])
def test_adjoint_kernels_match_the_oracle(emu, iters):
    d = np.load(os.path.join(GOLDEN, 'reference_smoke.npz'))
    fd = np.load(os.path.join(GOLDEN, 'reference_smoke_fd.npz'))
    e, o = _pair(emu, d, iters)
    for s in range(2):
        e.step(s, 10 * s); o.step(s, 10 * s)
    o.reset_grad()
    z = o._alloc(); z['v'], z['q'], z['p'] = fd['w_v'], fd['w_q'], fd['w_p']
    o.set_grad(2, z); e.set_grad(2, z)
    for s in (1, 0):
        e.step_grad(s, 10 * s); o.step_grad(s, 10 * s)
    a, b = e.get_state(0, grad=True), o.get_grad(0)
    for k in ('v', 'q', 'p'):
        assert np.abs(b[k]).max() > 0 and rel(a[k], b[k]) < 1e-4, (k, rel(a[k], b[k]))
    for f in (0, 10):
        ga = np.concatenate([e.air['gpos'][f], e.air['gquat'][f], [e.air['gs'][f]], [e.air['gr'][f]]])
        gb = o.aircon_grad(f)
        assert np.abs(ga - gb).max() < 1e-4 * np.abs(gb).max(), (f, ga, gb)
    if iters == 6:   # and against finite differences through the reference's own forward kernels (fp64), at fp32 accuracy
        for k in ('v', 'q', 'p'):
            an = float((a[k].astype(np.float64) * fd['dir_' + k]).sum())
            assert abs(an - float(fd['fd_' + k])) < 2e-3 * abs(float(fd['fd_' + k])), (k, an, float(fd['fd_' + k]))


def test_abi_rejects_bad_calls(emu):
    d = np.load(os.path.join(GOLDEN, 'reference_smoke.npz'))
    e, _ = _pair(emu, d, 2, grads=False)
    assert emu.fsmk_step(e.h, 99, 0, None) != 0 and b'out of range' in emu.fsmk_last_error(e.h)
    assert emu.fsmk_step_grad(e.h, 0, 0, None) != 0 and b'gradient buffers were not bound' in emu.fsmk_last_error(e.h)


def test_product_host_class_on_the_emulated_library(emu, monkeypatch):
    """fluidlab_b200.smoke.SmokeField — the product's host class (tensor layouts, state I/O, ring helpers, checkpoint, statics and
    air-conditioner registration) — driven end to end with the emulated library in place of libfluidmpm.so and CPU tensors in place of
    HBM: 3 steps with a ring wrap (copy_frame) then the backward pass, against the oracle.  (On a GPU box the same class runs on the real
    library: tests/test_gpu_parity.py.)"""
    import types
    import torch
    from fluidlab_b200 import smoke as smoke_mod, meshes
    from fluidlab_b200 import macros as M
    monkeypatch.setattr(smoke_mod._lib, 'load', lambda: emu)
    d = np.load(os.path.join(GOLDEN, 'reference_smoke.npz'))
    fd = np.load(os.path.join(GOLDEN, 'reference_smoke_fd.npz'))
    n = int(d['res'])
    Tsub, dev = 40, torch.device('cpu')
    z = lambda *s: torch.zeros(s, dtype=torch.float32)
    air = types.SimpleNamespace(pos=z(Tsub + 1, 3), quat=z(Tsub + 1, 4), s=z(Tsub + 1), r=z(Tsub + 1), gpos=z(Tsub + 1, 3), gquat=z(Tsub + 1, 4), gs=z(Tsub + 1), gr=z(Tsub + 1),
                                inject_v=np.asarray(d['inject_v']))
    air.quat[:, 0] = 1
    for f, a in zip(d['air_f'], d['air']):
        a = torch.from_numpy(a.astype(np.float32)); air.pos[int(f)] = a[:3]; air.quat[int(f)] = a[3:7]; air.s[int(f)] = a[7]; air.r[int(f)] = a[8]
    statics = meshes.Statics()
    for vox, T in zip(d['vox'], d['T_static']):
        statics.add_static(file='x.obj', material=M.PILLAR, has_dynamics=True, sdf=dict(voxels=vox, T_mesh_to_voxels=T))
    agent = types.SimpleNamespace(aircon=air)
    sim = types.SimpleNamespace(max_steps_local=4, agent=agent, device=dev, statics=statics, _stream=lambda: None)
    sf = smoke_mod.SmokeField(dim=3, ckpt_dest='cpu', res=n, dt=float(d['dt']), solver_iters=int(d['iters']), q_dim=int(d['q_dim']))
    sf.lower_y, sf.higher_y = int(d['lower_y']), int(d['higher_y'])
    sf.build(sim, agent)
    assert np.array_equal(sf.get_state(0)['q'], d['q_init']), 'init_fields'
    st0 = {k: d['st0_' + k] for k in ('v', 'v_tmp', 'div', 'p', 'q')}
    sf.set_state(0, st0)
    back = sf.get_state(0)
    assert all(np.array_equal(back[k], st0[k]) for k in st0), 'set_state / get_state round trip'
    ck = sf.get_ckpt('000000')
    for s in range(3):
        sf.step(s, 10 * s)
    assert np.array_equal(sf.is_free(0), d['free0'])
    r3 = sf.get_state(3)
    for k in ('v', 'p', 'q'):
        assert rel(r3[k], d['ref3_' + k]) < 2e-5, (k, rel(r3[k], d['ref3_' + k]))
    # ring helpers + checkpoint
    sf.copy_frame(3, 0)
    assert all(np.array_equal(sf.get_state(0)[k], r3[k]) for k in ('v', 'p', 'q'))
    sf.set_ckpt(ckpt_name='000000')
    assert all(np.array_equal(sf.get_state(0)[k], st0[k]) for k in st0)
    # backward of the first two steps against the oracle
    o = SmokeOracle(res=n, dt=float(d['dt']), solver_iters=int(d['iters']), q_dim=int(d['q_dim']), max_steps_local=4, max_substeps_local=40, lower_y=int(d['lower_y']),
                    higher_y=int(d['higher_y']), inject_v=tuple(d['inject_v']), precision=32)
    for vox, T in zip(d['vox'], d['T_static']):
        o.add_static(vox, T)
    for f, a in zip(d['air_f'], d['air']):
        o.set_aircon(int(f), a)
    o.set_state(0, st0); o.step(0, 0); o.step(1, 10)
    zz = o._alloc(); zz['v'], zz['q'], zz['p'] = fd['w_v'], fd['w_q'], fd['w_p']
    o.reset_grad(); o.set_grad(2, zz)
    sf.reset_grad(); sf.set_grad(2, zz)
    for s in (1, 0):
        sf.step_grad(s, 10 * s); o.step_grad(s, 10 * s)
    a, b = sf.get_grad(0), o.get_grad(0)
    for k in ('v', 'q', 'p'):
        assert rel(a[k], b[k]) < 1e-4, (k, rel(a[k], b[k]))
    gb = o.aircon_grad(0)
    ga = np.concatenate([air.gpos[0].numpy(), air.gquat[0].numpy(), [float(air.gs[0])], [float(air.gr[0])]])
    assert np.abs(ga - gb).max() < 1e-4 * np.abs(gb).max()
    sf.copy_grad(0, 4); sf.reset_grad_till_frame(2)
    assert np.array_equal(sf.get_grad(4)['v'], a['v']) and not np.any(sf.get_grad(0)['v']) and not np.any(sf.get_grad(1)['q'])


def test_smoke_timing_script_runs_on_the_emulated_library(emu, monkeypatch):
    """profiles/smoke_times.py (the per-phase timing script queued for the next GPU round) at a reduced size on the shim: script sanity only"""
    import sys
    import time
    import torch
    from fluidlab_b200 import smoke as smoke_mod
    monkeypatch.setattr(smoke_mod._lib, 'load', lambda: emu)
    sys.path.insert(0, os.path.join(ROOT, 'profiles'))
    import smoke_times

    class Ev:
        def __init__(self, **k): self.t = 0.0
        def record(self): self.t = time.perf_counter()
        def elapsed_time(self, o): return (o.t - self.t) * 1e3
    out = smoke_times.run(res=24, dev=torch.device('cpu'), reps=1, band=(8, 14), iters_list=(5,), event_cls=Ev, sync=lambda: None)
    ms = out['runs'][0]['ms']
    assert set(ms) >= {'step', 'step_grad', 'pressure', 'advect_grad'} and all(v > 0 for v in ms.values())


@pytest.mark.parametrize('band', [(3, 12), (2, 15)], ids=['H8_tiled', 'H12_sweep_per_launch'])
def test_other_band_heights_forward_and_adjoint(emu, band):
    """band of 8 layers: the largest the time-blocked tile kernel takes; 12 layers: the one-sweep-per-launch fallback (k_jacobi_simple).
    Same inputs as the reference-run fixture, other free band; against the oracle, forward and adjoint."""
    d = dict(np.load(os.path.join(GOLDEN, 'reference_smoke.npz')))
    fd = np.load(os.path.join(GOLDEN, 'reference_smoke_fd.npz'))
    d['lower_y'], d['higher_y'] = np.int64(band[0]), np.int64(band[1])
    e, o = _pair(emu, d, 11)
    for s in range(2):
        e.step(s, 10 * s); o.step(s, 10 * s)
    assert np.array_equal(e.a['is_free'][0].reshape((e.n,) * 3), o.is_free(0)) and o.is_free(0).sum() > 0
    a, b = e.get_state(2), o.get_state(2)
    for k in ('v', 'p', 'q'):
        assert rel(a[k], b[k]) < 1e-5, (band, k, rel(a[k], b[k]))
    o.reset_grad()
    z = o._alloc(); z['v'], z['q'], z['p'] = fd['w_v'], fd['w_q'], fd['w_p']
    o.set_grad(2, z); e.set_grad(2, z)
    for s in (1, 0):
        e.step_grad(s, 10 * s); o.step_grad(s, 10 * s)
    a, b = e.get_state(0, grad=True), o.get_grad(0)
    for k in ('v', 'q', 'p'):
        assert rel(a[k], b[k]) < 1e-4, (band, k, rel(a[k], b[k]))
