"""The smoke-solver oracle (oracle/smoke_oracle.hpp) against the reference's own kernels and against finite differences.

tests/golden/reference_smoke.npz holds a 3-step run of the UNMODIFIED fluidlab/fluidengine/simulators/smoke_field.py on the NumPy emulation
of the Taichi API (tests/golden/make_reference_smoke.py): free-space mask with two reference `Static` SDF volumes, RK3 advection,
air-conditioner impulse, divergence, Jacobi sweeps, projection.  reference_smoke_fd.npz holds central differences through that same
reference forward code run in float64 — the adjoint's pin."""
import os
import numpy as np
import pytest

from oracle.smoke import SmokeOracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-12))


def make(d, prec, max_steps_local=4):
    o = SmokeOracle(res=int(d['res']), dt=float(d['dt']), solver_iters=int(d['iters']), q_dim=int(d['q_dim']), max_steps_local=max_steps_local, max_substeps_local=40,
                    lower_y=int(d['lower_y']), higher_y=int(d['higher_y']), inject_v=tuple(d['inject_v']), precision=prec)
    for vox, T in zip(d['vox'], d['T_static']):
        o.add_static(vox, T)
    for f, a in zip(d['air_f'], d['air']):
        o.set_aircon(int(f), a)
    return o


@pytest.mark.parametrize('prec', [32, 64])
def test_oracle_reproduces_a_run_of_the_reference_smoke_kernels(prec):
    d = np.load(os.path.join(G, 'reference_smoke.npz'))
    o = make(d, prec)
    assert np.array_equal(o.get_state(0)['q'].astype(np.float32), d['q_init']), 'init_fields (SF:87-93)'
    o.set_state(0, {k: d['st0_' + k] for k in ('v', 'v_tmp', 'div', 'p', 'q')})
    for s in range(3):
        o.step(s, 10 * s)
        if s == 0:
            free = o.is_free(0)
            assert np.array_equal(free, d['free0']) and 0 < free.sum() < free.size
            band = (int(d['higher_y']) - int(d['lower_y']) - 1) * int(d['res']) ** 2
            assert free.sum() < band, 'the statics must block part of the band'
    tol = 2e-5 if prec == 32 else 1e-5     # the reference run is float32
    for s in (0, 1, 2):
        st = o.get_state(s)
        assert rel(st['v_tmp'], d[f'ref{s}_v_tmp']) < tol and rel(st['div'], d[f'ref{s}_div']) < 5 * tol, (s, rel(st['v_tmp'], d[f'ref{s}_v_tmp']), rel(st['div'], d[f'ref{s}_div']))
    for s in (1, 2, 3):
        st = o.get_state(s)
        for k in ('v', 'p', 'q'):
            assert rel(st[k], d[f'ref{s}_{k}']) < (5 * tol if k == 'p' else tol), (s, k, rel(st[k], d[f'ref{s}_{k}']))
    # the run must exercise everything: impulse, advection across cells, non-trivial pressure
    assert np.abs(d['ref3_p']).max() > 1e-2 and np.abs(d['ref3_q'] - d['st0_q']).max() > 0.1


def _fd_setup():
    d = np.load(os.path.join(G, 'reference_smoke.npz'))
    fd = np.load(os.path.join(G, 'reference_smoke_fd.npz'))
    o = make(d, 64)
    o.set_state(0, {k: d['st0_' + k].astype(np.float64) for k in ('v', 'v_tmp', 'div', 'p', 'q')})
    o.step(0, 0); o.step(1, 10)
    o.reset_grad()
    z = o._alloc(); z['v'], z['q'], z['p'] = fd['w_v'], fd['w_q'], fd['w_p']
    o.set_grad(2, z)
    o.step_grad(1, 10); o.step_grad(0, 0)
    return o, fd


def test_adjoint_matches_finite_differences_through_the_reference_forward_kernels():
    """d sum(w . state after 2 steps) / d (v0, q0, p0 along random directions; every air-conditioner parameter of both steps)"""
    o, fd = _fd_setup()
    g = o.get_grad(0)
    for k in ('v', 'q', 'p'):
        an = float((g[k] * fd['dir_' + k]).sum())
        assert abs(an - float(fd['fd_' + k])) < 2e-5 * abs(float(fd['fd_' + k])), (k, an, float(fd['fd_' + k]))
    for a_i, f in enumerate((0, 10)):
        ga = o.aircon_grad(f)
        ref = fd['fd_air'][a_i]
        assert np.abs(ref).min() > 1e-3
        assert np.abs(ga - ref).max() < 1e-6 * np.abs(ref).max(), (f, ga, ref)
