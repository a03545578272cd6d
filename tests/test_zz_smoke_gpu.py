"""GPU leg of the smoke solver: fluidlab_b200.smoke.SmokeField and the air-circulation stack on the real library.

Written after this round's GPU budget was spent: the kernels and the host class are verified on CPU through the CUDA execution-model
shim (tests/test_smoke_cuda_emu.py) but these tests have not run on a B200 yet.  The file sorts last so that a failure here cannot mask
the validated MPM parity tests under `pytest -x`."""
import os
import types
import numpy as np
import pytest
import torch

from oracle.smoke import SmokeOracle
from test_smoke_oracle import G as GOLDEN, rel

pytestmark = pytest.mark.gpu


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail('no CUDA device: the -m gpu tests must run on the B200 box')


def _oracle(d, iters=None, prec=32):
    o = SmokeOracle(res=int(d['res']), dt=float(d['dt']), solver_iters=int(d['iters']) if iters is None else iters, q_dim=int(d['q_dim']), max_steps_local=4,
                    max_substeps_local=40, lower_y=int(d['lower_y']), higher_y=int(d['higher_y']), inject_v=tuple(d['inject_v']), precision=prec)
    for vox, T in zip(d['vox'], d['T_static']):
        o.add_static(vox, T)
    for f, a in zip(d['air_f'], d['air']):
        o.set_aircon(int(f), a)
    return o


def _field(d, iters=None):
    from fluidlab_b200 import smoke as smoke_mod, meshes, macros as M
    emu = os.environ.get('FLUIDLAB_CUDA_EMU') == '1'      # development aid: conftest.py routes the library to the CPU execution-model shim
    dev = torch.device('cpu') if emu else torch.device('cuda', 0)
    Tsub = 40
    z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
    air = types.SimpleNamespace(pos=z(Tsub + 1, 3), quat=z(Tsub + 1, 4), s=z(Tsub + 1), r=z(Tsub + 1), gpos=z(Tsub + 1, 3), gquat=z(Tsub + 1, 4), gs=z(Tsub + 1), gr=z(Tsub + 1),
                                inject_v=np.asarray(d['inject_v']))
    air.quat[:, 0] = 1
    for f, a in zip(d['air_f'], d['air']):
        a = torch.from_numpy(a.astype(np.float32)).to(dev); air.pos[int(f)] = a[:3]; air.quat[int(f)] = a[3:7]; air.s[int(f)] = a[7]; air.r[int(f)] = a[8]
    statics = meshes.Statics()
    for vox, T in zip(d['vox'], d['T_static']):
        statics.add_static(file='x.obj', material=M.PILLAR, has_dynamics=True, sdf=dict(voxels=vox, T_mesh_to_voxels=T))
    agent = types.SimpleNamespace(aircon=air)
    import ctypes as C
    sim = types.SimpleNamespace(max_steps_local=4, agent=agent, device=dev, statics=statics, _stream=(lambda: None) if emu else (lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    sf = smoke_mod.SmokeField(dim=3, ckpt_dest='gpu', res=int(d['res']), dt=float(d['dt']), solver_iters=int(d['iters']) if iters is None else iters, q_dim=int(d['q_dim']))
    sf.lower_y, sf.higher_y = int(d['lower_y']), int(d['higher_y'])
    sf.build(sim, agent)
    return sf, air


@pytest.mark.parametrize('iters', [6, 21])
def test_smoke_forward_matches_the_oracle_and_the_reference_run(iters):
    _need_gpu()
    d = np.load(os.path.join(GOLDEN, 'reference_smoke.npz'))
    sf, _ = _field(d, iters)
    o = _oracle(d, iters)
    st0 = {k: d['st0_' + k] for k in ('v', 'v_tmp', 'div', 'p', 'q')}
    sf.set_state(0, st0); o.set_state(0, st0)
    for s in range(3):
        sf.step(s, 10 * s); o.step(s, 10 * s)
    assert np.array_equal(sf.is_free(0), o.is_free(0))
    for s in (1, 2, 3):
        a, b = sf.get_state(s), o.get_state(s)
        for k in ('v', 'p', 'q'):
            assert rel(a[k], b[k]) < 1e-5, (s, k, rel(a[k], b[k]))
    if iters == 6:
        a = sf.get_state(3)
        for k in ('v', 'p', 'q'):
            assert rel(a[k], d['ref3_' + k]) < 2e-5, (k, rel(a[k], d['ref3_' + k]))


def test_smoke_backward_matches_the_oracle_and_reference_finite_differences():
    _need_gpu()
    d = np.load(os.path.join(GOLDEN, 'reference_smoke.npz'))
    fd = np.load(os.path.join(GOLDEN, 'reference_smoke_fd.npz'))
    sf, air = _field(d)
    o = _oracle(d, prec=32)     # float32 like the device; the float64 side of the pin is the finite-difference fixture below
    st0 = {k: d['st0_' + k] for k in ('v', 'v_tmp', 'div', 'p', 'q')}
    sf.set_state(0, st0); o.set_state(0, st0)
    for s in range(2):
        sf.step(s, 10 * s); o.step(s, 10 * s)
    zz = o._alloc(); zz['v'], zz['q'], zz['p'] = fd['w_v'], fd['w_q'], fd['w_p']
    o.reset_grad(); o.set_grad(2, zz)
    sf.reset_grad(); sf.set_grad(2, zz)
    for s in (1, 0):
        sf.step_grad(s, 10 * s); o.step_grad(s, 10 * s)
    a, b = sf.get_grad(0), o.get_grad(0)
    for k in ('v', 'q', 'p'):
        assert rel(a[k], b[k]) < 1e-4, (k, rel(a[k], b[k]))
        an = float((a[k].astype(np.float64) * fd['dir_' + k]).sum())
        assert abs(an - float(fd['fd_' + k])) < 2e-3 * abs(float(fd['fd_' + k])), (k, an, float(fd['fd_' + k]))
    for f in (0, 10):
        ga = np.concatenate([air.gpos[f].cpu().numpy(), air.gquat[f].cpu().numpy(), [float(air.gs[f])], [float(air.gr[f])]])
        gb = o.aircon_grad(f)
        assert np.abs(ga - gb).max() < 1e-4 * np.abs(gb).max(), (f, ga, gb)


def test_circulation_stack_through_taichi_env():
    """envs/circulation_env.py's stack at its real size (128^3 smoke grid, 50 Jacobi sweeps, q_dim 1, AgentCirculation + AirCon with the
    8-component action, 10 parked MPM particles, CirculationLoss): 3 steps forward + backward through TaichiEnv; see
    tests/circulation_case.py for what is compared.  The same case runs on CPU at a reduced size in tests/test_cuda_emu_mpm.py."""
    _need_gpu()
    from circulation_case import run_circulation_stack
    run_circulation_stack(device=None)
    # reduced size with a ring of 2 steps for 3 steps: MPM and smoke rings wrap, the backward pass re-runs the first chunk
    dets = [[5, 16], [7, 16], [3, 16], [5, 14], [5, 18], [5, 8], [7, 8], [3, 8], [5, 6], [5, 10], [20, 12], [21, 12], [18, 12], [20, 9], [20, 16]]
    run_circulation_stack(device=None, res=24, iters=10, band=(8, 14), detectors=dets, detector_h=11, n_steps=3, max_substeps_local=20, ring_wraps=True)


def test_circulation_stack_equals_a_run_of_the_real_reference_stack():
    """tests/golden/reference_circulation.npz: the reference's OWN MPMSimulator + AgentCirculation + AirCon + SmokeField stepped on the Taichi
    emulation; the product's TaichiEnv stack on the B200 must reproduce the air conditioner's trajectory and the smoke states."""
    _need_gpu()
    from circulation_case import run_reference_stack_case
    run_reference_stack_case(device=None)
