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
    dev = torch.device('cuda', 0)
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
    sim = types.SimpleNamespace(max_steps_local=4, agent=agent, device=dev, statics=statics, _stream=lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
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
    o = _oracle(d, prec=64)
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
    8-component action, 10 parked MPM particles, CirculationLoss): 3 steps forward + backward through TaichiEnv; the smoke state must
    equal the oracle fed with the SAME air-conditioner trajectory, and the air conditioner's adjoints / the strength and radius
    components of dLoss/dAction must equal the oracle's."""
    _need_gpu()
    from fluidlab_b200 import TaichiEnv, CirculationLoss, macros as M
    env = TaichiEnv(dim=3, particle_density=1e6, max_substeps_local=100, gravity=(0.0, -20.0, 0.0), horizon=20, ckpt_dest='gpu')
    env.setup_agent(dict(type='AgentCirculation', effectors=[dict(type='AirCon', params=dict(init_pos=(0.8, 0.8, 0.5), action_dim=8, action_scale_p=(1.0,) * 8,
                                                                                            action_scale_v=(1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 100000.0, 50.0)),
                                                                 boundary=dict(type='cube', lower=(0.05, 0.05, 0.05), upper=(0.95, 0.95, 0.95)))]))
    env.add_body(type='nowhere', n_particles=10, material=M.WATER)
    env.setup_smoke_field(res=128, dt=0.03, solver_iters=50, decay=0.99, q_dim=1)
    env.setup_loss(loss_cls=CirculationLoss, type='diff', weights={'temp': 1.0})
    env.build()
    sf, air = env.smoke_field, env.agent.aircon
    env.apply_agent_action_p(np.array([0.55, 0.5, 0.27, 0.0, 0.0, 0.0, 0.0, 0.0]))      # demo_policy, circulation_env.py:113-120
    act = np.array([0.01, 0.0, 0.005, 0.0, 0.1, 0.0, 0.02, 0.04])
    env.set_state(env.get_state()['state'], grad_enabled=True)
    n_steps = 3
    for _ in range(n_steps):
        env.step(act)
    o = SmokeOracle(res=128, dt=0.03, solver_iters=50, q_dim=1, max_steps_local=10, max_substeps_local=100, inject_v=tuple(air.inject_v), precision=32)
    for s in range(n_steps):
        f = 10 * s
        o.set_aircon(f, np.concatenate([air.pos[f].cpu().numpy(), air.quat[f].cpu().numpy(), [float(air.s[f])], [float(air.r[f])]]))
        o.step(s, f)
    assert abs(float(air.s[0]) - 0.02 * 100000.0) < 1e-2 and abs(float(air.r[0]) - 0.04 * 50.0) < 1e-5
    for s in range(1, n_steps + 1):
        a, b = sf.get_state(s), o.get_state(s)
        for k in ('v', 'q', 'p'):
            assert rel(a[k], b[k]) < 2e-5, (s, k, rel(a[k], b[k]))
    assert np.abs(sf.get_state(n_steps)['v']).max() > 1e-3, 'the air conditioner must move the air'
    info = env.get_final_loss()
    # backward
    env.reset_grad(); env.get_final_loss_grad()
    for _ in range(n_steps):
        env.step_grad(act)
    # the oracle's backward with the same loss seeds (sign(q - target) at the detectors of every step frame 1..n_steps)
    from fluidlab_b200.losses import CirculationLoss as CL
    o.reset_grad()
    tgt = [1.0] * 5 + [0.0] * 10
    loss_o = 0.0
    for s in range(n_steps, 0, -1):
        q = o.get_state(s)['q']
        g = o.get_grad(s)
        for (x, zc), t in zip(CL.DETECTORS, tgt):
            g['q'][x, 64, zc, 0] += np.sign(q[x, 64, zc, 0] - t)
            loss_o += abs(q[x, 64, zc, 0] - t)
        o.set_grad(s, g)
        o.step_grad(s - 1, 10 * (s - 1))
    assert abs(info['loss'] - loss_o) < 1e-4 * abs(loss_o), (info['loss'], loss_o)
    gs_sum = gr_sum = 0.0
    for s in range(n_steps):
        f = 10 * s
        gb = o.aircon_grad(f)
        ga = np.concatenate([air.gpos[f].cpu().numpy() * 0, air.gquat[f].cpu().numpy() * 0, [float(air.gs[f])], [float(air.gr[f])]])
        assert abs(ga[7] - gb[7]) <= 1e-3 * max(abs(gb[7]), 1e-6) and abs(ga[8] - gb[8]) <= 1e-3 * max(abs(gb[8]), 1e-6), (s, ga[7:], gb[7:])
    grad = env.agent.get_grad(n_steps)
    assert grad.shape == (n_steps + 1, 8)
    for s in range(n_steps):
        gb = o.aircon_grad(10 * s)
        assert abs(grad[s, 6] - gb[7] * 100000.0) <= 1e-3 * max(abs(gb[7] * 100000.0), 1e-6), (s, grad[s, 6], gb[7] * 100000.0)
        assert abs(grad[s, 7] - gb[8] * 50.0) <= 1e-3 * max(abs(gb[8] * 50.0), 1e-6), (s, grad[s, 7], gb[8] * 50.0)
