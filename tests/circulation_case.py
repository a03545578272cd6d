"""The air-circulation stack of envs/circulation_env.py driven through the product's TaichiEnv and checked against the smoke oracle.
Shared by the GPU test (tests/test_zz_smoke_gpu.py, real size) and the CUDA-execution-model test (tests/test_cuda_emu_mpm.py, reduced size)."""
import numpy as np

from oracle.smoke import SmokeOracle
from test_smoke_oracle import rel


def run_circulation_stack(device=None, res=128, iters=50, band=None, detectors=None, detector_h=None, n_steps=3, max_substeps_local=100, ring_wraps=False):
    """AgentCirculation + AirCon (8-component action) + SmokeField + CirculationLoss + 10 parked MPM particles: n_steps forward and backward
    through TaichiEnv.  The smoke state must equal the oracle fed with the SAME air-conditioner trajectory; the loss, the air conditioner's
    strength / radius adjoints and components 6, 7 of dLoss/dAction must equal the oracle's.
    ring_wraps: max_substeps_local is shorter than the trajectory, so the MPM ring AND the smoke field's step ring wrap (memory_to_cache:
    smoke checkpoint + copy_frame; backward: memory_from_cache re-runs the chunk incl. the smoke steps, mpm_simulator.py:777-912); only
    quantities that survive the wrap are compared then (final smoke state, loss, dLoss/dAction)."""
    from fluidlab_b200 import TaichiEnv, CirculationLoss, macros as M
    from fluidlab_b200.losses import CirculationLoss as CL
    env = TaichiEnv(dim=3, particle_density=1e6, max_substeps_local=max_substeps_local, gravity=(0.0, -20.0, 0.0), horizon=20, ckpt_dest='gpu' if device is None else 'cpu',
                    device=device)
    env.simulator.use_graphs = device is None
    env.setup_agent(dict(type='AgentCirculation', effectors=[dict(type='AirCon', params=dict(init_pos=(0.8, 0.8, 0.5), action_dim=8, action_scale_p=(1.0,) * 8,
                                                                                            action_scale_v=(1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 100000.0, 50.0)),
                                                                 boundary=dict(type='cube', lower=(0.05, 0.05, 0.05), upper=(0.95, 0.95, 0.95)))]))
    env.add_body(type='nowhere', n_particles=10, material=M.WATER)
    env.setup_smoke_field(res=res, dt=0.03, solver_iters=iters, decay=0.99, q_dim=1)
    if band is not None:
        env.smoke_field.lower_y, env.smoke_field.higher_y = band
    env.setup_loss(loss_cls=CirculationLoss, type='diff', weights={'temp': 1.0})
    if detectors is not None:
        env.loss.DETECTORS = detectors; env.loss.detector_h = detector_h
    env.build()
    sf, air = env.smoke_field, env.agent.aircon
    dets, dh = env.loss.DETECTORS, env.loss.detector_h
    env.apply_agent_action_p(np.array([0.55, 0.5, 0.27, 0.0, 0.0, 0.0, 0.0, 0.0]))      # demo_policy, circulation_env.py:113-120
    act = np.array([0.01, 0.0, 0.005, 0.0, 0.1, 0.0, 0.02, 0.04])
    env.set_state(env.get_state()['state'], grad_enabled=True)
    air_traj = []
    for _ in range(n_steps):
        f0 = env.simulator.cur_substep_local
        env.step(act)
        air_traj.append(np.concatenate([air.pos[f0].cpu().numpy(), air.quat[f0].cpu().numpy(), [float(air.s[f0])], [float(air.r[f0])]]))   # what the smoke step of this step saw
    kw = {} if band is None else dict(lower_y=band[0], higher_y=band[1])
    o = SmokeOracle(res=res, dt=0.03, solver_iters=iters, q_dim=1, max_steps_local=n_steps + 1, max_substeps_local=10 * (n_steps + 1),
                    inject_v=tuple(air.inject_v), precision=32, **kw)
    for s in range(n_steps):
        o.set_aircon(10 * s, air_traj[s])
        o.step(s, 10 * s)
    if ring_wraps:
        assert n_steps * 10 > max_substeps_local
        a, b = sf.get_state(env.simulator.cur_step_local), o.get_state(n_steps)
        for k in ('v', 'q', 'p'):
            assert rel(a[k], b[k]) < 2e-5, ('final', k, rel(a[k], b[k]))
    assert abs(air_traj[0][7] - 0.02 * 100000.0) < 1e-2 and abs(air_traj[0][8] - 0.04 * 50.0) < 1e-5
    assert np.abs(air_traj[1][:3] - air_traj[0][:3]).max() > 1e-3 and abs(air_traj[1][3] - 1.0) > 1e-4, 'the pose chain must move and rotate'
    if not ring_wraps:
        for s in range(1, n_steps + 1):
            a, b = sf.get_state(s), o.get_state(s)
            for k in ('v', 'q', 'p'):
                assert rel(a[k], b[k]) < 2e-5, (s, k, rel(a[k], b[k]))
    assert np.abs(o.get_state(n_steps)['v']).max() > 1e-3, 'the air conditioner must move the air'
    st = env.get_state()['state']
    assert st['smoke_field']['q'].shape == (res, res, res, 1) and len(st['agent'][0]) == 9
    info = env.get_final_loss()
    # backward
    env.reset_grad(); env.get_final_loss_grad()
    for _ in range(n_steps):
        env.step_grad(act)
    # the oracle's backward with the same loss seeds (sign(q - target) at the detectors of every step frame 1..n_steps)
    o.reset_grad()
    tgt = [1.0] * 5 + [0.0] * 10
    loss_o = 0.0
    for s in range(n_steps, 0, -1):
        q = o.get_state(s)['q']
        g = o.get_grad(s)
        for (x, zc), t in zip(dets, tgt):
            g['q'][x, dh, zc, 0] += np.sign(q[x, dh, zc, 0] - t)
            loss_o += abs(q[x, dh, zc, 0] - t)
        o.set_grad(s, g)
        o.step_grad(s - 1, 10 * (s - 1))
    assert abs(info['loss'] - loss_o) < 1e-4 * abs(loss_o), (info['loss'], loss_o)
    grad = env.agent.get_grad(n_steps)
    assert grad.shape == (n_steps + 1, 8)
    seen = 0.0
    for s in range(n_steps):
        f = 10 * s
        gb = o.aircon_grad(f)
        if not ring_wraps:   # per-frame adjoints of the effector are ring-local
            ga = np.concatenate([air.gpos[f].cpu().numpy(), air.gquat[f].cpu().numpy(), [float(air.gs[f])], [float(air.gr[f])]])
            scale = max(np.abs(gb[7:]).max(), 1e-12)
            assert np.abs(ga[7:] - gb[7:]).max() <= 1e-3 * scale, (s, ga[7:], gb[7:])
        assert abs(grad[s, 6] - gb[7] * 100000.0) <= 1e-3 * max(abs(gb[7] * 100000.0), 1e-9), (s, grad[s, 6], gb[7] * 100000.0)
        assert abs(grad[s, 7] - gb[8] * 50.0) <= 1e-3 * max(abs(gb[8] * 50.0), 1e-9), (s, grad[s, 7], gb[8] * 50.0)
        seen = max(seen, abs(gb[7]), abs(gb[8]))
    assert seen > 0, 'the loss must depend on the air conditioner'
    assert np.abs(grad[:n_steps, :6]).max() > 0, 'the pose components of dLoss/dAction must be populated (pose adjoints flow through the effector chain)'
    return env, o


def run_reference_stack_case(device=None):
    """tests/golden/reference_circulation.npz: the reference's OWN MPMSimulator + AgentCirculation + AirCon + SmokeField (10 parked
    particles, a Static 'room'), 3 x MPMSimulator.step(action) on the Taichi emulation (tests/golden/make_reference_smoke.py stack).  The
    product's TaichiEnv stack on the emulated device must reproduce the air conditioner's trajectory (pose chain with rotation, strength
    and radius channels of the 8-component action, apply_action_p) and the smoke state after every step."""
    from fluidlab_b200 import TaichiEnv, macros as M
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_circulation.npz'))
    env = TaichiEnv(dim=3, quality=0.25, particle_density=1e6, max_substeps_local=int(d['T']), gravity=(0.0, -20.0, 0.0), horizon=10, ckpt_dest='gpu' if device is None else 'cpu', device=device)
    env.simulator.use_graphs = device is None
    env.setup_agent(dict(type='AgentCirculation', effectors=[dict(type='AirCon', params=dict(init_pos=tuple(d['init_pos']), action_dim=8, action_scale_p=(1.0,) * 8,
                                                                                            action_scale_v=tuple(d['scale_v']), inject_v=tuple(d['inject_v'])),
                                                                 boundary=dict(type='cube', lower=tuple(d['e_lower']), upper=tuple(d['e_upper'])))]))
    env.add_static(file='room.obj', material=M.PILLAR, has_dynamics=True, sdf=dict(voxels=d['room_vox'], T_mesh_to_voxels=d['room_T']))
    env.add_body(type='nowhere', n_particles=10, material=M.WATER)
    env.setup_smoke_field(res=int(d['res']), dt=float(d['dt']), solver_iters=int(d['iters']), decay=0.99, q_dim=int(d['q_dim']))
    env.smoke_field.lower_y, env.smoke_field.higher_y = int(d['lower_y']), int(d['higher_y'])
    env.build()
    env.apply_agent_action_p(d['action_p'])
    for a in d['actions']:
        env.step(a)
    air, sf = env.agent.aircon, env.smoke_field
    nf = 10 * len(d['actions'])
    assert rel(air.pos[:nf + 1].cpu().numpy(), d['ref_pos']) < 1e-6 and rel(air.quat[:nf + 1].cpu().numpy(), d['ref_quat']) < 1e-6
    assert rel(air.s[:nf].cpu().numpy(), d['ref_s']) < 1e-6 and rel(air.r[:nf].cpu().numpy(), d['ref_r']) < 1e-6
    assert np.abs(d['ref_quat'][nf][1:]).max() > 0.05 and np.ptp(d['ref_s']) > 0, 'the reference trajectory must rotate and vary its strength'
    assert rel(env.agent.get_state(nf)[0], d['ref_state']) < 1e-6
    for s in (1, 2, 3):
        st = sf.get_state(s)
        for k in ('v', 'p', 'q'):
            assert rel(st[k], d[f'ref{s}_{k}']) < 2e-5, (s, k, rel(st[k], d[f'ref{s}_{k}']))
    assert np.abs(d['ref3_v']).max() > 1.0
