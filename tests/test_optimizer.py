"""Trajectory optimiser (on-device Adam; fluidlab/optimizer/optim.py:22-41, policies.py:131-164, solver.py:14-59).
Fixture tests/golden/reference_optim.npz: action tables produced by the reference's OWN Adam / TrainablePolicy classes
(tests/golden/make_reference_optim.py).  CPU: the oracle restatement and — on the CUDA execution-model shim — the product's device kernel
reproduce them bit for bit, and the Solver loop equals the same loop driven by the oracle optimiser.  `-m gpu`: the kernel on hardware."""
import os
import sys
import types
import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'cuda_emu'))
from conftest import make_particles  # noqa: E402

SCENES = ('latteart', 'pouring')


def _golden():
    return np.load(os.path.join(HERE, 'golden', 'reference_optim.npz'))


def _cfg(d, name):
    return types.SimpleNamespace(type='Adam', lr=float(d[f'{name}_lr']), beta_1=float(d['beta_1']), beta_2=float(d['beta_2']), epsilon=float(d['epsilon']))


@pytest.mark.parametrize('name', SCENES)
def test_oracle_reproduces_the_reference_optimiser_bit_for_bit(name):
    from oracle import optim as oo
    d = _golden()
    tables, grads = d[f'{name}_tables'], d[f'{name}_grads']
    c = _cfg(d, name)
    adam = oo.AdamOracle(tables[0].shape, c.lr, c.beta_1, c.beta_2, c.epsilon)
    fix = list(d[f'{name}_fix_dim']) or None
    av, ap = tables[0][:-1].copy(), tables[0][-1].copy()
    for it in range(len(grads)):
        av, ap = oo.policy_optimize(adam, av, ap, grads[it], tuple(d['action_range']), trainable=d[f'{name}_trainable'], fix_dim=fix)
        assert np.array_equal(np.vstack([av, ap[None]]), tables[it + 1]), it
    assert np.array_equal(adam.m, d[f'{name}_m']) and np.array_equal(adam.v, d[f'{name}_v'])
    assert np.abs(tables[-1] - tables[0]).max() > 1e-3          # the fixture moves


def _tiny_sim(device):
    from fluidlab_b200 import MPMSimulator, macros as M
    x = np.random.RandomState(0).uniform(0.4, 0.6, size=(40, 3)).astype(np.float32)
    P = make_particles(x, np.full(40, M.WATER), 16)
    s = MPMSimulator(dim=3, quality=0.25, gravity=(0, -10, 0), horizon=10, max_substeps_local=20, max_substeps_global=1000, ckpt_dest='cpu', device=device)
    s.setup_boundary(type='cube', lower=(0.1, 0.1, 0.1), upper=(0.9, 0.9, 0.9)); s.build(None, None, [], P)
    return s


def _check_product_against_golden(sim, name):
    from fluidlab_b200 import TrainablePolicy
    d = _golden()
    tables, grads = d[f'{name}_tables'], d[f'{name}_grads']
    H, D = tables[0].shape[0] - 1, tables[0].shape[1]
    pol = TrainablePolicy(_cfg(d, name), types.SimpleNamespace(v=(-0.05, 0.05), p=(0.4, 0.6)), D, H, tuple(d['action_range']), fix_dim=list(d[f'{name}_fix_dim']) or None)
    pol.actions_v, pol.actions_p = tables[0][:-1].copy(), tables[0][-1].copy()
    pol.trainable[:] = d[f'{name}_trainable']
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        pol.optimize(grads[0])
    pol.bind(sim)
    for it in range(len(grads)):
        g = grads[it] if it % 2 else torch.from_numpy(grads[it]).to(sim.device)     # host array and device tensor inputs
        pol.optimize(g, {})
        assert np.array_equal(pol.comp_actions, tables[it + 1]), (it, np.abs(pol.comp_actions - tables[it + 1]).max())
    assert np.array_equal(pol.optim.momentum_buffer.cpu().numpy(), d[f'{name}_m']) and np.array_equal(pol.optim.v_buffer.cpu().numpy(), d[f'{name}_v'])
    assert pol.optim.iter == len(grads)


TASKS = ('LatteArtStirPolicy', 'IceCreamDynamicPolicy', 'IceCreamStaticPolicy', 'TransportingPolicy')


def _check_task_policy(sim, name):
    """the rule-table policies against the reference's own classes: trainable rows, lr and freeze schedules over loss_info['temporal_range'],
    IceCreamStatic's gradient clip; tables bit for bit"""
    import fluidlab_b200
    d = _golden()
    tabs, lrs, trains = d[f'task_{name}_tables'], d[f'task_{name}_lr'], d[f'task_{name}_trainable']
    H, D = tabs[0].shape[0] - 1, tabs[0].shape[1]
    cfg = types.SimpleNamespace(type='Adam', lr=0.01, beta_1=0.9, beta_2=0.999, epsilon=1e-8)
    pol = getattr(fluidlab_b200, name)(cfg, types.SimpleNamespace(v=(-0.05, 0.05), p=(0.4, 0.6)), D, H, (-0.1, 0.1), fix_dim=None, sim=sim)
    pol.actions_v, pol.actions_p = tabs[0][:-1].copy(), tabs[0][-1].copy()
    assert np.array_equal(pol.trainable, trains[0])
    for it, tr in enumerate(d['task_tranges']):
        g = d['task_grads'][it]
        pol.optimize(g.copy() if it % 2 else torch.from_numpy(g).to(sim.device), {'temporal_range': int(tr)})
        assert np.array_equal(pol.comp_actions, tabs[it + 1]), (name, it)
        assert pol.optim.lr == lrs[it] and np.array_equal(pol.trainable, trains[it + 1]), (name, it)


PHASED = ('GatheringPolicy', 'GatheringOPolicy', 'MixingPolicy')


def _check_phased_policy(sim, name):
    """Gathering / GatheringO / Mixing: status / trainable layout, the scripted phases of get_action_v(update=True) (reading agent.rigid.latest_pos),
    freeze schedules and tables against the reference's own classes"""
    import fluidlab_b200
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    from make_reference_optim import fake_latest_pos
    d = _golden()
    scr, tabs, trains = d[f'phased_{name}_scripted'], d[f'phased_{name}_tables'], d[f'phased_{name}_trainable']
    H, D = tabs[0].shape[0] - 1, tabs[0].shape[1]
    cfg = types.SimpleNamespace(type='Adam', lr=0.002, beta_1=0.9, beta_2=0.999, epsilon=1e-8)
    pol = getattr(fluidlab_b200, name)(cfg, types.SimpleNamespace(v=(-0.005, 0.005), p=(0.4, 0.6)), D, H, (-0.01, 0.01), fix_dim=None, sim=sim)
    pol.actions_v, pol.actions_p = tabs[0][:-1].copy(), tabs[0][-1].copy()
    assert np.array_equal(pol.status, d[f'phased_{name}_status']) and np.array_equal(pol.trainable, trains[0])
    for it, tr in enumerate(d['phased_tranges']):
        for i in range(H):
            agent = types.SimpleNamespace(rigid=types.SimpleNamespace(latest_pos=types.SimpleNamespace(to_numpy=lambda it=it, i=i: fake_latest_pos(it, i))))
            pol.get_action_v(i, agent=agent, update=True)
        assert np.array_equal(pol.comp_actions, scr[it]), (name, it)
        pol.optimize(d['phased_grads'][it].copy(), {'temporal_range': int(tr)})
        assert np.array_equal(pol.comp_actions, tabs[it + 1]), (name, it)
        assert np.array_equal(pol.trainable, trains[it + 1]) and pol.freeze_till == int(d[f'phased_{name}_freeze'][it]), (name, it)


@pytest.mark.parametrize('name', PHASED)
def test_scripted_phase_policies_equal_the_reference_classes_on_the_shim(emu, name):
    _check_phased_policy(_tiny_sim('cpu'), name)


def test_effector_latest_pos_follows_the_last_moved_substep(emu):
    """effector.py:146-152: latest_pos = pos[f] of the step's last move; zeros before any move"""
    env, n_steps = _latteart_env('cpu')
    eff = env.agent.effectors[0]
    assert np.array_equal(eff.latest_pos.to_numpy(), np.zeros((1, 3), np.float32))
    env.set_state(env.get_state()['state'], grad_enabled=False)
    env.step(np.array([0.004, 0.0, -0.002], np.float32))
    n = env.simulator.n_substeps
    assert np.array_equal(eff.latest_pos.to_numpy()[0], eff.get_state(n - 1)[:3]) and not np.array_equal(eff.get_state(n - 1)[:3], eff.get_state(0)[:3])


@pytest.fixture
def emu():
    import harness
    harness.enable()
    yield harness
    harness.disable()


@pytest.mark.parametrize('name', SCENES)
def test_device_adam_reproduces_the_reference_bit_for_bit_on_the_shim(emu, name):
    _check_product_against_golden(_tiny_sim('cpu'), name)


@pytest.mark.parametrize('name', TASKS)
def test_task_policies_equal_the_reference_classes_on_the_shim(emu, name):
    _check_task_policy(_tiny_sim('cpu'), name)


@pytest.mark.gpu
@pytest.mark.parametrize('name', SCENES)
def test_device_adam_reproduces_the_reference_bit_for_bit(name):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    _check_product_against_golden(_tiny_sim(None), name)


@pytest.mark.gpu
@pytest.mark.parametrize('name', TASKS)
def test_task_policies_equal_the_reference_classes(name):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    _check_task_policy(_tiny_sim(None), name)


def _latteart_env(device):
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
    tgt = [np.tile(np.array([0.55, 0.4, 0.45], np.float32), (len(x), 1)) for _ in range(n_steps)]
    env = TaichiEnv(quality=n_grid / 64, max_substeps_local=T, gravity=(0.0, -20.0, 0.0), horizon=n_steps, ckpt_dest='cpu', device=device)
    env.simulator.use_graphs = False
    np.random.seed(5)
    env.setup_agent(cfg)
    env.particle_bodies.get = lambda: P
    env.setup_boundary(**bnd)
    env.setup_loss(loss_cls=LatteArtLoss, type='diff', target=tgt, weights={'chamfer': 1.0})
    env.build()
    return env, n_steps


def test_solver_loop_equals_the_same_loop_with_the_oracle_optimiser(emu):
    """Solver.solve (solver.py:14-59) over a miniature LatteArt rollout: device gradients -> device Adam -> host mirror, 4 iterations; the final
    action table equals the one obtained with agent.get_grad (host) + the oracle's NumPy Adam, and the loss goes down"""
    from fluidlab_b200 import Solver, TrainablePolicy, forward_backward
    from oracle import optim as oo
    env, n_steps = _latteart_env('cpu')
    ocfg = types.SimpleNamespace(type='Adam', lr=0.01, beta_1=0.9, beta_2=0.999, epsilon=1e-8)
    irange = types.SimpleNamespace(v=(-0.004, 0.004), p=(0.45, 0.55))

    def make_policy(*_):
        np.random.seed(3)
        return TrainablePolicy(ocfg, irange, 3, n_steps, (-0.02, 0.02), fix_dim=[1])
    from fluidlab_b200 import trainable_policy
    np.random.seed(3)
    bound = trainable_policy(env, TrainablePolicy, ocfg, irange, n_steps, (-0.02, 0.02), fix_dim=[1])
    assert bound._sim is env.simulator and np.array_equal(bound.comp_actions, make_policy().comp_actions)
    wrapper = types.SimpleNamespace(taichi_env=env, horizon=n_steps, horizon_action=n_steps, trainable_policy=make_policy)
    losses = []
    pol = Solver(wrapper, cfg=types.SimpleNamespace(optim=ocfg, init_range=irange, n_iters=4)).solve(callback=lambda it, info: losses.append(info['loss']))
    final = pol.comp_actions.copy()
    assert losses[-1] < losses[0], losses
    # the same loop with the oracle optimiser on host gradients
    ref = make_policy()
    adam = oo.AdamOracle(ref.comp_actions_shape, ocfg.lr, ocfg.beta_1, ocfg.beta_2, ocfg.epsilon)
    env2, _ = _latteart_env('cpu')
    st0 = env2.get_state()['state']
    losses2 = []
    for it in range(4):
        info, g = forward_backward(env2, st0, ref, n_steps, device_grad=False)
        losses2.append(info['loss'])
        ref.actions_v, ref.actions_p = oo.policy_optimize(adam, ref.actions_v, ref.actions_p, g, ref.action_range, trainable=ref.trainable, fix_dim=ref.fix_dim)
    # the two rollouts are separate simulations (the order of the scatter atomics is not reproducible), so their gradients agree to fp32
    # round-off, not bit for bit; the optimiser's own bit-exactness is pinned by the golden tests above
    assert np.abs(final - ref.comp_actions).max() < 1e-6, np.abs(final - ref.comp_actions).max()
    assert np.allclose(losses, losses2, rtol=1e-5, atol=0)
    assert np.array_equal(final[:, 1], make_policy().comp_actions[:, 1])      # fix_dim column untouched
