"""Generates tests/golden/reference_optim.npz with the reference's OWN optimiser classes: Adam from fluidlab/optimizer/optim.py (loaded by
path: it imports NumPy only) and TrainablePolicy from fluidlab/optimizer/policies.py (its class source is compiled on its own — the module
imports input-device packages that are not installed).  Run in the build container (needs /root/reference); the fixture travels.

    python tests/golden/make_reference_optim.py
"""
import ast
import importlib.util
import os
import types
import numpy as np

REF = '/root/reference/fluidlab/optimizer'
TASK_POLICIES = ('LatteArtStirPolicy', 'IceCreamDynamicPolicy', 'IceCreamStaticPolicy', 'TransportingPolicy')
PHASED_POLICIES = ('GatheringPolicy', 'GatheringOPolicy', 'MixingPolicy')
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_classes():
    spec = importlib.util.spec_from_file_location('ref_optim', os.path.join(REF, 'optim.py'))
    optim = importlib.util.module_from_spec(spec); spec.loader.exec_module(optim)
    tree = ast.parse(open(os.path.join(REF, 'policies.py')).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ('TrainablePolicy',) + TASK_POLICIES + PHASED_POLICIES]
    ns = {'np': np, 'Adam': optim.Adam, 'print': lambda *a, **k: None}
    exec(compile(ast.Module(body=cls, type_ignores=[]), 'policies.py', 'exec'), ns)
    return optim.Adam, ns['TrainablePolicy'], ns


def fake_latest_pos(it, i):
    return np.array([[0.3 + 0.001 * i + 0.01 * it, 0.5 - 0.0005 * i, 0.6 - 0.0007 * i]], np.float32)


def main():
    Adam, TrainablePolicy, ns = reference_classes()
    out = {}
    for name, (H, D, fix_dim, frozen, lr) in dict(latteart=(12, 3, None, 0, 0.05), pouring=(9, 6, [1, 4], 3, 0.003)).items():
        cfg = types.SimpleNamespace(type='Adam', lr=lr, beta_1=0.9, beta_2=0.999, epsilon=1e-8)
        rng = types.SimpleNamespace(v=(-0.05, 0.05), p=(0.4, 0.6))
        np.random.seed(7)
        pol = TrainablePolicy(cfg, rng, D, H, (-0.1, 0.1), fix_dim=fix_dim)
        pol.trainable[:frozen] = False
        rs = np.random.RandomState(11)
        n_iters = 6
        grads = (rs.randn(n_iters, H + 1, D) * 10.0 ** rs.uniform(-4, 1, size=(n_iters, 1, 1))).astype(np.float32)
        grads[2, 1] = 0.0                      # a zero gradient row: sqrt(v) + eps path
        tables = [pol.comp_actions.copy()]
        for it in range(n_iters):
            pol.optimize(grads[it].copy(), {})
            tables.append(pol.comp_actions.copy())
        out.update({f'{name}_grads': grads, f'{name}_tables': np.stack(tables), f'{name}_trainable': pol.trainable.copy(),
                    f'{name}_fix_dim': np.array([] if fix_dim is None else fix_dim, np.int32), f'{name}_lr': lr,
                    f'{name}_m': pol.optim.momentum_buffer.copy(), f'{name}_v': pol.optim.v_buffer.copy()})
    # task policies (rule tables on loss_info['temporal_range']): horizon 180, temporal ranges that cross their lr / freeze thresholds
    H, D = 180, 3
    tranges = np.array([90, 120, 160, 210, 260, 460], np.int32)
    rs = np.random.RandomState(5)
    tgrads = (rs.randn(len(tranges), H + 1, D) * 10.0 ** rs.uniform(-2, 7, size=(len(tranges), 1, 1))).astype(np.float32)   # some beyond the 1e5 clip
    for name in TASK_POLICIES:
        cfg = types.SimpleNamespace(type='Adam', lr=0.01, beta_1=0.9, beta_2=0.999, epsilon=1e-8)
        np.random.seed(9)
        pol = ns[name](cfg, types.SimpleNamespace(v=(-0.05, 0.05), p=(0.4, 0.6)), D, H, (-0.1, 0.1), fix_dim=None)
        tabs, lrs, trains = [pol.comp_actions.copy()], [], [pol.trainable.copy()]
        for it, tr in enumerate(tranges):
            pol.optimize(tgrads[it].copy(), {'temporal_range': int(tr)})
            tabs.append(pol.comp_actions.copy()); lrs.append(pol.optim.lr); trains.append(pol.trainable.copy())
        out.update({f'task_{name}_tables': np.stack(tabs), f'task_{name}_lr': np.array(lrs), f'task_{name}_trainable': np.stack(trains)})
    out.update(task_grads=tgrads, task_tranges=tranges)
    # scripted-phase policies: get_action_v(i, agent, update=True) over the horizon with a stand-in agent whose latest position follows
    # fake_latest_pos(it, i), then optimize
    H, D = 250, 3
    ptr = np.array([100, 130, 250, 370], np.int32)
    rs = np.random.RandomState(6)
    pgrads = (rs.randn(len(ptr), H + 1, D) * 0.1).astype(np.float32)
    for name in PHASED_POLICIES:
        cfg = types.SimpleNamespace(type='Adam', lr=0.002, beta_1=0.9, beta_2=0.999, epsilon=1e-8)
        np.random.seed(10)
        pol = ns[name](cfg, types.SimpleNamespace(v=(-0.005, 0.005), p=(0.4, 0.6)), D, H, (-0.01, 0.01), fix_dim=None)
        scripted, tabs, trains, freezes = [], [pol.comp_actions.copy()], [pol.trainable.copy()], []
        for it, tr in enumerate(ptr):
            for i in range(H):
                agent = types.SimpleNamespace(rigid=types.SimpleNamespace(latest_pos=types.SimpleNamespace(to_numpy=lambda it=it, i=i: fake_latest_pos(it, i))))
                pol.get_action_v(i, agent=agent, update=True)
            scripted.append(pol.comp_actions.copy())
            pol.optimize(pgrads[it].copy(), {'temporal_range': int(tr)})
            tabs.append(pol.comp_actions.copy()); trains.append(pol.trainable.copy()); freezes.append(pol.freeze_till)
        out.update({f'phased_{name}_scripted': np.stack(scripted), f'phased_{name}_tables': np.stack(tabs), f'phased_{name}_trainable': np.stack(trains),
                    f'phased_{name}_freeze': np.array(freezes), f'phased_{name}_status': pol.status.copy()})
    out.update(phased_grads=pgrads, phased_tranges=ptr)
    np.savez_compressed(os.path.join(HERE, 'reference_optim.npz'), beta_1=0.9, beta_2=0.999, epsilon=1e-8, action_range=np.array([-0.1, 0.1]), **out)
    print('wrote reference_optim.npz', {k: v.shape for k, v in out.items() if hasattr(v, 'shape') and v.ndim > 0})


if __name__ == '__main__':
    main()
