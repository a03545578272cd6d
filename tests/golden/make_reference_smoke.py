"""Fixtures from RUNS OF THE REAL REFERENCE SMOKE SOLVER (build container only).

The unmodified `fluidlab/fluidengine/simulators/smoke_field.py` executes on the NumPy emulation of the Taichi API
(tests/golden/taichi_emu.py) with a real reference `Static` (synthetic baked SDF) blocking part of the free band and a stand-in air
conditioner exposing exactly the fields the kernels read (`agent.aircon.pos/quat/s/r[f]`, `inject_v`; effectors/aircon.py:18-26).
Two products, both in tests/golden/reference_smoke.npz:

  * float32: initial state + state after 3 steps (SF:95-110) -> tests/test_smoke_oracle.py requires the oracle to reproduce it;
  * float64 (the reference's own `dprecision = 64` switch): central finite differences of sum(w . state after 2 steps) with respect to
    random directions of the initial v / q / p and to every air-conditioner parameter -> pins the oracle's hand-written ADJOINT to the
    reference's own forward code (Taichi's autodiff cannot run here).

The only patches: `lower_y` / `higher_y` (instance attributes, SF:26-27) are lowered so a 20^3 grid has a free band, and the module's bare
`max` / `min` (Taichi built-ins inside @ti.func, SF:304) are bound to the emulated ones.

    python tests/golden/make_reference_smoke.py          # float32 run
    python tests/golden/make_reference_smoke.py fd       # float64 finite differences
    python tests/golden/make_reference_smoke.py stack    # the real MPMSimulator + AgentCirculation + AirCon + SmokeField stack
"""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('FLUIDLAB_REFERENCE', '/root/reference')
sys.path.insert(0, HERE); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))

RES, LOWER_Y, HIGHER_Y, ITERS, DT, QD = 20, 6, 12, 6, 0.5, 1
INJECT_V = (-0.3, 0.0, 1.0)


def load(fp64):
    import taichi_emu
    if fp64:
        taichi_emu.default_fp = np.float64
    ti = taichi_emu.install()
    taichi_emu.stub_optional_dependencies()
    taichi_emu.install_value_semantics(REF)
    sys.path.insert(0, REF)
    macros = importlib.import_module('fluidlab.configs.macros')
    if fp64:
        import torch
        macros.dprecision, macros.DTYPE_TI, macros.DTYPE_NP, macros.DTYPE_TC = 64, ti.f64, np.float64, torch.float64
    import make_reference_run as mrr
    R = dict(macros=macros, meshes=importlib.import_module('fluidlab.fluidengine.meshes'), ti=ti,
             smoke=importlib.import_module('fluidlab.fluidengine.simulators.smoke_field'))
    mrr.patch_mesh_io(R)
    R['smoke'].max, R['smoke'].min = ti.max, ti.min   # Taichi built-ins inside @ti.func
    R['mrr'] = mrr
    return R


def inputs():
    from conftest import sphere_sdf
    rng = np.random.RandomState(401)
    n = RES
    st0 = dict(v=(rng.randn(n, n, n, 3) * 0.8), v_tmp=np.zeros((n, n, n, 3)), div=np.zeros((n, n, n)), p=rng.randn(n, n, n) * 0.1, q=rng.rand(n, n, n, QD))
    unit = lambda q: q / np.linalg.norm(q)
    air = {0: np.array([0.40, 0.50, 0.45, *unit(np.array([0.9, 0.1, -0.3, 0.2])), 3.0, 2.5]),
           10: np.array([0.45, 0.47, 0.50, *unit(np.array([0.8, -0.2, 0.3, 0.1])), 2.0, 3.0]),
           20: np.array([0.55, 0.45, 0.52, *unit(np.array([0.7, 0.3, 0.1, -0.4])), 4.0, 1.5])}
    from conftest import box_sdf
    vox, T = sphere_sdf(0.12, 0.25, res=24)
    T = T.copy(); T[:3, 3] -= T[:3, :3] @ np.array([0.62, 0.48, 0.40])   # the volume's mesh frame sits at this world position
    # a "room" like envs/circulation_env.py:47-56: everything outside an inner box is solid, so no free cell touches the x / z domain
    # edge (there the reference's compute_location would index out of bounds, see oracle/smoke_oracle.hpp)
    rvox, rT = box_sdf(np.array([0.36, 0.60, 0.36]), 0.55, res=24)
    rvox = -rvox
    rT = rT.copy(); rT[:3, 3] -= rT[:3, :3] @ np.array([0.5, 0.5, 0.5])
    vox, T = np.stack([vox.reshape(24, 24, 24), rvox.reshape(24, 24, 24)]), np.stack([T, rT])
    w = dict(v=rng.randn(n, n, n, 3), q=rng.randn(n, n, n, QD), p=rng.randn(n, n, n))
    dirs = {k: rng.randn(*st0[k].shape) for k in ('v', 'q', 'p')}
    return st0, air, vox, T, w, dirs


def build(R, air, vox, T):
    M = R['macros']; ti = R['ti']
    statics = R['meshes'].Statics()
    for i, name in enumerate(('blob.obj', 'room.obj')):
        R['mrr'].SDF_REGISTRY[name] = dict(voxels=vox[i], T_mesh_to_voxels=T[i])
        statics.add_static(file=name, material=M.PILLAR, has_dynamics=True, pos=(0.0, 0.0, 0.0))
    T_used = np.stack([np.asarray(st.T_mesh_to_voxels_np, dtype=np.float64) for st in statics])
    S = R['smoke'].SmokeField(dim=3, ckpt_dest='cpu', res=RES, dt=DT, solver_iters=ITERS, q_dim=QD)
    S.lower_y, S.higher_y = LOWER_Y, HIGHER_Y
    Tsub = 40
    aircon = types.SimpleNamespace(pos=ti.Vector.field(3, M.DTYPE_TI, shape=(Tsub + 1,)), quat=ti.Vector.field(4, M.DTYPE_TI, shape=(Tsub + 1,)),
                                   s=ti.field(M.DTYPE_TI, shape=(Tsub + 1,)), r=ti.field(M.DTYPE_TI, shape=(Tsub + 1,)), inject_v=ti.Vector(list(INJECT_V)))
    for f, a in air.items():
        aircon.pos[f] = a[0:3]; aircon.quat[f] = a[3:7]; aircon.s[f] = a[7]; aircon.r[f] = a[8]
    agent = types.SimpleNamespace(aircon=aircon)
    sim = types.SimpleNamespace(max_steps_local=4, n_grid=16, agent=agent, n_statics=len(statics), statics=statics)
    S.build(sim, agent)
    return S, T_used


def run(R, st0, air, vox, T, n_steps):
    S, T_used = build(R, air, vox, T)
    q_init = S.get_state(0)['q'].copy()     # init_fields, SF:87-93
    dt_np = R['macros'].DTYPE_NP
    S.set_state(0, {k: a.astype(dt_np) for k, a in st0.items()})
    for s in range(n_steps):
        S.step(s, 10 * s)
    return S, T_used, q_init


def circulation_stack(R):
    """THE REAL STACK of envs/circulation_env.py, reduced: the reference's MPMSimulator (10 parked particles) + AgentCirculation + AirCon
    (8-component action through set_action_kernel / set_velocity / move_kernel) + SmokeField, stepped with MPMSimulator.step(action)
    (mpm_simulator.py:734-753: set_action, smoke step at step level, 10 substeps with agent.move) -> reference_circulation.npz"""
    M, ti = R['macros'], R['ti']
    agents = importlib.import_module('fluidlab.fluidengine.agents')
    sim_mod = importlib.import_module('fluidlab.fluidengine.simulators.mpm_simulator')
    T = 40
    agent = agents.AgentCirculation(max_substeps_local=T, max_substeps_global=1000, max_action_steps_global=20, ckpt_dest='cpu')
    params = dict(init_pos=(0.8, 0.8, 0.5), action_dim=8, action_scale_p=(1.0,) * 8, action_scale_v=(1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1000.0, 50.0), inject_v=INJECT_V)
    ebnd = dict(type='cube', lower=(0.05, 0.05, 0.05), upper=(0.95, 0.95, 0.95))
    agent.add_effector(type='AirCon', params=params, mesh_cfg=None, boundary_cfg=ebnd)
    S = sim_mod.MPMSimulator(dim=3, quality=0.25, gravity=(0.0, -20.0, 0.0), horizon=10, max_substeps_local=T, max_substeps_global=1000, ckpt_dest='cpu')
    smoke = R['smoke'].SmokeField(dim=3, ckpt_dest='cpu', res=RES, dt=DT, solver_iters=ITERS, q_dim=QD)
    smoke.lower_y, smoke.higher_y = LOWER_Y, HIGHER_Y
    N = 10
    x = np.tile(np.array([-100.0, -100.0, -100.0], dtype=np.float32), (N, 1))
    statics = R['meshes'].Statics()
    _, _, vox, Tm, _, _ = inputs()     # the "room" volume (second entry): no free cell touches the x / z domain edge
    R['mrr'].SDF_REGISTRY['room.obj'] = dict(voxels=vox[1], T_mesh_to_voxels=Tm[1])
    statics.add_static(file='room.obj', material=M.PILLAR, has_dynamics=True, pos=(0.0, 0.0, 0.0))
    S.build(agent, smoke, statics, dict(x=x, used=np.zeros(N), mat=np.full(N, M.WATER), rho=np.ones(N, np.float32), body_id=np.zeros(N), bodies={'n': 1}))
    agent.build(S)
    smoke.build(S, agent)
    action_p = np.array([0.55, 0.5, 0.45, 0.0, 0.0, 0.0, 0.0, 0.0], dtype=np.float32)
    actions = np.array([[0.010, 0.002, 0.005, 0.00, 0.10, 0.00, 0.020, 0.040], [0.005, -0.004, 0.010, 0.05, 0.05, -0.02, 0.030, 0.050],
                        [-0.008, 0.003, 0.002, -0.03, 0.08, 0.04, 0.015, 0.030]], dtype=np.float32)
    agent.apply_action_p(action_p)
    for a in actions:
        S.step(a)
    air = agent.aircon
    nf = 10 * len(actions)
    out = dict(res=RES, lower_y=LOWER_Y, higher_y=HIGHER_Y, iters=ITERS, dt=DT, q_dim=QD, inject_v=np.array(INJECT_V), T=T, action_p=action_p, actions=actions,
               scale_v=np.array(params['action_scale_v']), e_lower=np.array(ebnd['lower']), e_upper=np.array(ebnd['upper']), init_pos=np.array(params['init_pos']),
               ref_pos=np.stack([np.asarray(air.pos[f]) for f in range(nf + 1)]), ref_quat=np.stack([np.asarray(air.quat[f]) for f in range(nf + 1)]),
               ref_s=np.array([float(air.s[f]) for f in range(nf)]), ref_r=np.array([float(air.r[f]) for f in range(nf)]),
               ref_state=np.asarray(agent.get_state(nf)[0], dtype=np.float64), room_vox=vox[1].astype(np.float32),
               room_T=np.asarray(statics[0].T_mesh_to_voxels_np, dtype=np.float64))
    for s_ in (1, 2, 3):
        st = smoke.get_state(s_)
        for k in ('v', 'p', 'q'):
            out[f'ref{s_}_{k}'] = st[k].astype(np.float32)
    path = os.path.join(HERE, 'reference_circulation.npz')
    np.savez_compressed(path, **out)
    print('reference_circulation.npz', os.path.getsize(path), 'bytes; |v| max', float(np.abs(out['ref3_v']).max()))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'stack':
        return circulation_stack(load(False))
    fp64 = len(sys.argv) > 1 and sys.argv[1] == 'fd'
    R = load(fp64)
    st0, air, vox, T, w, dirs = inputs()
    out_path = os.path.join(HERE, 'reference_smoke_fd.npz' if fp64 else 'reference_smoke.npz')
    if not fp64:
        S, T_used, q_init = run(R, st0, air, vox, T, 3)
        out = dict(res=RES, lower_y=LOWER_Y, higher_y=HIGHER_Y, iters=ITERS, dt=DT, q_dim=QD, inject_v=np.array(INJECT_V), vox=vox.astype(np.float32), T_static=T_used,
                   air_f=np.array(sorted(air)), air=np.stack([air[f] for f in sorted(air)]), q_init=q_init.astype(np.float32),
                   free0=np.asarray(S.grid_ng.is_free.to_numpy()[0], dtype=np.int8))
        for k, a in st0.items():
            out['st0_' + k] = a.astype(np.float32)
        for s in (1, 2, 3):
            st = S.get_state(s)
            for k in ('v', 'p', 'q'):
                out[f'ref{s}_{k}'] = st[k].astype(np.float32)
            if s < 3:
                pass
        for s in (0, 1, 2):
            st = S.get_state(s)
            out[f'ref{s}_v_tmp'] = st['v_tmp'].astype(np.float32); out[f'ref{s}_div'] = st['div'].astype(np.float32)
        np.savez_compressed(out_path, **out)
        print('reference_smoke.npz', os.path.getsize(out_path), 'bytes; free cells', int(out['free0'].sum()))
        return

    def loss(st, air_):
        S, _, _ = run(R, st, air_, vox, T, 2)
        r = S.get_state(2)
        return float((r['v'] * w['v']).sum() + (r['q'] * w['q']).sum() + (r['p'] * w['p']).sum())
    fd = {}
    eps = 1e-6
    for k, d in dirs.items():
        sp = {a: b.copy() for a, b in st0.items()}; sm = {a: b.copy() for a, b in st0.items()}
        sp[k] += eps * d; sm[k] -= eps * d
        fd['fd_' + k] = (loss(sp, air) - loss(sm, air)) / (2 * eps)
        print(k, fd['fd_' + k])
    fd_air = np.zeros((2, 9))
    for a_i, f in enumerate((0, 10)):
        for idx in range(9):
            ap = {a: b.copy() for a, b in air.items()}; am = {a: b.copy() for a, b in air.items()}
            ap[f][idx] += eps; am[f][idx] -= eps
            fd_air[a_i, idx] = (loss(st0, ap) - loss(st0, am)) / (2 * eps)
        print('air', f, fd_air[a_i])
    np.savez_compressed(out_path, fd_air=fd_air, **fd, **{'w_' + k: a for k, a in w.items()}, **{'dir_' + k: a for k, a in dirs.items()})
    print('reference_smoke_fd.npz', os.path.getsize(out_path), 'bytes')


if __name__ == '__main__':
    main()
