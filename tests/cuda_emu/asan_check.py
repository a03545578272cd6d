"""TEST INFRASTRUCTURE / development aid: the product's kernels under AddressSanitizer on the CUDA execution-model shim — the CPU stand-in
for `compute-sanitizer --tool memcheck` while no GPU is available.  Builds the .cu translation units with -fsanitize=address (ucontext
fibers: ASan understands swapcontext), then runs MPM forward + backward (stored-grid path), the fused g2p2g path and the smoke solver
forward + backward on small scenes with unused slots and walls.  Any out-of-bounds global access aborts with an ASan report.

    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 python tests/cuda_emu/asan_check.py

Last run (end of round 1): clean."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, HERE)
import harness  # noqa: E402


def _asan_library():
    out = os.path.join(HERE, '_build', 'libfluidmpm_emu_asan.so')
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(['/usr/bin/g++', '-std=c++20', '-O1', '-g', '-fsanitize=address', '-fno-omit-frame-pointer', '-DCUEMU_USE_UCONTEXT', '-fPIC', '-shared', '-pthread',
                           '-x', 'c++', '-I', HERE, '-DFMPM_BUILD'] + [os.path.join(harness.CSRC, s) for s in harness.SRCS] + ['-o', out])
    return out


harness.build_library = _asan_library
import numpy as np, torch
harness.enable()
from fluidlab_b200 import MPMSimulator, macros as M
from conftest import make_particles
rng=np.random.RandomState(3); n,N=16,333
x=rng.uniform(0.33,0.67,size=(N,3)).astype(np.float32)
mat=np.array([[M.WATER,M.ELASTIC,M.ICECREAM][i%3] for i in range(N)],dtype=np.int32)
used=(rng.rand(N)>0.1).astype(np.int32)
P=make_particles(x,mat,n,used=used)
for fuse in (False, True):
    s=MPMSimulator(dim=3,quality=n/64,gravity=(0,-10,0),horizon=50,max_substeps_local=20,max_substeps_global=1000,ckpt_dest='cpu',device='cpu')
    s.use_graphs=False; s.fuse_g2p2g=fuse
    s.setup_boundary(type='cube',lower=(0.3,0.3,0.3),upper=(0.7,0.7,0.7)); s.build(None,None,[],P)
    st=s.get_state(); st['v'][:]=rng.randn(N,3)*0.5; st['F'][:]=np.eye(3)+rng.randn(N,3,3)*0.05; s.set_state(0,st)
    if not fuse: s.enable_grad()
    s.step(None); s.step(None)
    if not fuse:
        s.reset_grad(); z9=np.zeros((N,3,3),np.float32); s.set_grad(rng.randn(N,3).astype(np.float32),np.zeros((N,3),np.float32),z9,z9)
        s.step_grad(None); s.step_grad(None)
    print('mpm ok fuse',fuse, flush=True)
# smoke
import types
from fluidlab_b200 import smoke as smoke_mod, meshes
d=np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_smoke.npz'))
z=lambda *s_: torch.zeros(s_,dtype=torch.float32)
air=types.SimpleNamespace(pos=z(41,3),quat=z(41,4),s=z(41),r=z(41),gpos=z(41,3),gquat=z(41,4),gs=z(41),gr=z(41),inject_v=np.asarray(d['inject_v']))
air.quat[:,0]=1
for f,a in zip(d['air_f'],d['air']):
    a=torch.from_numpy(a.astype(np.float32)); air.pos[int(f)]=a[:3]; air.quat[int(f)]=a[3:7]; air.s[int(f)]=a[7]; air.r[int(f)]=a[8]
statics=meshes.Statics()
for vox,T in zip(d['vox'],d['T_static']): statics.add_static(file='x.obj',material=M.PILLAR,has_dynamics=True,sdf=dict(voxels=vox,T_mesh_to_voxels=T))
agent=types.SimpleNamespace(aircon=air)
sim=types.SimpleNamespace(max_steps_local=4,agent=agent,device=torch.device('cpu'),statics=statics,_stream=lambda: None)
sf=smoke_mod.SmokeField(dim=3,ckpt_dest='cpu',res=int(d['res']),dt=float(d['dt']),solver_iters=11,q_dim=1)
sf.lower_y,sf.higher_y=int(d['lower_y']),int(d['higher_y'])
sf.build(sim,agent)
sf.set_state(0,{k:d['st0_'+k] for k in ('v','v_tmp','div','p','q')})
for s_ in range(3): sf.step(s_,10*s_)
sf.reset_grad(); sf._ensure_grad_buffers(); sf._gv[3].normal_(); sf._gq[3].normal_(); sf._gp[3].normal_()
for s_ in (2,1,0): sf.step_grad(s_,10*s_)
print('smoke ok', flush=True)

# agent scenes (injectors, collectors, 6-DOF rigid SDF collider at grid + particle level, static collider, plastic material) and the
# adjoint kernels on a multi-material cloud
import reference_scene_cases as cases
for name in ('latteart', 'jetbot', 'pouring', 'icecream'):
    getattr(cases, f'run_{name}_case')(device='cpu')
    print('scene ok', name, flush=True)
cases.FUSE[0] = True          # the same scenes through the g2p2g path (kAgent kernels, k_p2g_injected)
for name in ('latteart', 'jetbot', 'pouring', 'icecream'):
    getattr(cases, f'run_{name}_case')(device='cpu')
    print('fused scene ok', name, flush=True)
cases.FUSE[0] = False
cases.run_cloud_adjoint_case(device='cpu')
print('cloud adjoint ok', flush=True)
# grad-mode fused forward + backward
s = MPMSimulator(dim=3, quality=n / 64, gravity=(0, -10, 0), horizon=50, max_substeps_local=20, max_substeps_global=1000, ckpt_dest='cpu', device='cpu')
s.use_graphs = False; s.fuse_g2p2g = True
s.setup_boundary(type='cube', lower=(0.3, 0.3, 0.3), upper=(0.7, 0.7, 0.7)); s.build(None, None, [], P)
s.enable_grad(); s.step(None)
s.reset_grad(); z9 = np.zeros((N, 3, 3), np.float32); s.set_grad(rng.randn(N, 3).astype(np.float32), np.zeros((N, 3), np.float32), z9, z9)
s.step_grad(None)
print('grad-mode fused ok', flush=True)
# MAT_RIGID bodies through the fused steps (k_g2p2g gather half, fmpm_advect_rigid, k_p2g_rigid), forward-only and grad mode
dr = np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_run_rigid_bodies.npz'))
Pr = make_particles(dr['x0'], dr['mat'], int(dr['n_grid']))
Pr['body_id'] = dr['body_id']; Pr['bodies'] = {'n': 4}
for grad in (False, True):
    s = MPMSimulator(dim=3, quality=int(dr['n_grid']) / 64, gravity=(0, -10, 0), horizon=50, max_substeps_local=20, max_substeps_global=1000, ckpt_dest='cpu', device='cpu')
    s.use_graphs = False; s.fuse_g2p2g = True
    s.setup_boundary(type='cylinder', xz_radius=float(dr['xz_radius']), xz_center=tuple(dr['xz_center']), y_range=tuple(dr['y_range'])); s.build(None, None, [], Pr)
    if grad: s.enable_grad()
    s.setframe(0, dr['x0'], dr['v0'], dr['C0'], dr['F0'], np.ones(len(dr['x0']), np.int32)); s.sort_frame(0)
    s.step(None)
    if grad:
        Nr = len(dr['x0']); z9 = np.zeros((Nr, 3, 3), np.float32)
        s.reset_grad(); s.set_grad(rng.randn(Nr, 3).astype(np.float32), np.zeros((Nr, 3), np.float32), z9, z9); s.step_grad(None)
    print('rigid fused ok grad', grad, flush=True)
# device Adam on a 251 x 3 table (trainable mask + fix_dim + clip)
import types as _t
from fluidlab_b200 import TrainablePolicy
pol = TrainablePolicy(_t.SimpleNamespace(type='Adam', lr=0.05, beta_1=0.9, beta_2=0.999, epsilon=1e-8), _t.SimpleNamespace(v=(-0.05, 0.05), p=(0.4, 0.6)), 3, 250, (-0.1, 0.1),
                      fix_dim=[1], sim=s)
pol.trainable[:7] = False
for _ in range(3):
    pol.optimize(rng.randn(251, 3).astype(np.float32), {})
print('adam ok', flush=True)
