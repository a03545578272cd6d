"""TEST INFRASTRUCTURE ONLY: run the product (fluidlab_b200, unchanged) on a machine without a GPU.

`enable()` (1) builds every .cu translation unit of fluidlab_b200/csrc with g++ against tests/cuda_emu/cuda_runtime.h — a small model of the
CUDA execution model, one host thread per CUDA thread — into tests/cuda_emu/_build/libfluidmpm_emu.so, (2) makes `fluidlab_b200._lib.load()`
return that library instead of the nvcc-built libfluidmpm.so and (3) replaces the handful of torch.cuda calls of the Python host (streams,
events, pinned memory, memory info) by no-ops, so `MPMSimulator(device='cpu')`, `TaichiEnv`, `SmokeField` and `SlabMPMSimulator` drive the
real kernel code on CPU tensors.  Nothing here is reachable from the product; `disable()` undoes every patch."""
import ctypes as C
import hashlib
import os
import subprocess

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EMU_DIR = os.path.join(ROOT, 'tests', 'cuda_emu')
CSRC = os.path.join(ROOT, 'fluidlab_b200', 'csrc')
SRCS = ['fmpm_forward.cu', 'fmpm_backward.cu', 'fmpm_io.cu', 'fmpm_rigid.cu', 'fsmk_smoke.cu']
_state = {}


def build_library():
    # CUEMU_CXXFLAGS: extra compiler flags, e.g. "-mfma -ffp-contract=fast" for a build that contracts a*b+c into FMAs as nvcc does
    # (a second rounding variant of every kernel: parity bars that only hold for one instruction selection show up here, before the GPU)
    extra = os.environ.get('CUEMU_CXXFLAGS', '').split()
    tag = ('_' + hashlib.md5(' '.join(extra).encode()).hexdigest()[:8]) if extra else ''
    out = os.path.join(EMU_DIR, '_build', f'libfluidmpm_emu{tag}.so')
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh'))] + [os.path.join(EMU_DIR, 'cuda_runtime.h'),
            os.path.join(EMU_DIR, 'cub', 'device', 'device_radix_sort.cuh'), os.path.join(ROOT, 'include', 'fluidmpm.h'), os.path.join(ROOT, 'include', 'fluidsmoke.h')]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        tmp = out + f'.{os.getpid()}.tmp'
        subprocess.check_call(['/usr/bin/g++', '-std=c++20', '-O1', '-fPIC', '-shared', '-pthread'] + extra + ['-x', 'c++', '-I', EMU_DIR, '-DFMPM_BUILD'] +
                              [os.path.join(CSRC, s) for s in SRCS] + ['-o', tmp])
        os.replace(tmp, out)
    return out


class _Stream:
    cuda_stream = 0

    def synchronize(self):
        pass


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *a, **k):
        pass

    def synchronize(self):
        pass


def enable():
    """returns the emulated library (ctypes)"""
    if _state:
        return _state['lib']
    from fluidlab_b200 import _lib
    L = C.CDLL(build_library())
    for name, (res, args) in _lib._PROTOS.items():
        fn = getattr(L, name); fn.restype = res; fn.argtypes = args
    _lib.attach_smoke_protos(L)
    saved = dict(load=_lib.load, LIB=_lib._LIB, current_stream=torch.cuda.current_stream, Event=torch.cuda.Event, pin=torch.Tensor.pin_memory, empty=torch.empty,
                 mem=torch.cuda.mem_get_info, cur=torch.cuda.current_device, sync=torch.cuda.synchronize)
    _lib.load = lambda: L
    _lib._LIB = L
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    torch.cuda.Event = _Event
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    real_empty = saved['empty']
    torch.empty = lambda *a, **k: real_empty(*a, **{kk: vv for kk, vv in k.items() if kk != 'pin_memory'})
    torch.cuda.mem_get_info = lambda *a, **k: (1 << 40, 1 << 40)
    torch.cuda.current_device = lambda: 0
    torch.cuda.synchronize = lambda *a, **k: None
    _state.update(lib=L, saved=saved)
    return L


def disable():
    if not _state:
        return
    from fluidlab_b200 import _lib
    s = _state['saved']
    _lib.load, _lib._LIB = s['load'], s['LIB']
    torch.cuda.current_stream, torch.cuda.Event, torch.Tensor.pin_memory, torch.empty = s['current_stream'], s['Event'], s['pin'], s['empty']
    torch.cuda.mem_get_info, torch.cuda.current_device, torch.cuda.synchronize = s['mem'], s['cur'], s['sync']
    _state.clear()
