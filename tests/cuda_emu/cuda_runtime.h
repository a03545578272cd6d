// tests/cuda_emu/cuda_runtime.h — TEST INFRASTRUCTURE ONLY: a tiny model of the CUDA execution model for the HOST compiler.
//
// There is no GPU in the build container.  With this directory first on the include path, `g++ -x c++ file.cu` compiles a .cu translation
// unit of the product unchanged: kernels become ordinary functions and FMPM_LAUNCH (the product's launch macro) runs them with one real
// host thread per CUDA thread, one block at a time — so __syncthreads, __shared__ memory, warp shuffles and atomics behave as on the
// device.  tests/test_smoke_cuda_emu.py uses it to check the kernel bodies AND the host launch logic of fluidlab_b200/csrc/fsmk_smoke.cu
// against the oracle before any GPU time is spent.  Nothing here is reachable from the product: libfluidmpm.so is built by nvcc only.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)
#define FMPM_HOST_EMU 1

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
struct int3 { int x, y, z; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
using std::max;
using std::min;

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "cuda_emu"; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s_, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s_, n); return cudaSuccess; }
struct cudaDeviceProp { int multiProcessorCount; };
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 2; return cudaSuccess; }   // few SMs: small grids, fewer host threads

namespace cuemu {
struct Warp { unsigned long long slot[32]; std::unique_ptr<std::barrier<>> bar; };
struct Block { std::unique_ptr<std::barrier<>> bar; std::vector<Warp> warps; std::vector<unsigned char> smem; };
inline thread_local uint3 t_threadIdx, t_blockIdx;
inline thread_local dim3 t_blockDim, t_gridDim;
inline thread_local Block* t_block = nullptr;
inline thread_local int t_lin = 0;

template <class F> inline void launch(dim3 grid, dim3 block, size_t smem, F&& body) {
  const int nt = (int)(block.x * block.y * block.z);
  if (nt % 32 != 0) abort();   // the shuffle model needs whole warps
  for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
    Block B;
    B.bar = std::make_unique<std::barrier<>>(nt);
    B.warps.resize(nt / 32);
    for (auto& w : B.warps) w.bar = std::make_unique<std::barrier<>>(32);
    B.smem.assign(smem + 16, 0);
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 0; t < nt; t++) {
      th.emplace_back([&, t]() {
        t_threadIdx = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
        t_blockIdx = uint3{bx, by, bz}; t_blockDim = block; t_gridDim = grid; t_block = &B; t_lin = t;
        body();
        // like the hardware, a thread that has exited no longer takes part in the barriers / warp collectives of the others
        B.warps[t / 32].bar->arrive_and_drop();
        B.bar->arrive_and_drop();
      });
    }
    for (auto& x : th) x.join();
  }
}
inline void* dyn_smem() { return t_block->smem.data(); }
}  // namespace cuemu

#define threadIdx (cuemu::t_threadIdx)
#define blockIdx (cuemu::t_blockIdx)
#define blockDim (cuemu::t_blockDim)
#define gridDim (cuemu::t_gridDim)

static inline void __syncthreads() { cuemu::t_block->bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cuemu::t_block->warps[cuemu::t_lin / 32].bar->arrive_and_wait(); }
// every lane of the warp publishes a value, then reads what it needs (lanes that have exited keep their last published value)
template <class T, class F> static inline auto cuemu_collective(T x, F&& read) {
  static_assert(sizeof(T) <= 8, "warp slots are 8 bytes");
  cuemu::Warp& w = cuemu::t_block->warps[cuemu::t_lin / 32];
  const int lane = cuemu::t_lin % 32;
  unsigned long long raw = 0; memcpy(&raw, &x, sizeof(T));
  w.slot[lane] = raw;
  w.bar->arrive_and_wait();
  auto get = [&](int l) { T v; memcpy(&v, &w.slot[l], sizeof(T)); return v; };
  auto r = read(get, lane);
  w.bar->arrive_and_wait();
  return r;
}
template <class T> static inline T __shfl_sync(unsigned, T x, int src) { return cuemu_collective(x, [&](auto get, int) { return get(src & 31); }); }
template <class T> static inline T __shfl_down_sync(unsigned, T x, int off) { return cuemu_collective(x, [&](auto get, int lane) { return lane + off < 32 ? get(lane + off) : x; }); }
template <class T> static inline T __shfl_up_sync(unsigned, T x, int off) { return cuemu_collective(x, [&](auto get, int lane) { return lane - off >= 0 ? get(lane - off) : x; }); }
template <class T> static inline T __shfl_xor_sync(unsigned, T x, int m) { return cuemu_collective(x, [&](auto get, int lane) { return get(lane ^ m); }); }
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  return cuemu_collective((int)(pred != 0), [&](auto get, int) { unsigned b = 0; for (int l = 0; l < 32; l++) if (((mask >> l) & 1u) && get(l)) b |= 1u << l; return b; });
}
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
template <class T> static inline T __reduce_min_sync(unsigned mask, T x) {
  return cuemu_collective(x, [&](auto get, int) { T m = x; for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) m = std::min(m, get(l)); return m; });
}
template <class T> static inline T __reduce_max_sync(unsigned mask, T x) {
  return cuemu_collective(x, [&](auto get, int) { T m = x; for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) m = std::max(m, get(l)); return m; });
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
static inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }
static inline int atomicAdd(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v, std::memory_order_relaxed); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
