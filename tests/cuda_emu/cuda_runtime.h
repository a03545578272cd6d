// tests/cuda_emu/cuda_runtime.h — TEST INFRASTRUCTURE ONLY: a tiny model of the CUDA execution model for the HOST compiler.
//
// There is no GPU in the build container.  With this directory first on the include path, `g++ -x c++ file.cu` compiles a .cu translation
// unit of the product unchanged: kernels become ordinary functions and FMPM_LAUNCH (the product's launch macro) runs them with one
// user-space fiber per CUDA thread, a block per OS worker thread — so __syncthreads, __shared__ memory, warp shuffles and atomics
// behave as on the device.  tests/test_smoke_cuda_emu.py uses it to check the kernel bodies AND the host launch logic of fluidlab_b200/csrc/fsmk_smoke.cu
// against the oracle before any GPU time is spent.  Nothing here is reachable from the product: libfluidmpm.so is built by nvcc only.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static thread_local   /* one block at a time per OS worker thread */
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
// One CUDA thread = one user-space fiber (ucontext); one CUDA block = the fibers of one OS worker thread, scheduled round-robin and
// switched only at barriers / warp collectives; blocks of a grid run concurrently on a few OS workers.  A barrier that can never complete
// (a genuine synchronisation bug in a kernel) is detected and aborts with a message instead of hanging.
enum { RUN = 0, WAIT_BLOCK = 1, WAIT_WARP = 2, DONE = 3 };
#if defined(__x86_64__) && !defined(CUEMU_USE_UCONTEXT)
// user-space context switch without the two rt_sigprocmask system calls swapcontext makes: callee-saved registers + stack pointer
#define CUEMU_FAST_SWITCH 1
__attribute__((naked, noinline)) static void cuemu_switch(void** /*save_sp: rdi*/, void* /*load_sp: rsi*/) {
  asm volatile(
      "pushq %rbp\n\tpushq %rbx\n\tpushq %r12\n\tpushq %r13\n\tpushq %r14\n\tpushq %r15\n\t"
      "movq %rsp, (%rdi)\n\t"
      "movq %rsi, %rsp\n\t"
      "popq %r15\n\tpopq %r14\n\tpopq %r13\n\tpopq %r12\n\tpopq %rbx\n\tpopq %rbp\n\t"
      "ret\n\t");
}
struct Fiber { void* sp; int state; uint3 tid; int lin; };
#else
struct Fiber { ucontext_t ctx; int state; uint3 tid; int lin; };
#endif
struct BlockRun {
  std::vector<Fiber> fib;
#ifdef CUEMU_FAST_SWITCH
  void* sched = nullptr;
#else
  ucontext_t sched;
#endif
  int cur = 0, nt = 0, alive = 0, arrived = 0;
  std::vector<int> warp_alive, warp_arrived;
  std::vector<unsigned long long> slots;   // 32 per warp
  std::vector<unsigned char> smem;
  uint3 bid; dim3 bdim, gdim;
  const std::function<void()>* body = nullptr;
};
inline thread_local BlockRun* t_run = nullptr;
inline Fiber& cur() { return t_run->fib[t_run->cur]; }
#ifdef CUEMU_FAST_SWITCH
inline void yield_to_sched() { BlockRun* r = t_run; cuemu_switch(&r->fib[r->cur].sp, r->sched); }
#else
inline void yield_to_sched() { BlockRun* r = t_run; swapcontext(&r->fib[r->cur].ctx, &r->sched); }
#endif
inline void release(BlockRun* r, int what, int warp) {
  for (auto& f : r->fib) if (f.state == what && (what == WAIT_BLOCK || f.lin / 32 == warp)) f.state = RUN;
}
inline void sync_block() {
  BlockRun* r = t_run;
  if (++r->arrived == r->alive) { r->arrived = 0; release(r, WAIT_BLOCK, 0); return; }
  cur().state = WAIT_BLOCK; yield_to_sched();
}
inline void sync_warp() {
  BlockRun* r = t_run; const int w = cur().lin / 32;
  if (++r->warp_arrived[w] == r->warp_alive[w]) { r->warp_arrived[w] = 0; release(r, WAIT_WARP, w); return; }
  cur().state = WAIT_WARP; yield_to_sched();
}
inline void fiber_main() {
  BlockRun* r = t_run;
  (*r->body)();
  // like the hardware, a thread that has exited no longer takes part in the barriers / warp collectives of the others
  Fiber& f = cur(); const int w = f.lin / 32;
  f.state = DONE;
  r->alive--; r->warp_alive[w]--;
  if (r->alive > 0 && r->arrived == r->alive) { r->arrived = 0; release(r, WAIT_BLOCK, 0); }
  if (r->warp_alive[w] > 0 && r->warp_arrived[w] == r->warp_alive[w]) { r->warp_arrived[w] = 0; release(r, WAIT_WARP, w); }
  yield_to_sched();
}
struct Worker {   // per OS worker thread of one launch; fiber stacks come from a process-wide cache and go back to it
  static constexpr size_t STACK = 512 * 1024;
  std::vector<char*> stacks;
  static std::mutex& mu() { static std::mutex m; return m; }
  static std::vector<char*>& cache() { static std::vector<char*>* c = new std::vector<char*>(); return *c; }
  char* stack(int i) {
    while ((int)stacks.size() <= i) {
      char* p = nullptr;
      { std::lock_guard<std::mutex> lk(mu()); if (!cache().empty()) { p = cache().back(); cache().pop_back(); } }
      if (!p) p = (char*)aligned_alloc(64, STACK);
      stacks.push_back(p);
    }
    return stacks[i];
  }
  ~Worker() { std::lock_guard<std::mutex> lk(mu()); for (char* p : stacks) cache().push_back(p); }
};
inline void run_block(Worker& wk, uint3 bid, dim3 block, dim3 grid, size_t smem, const std::function<void()>& body) {
  const int nt = (int)(block.x * block.y * block.z);
  BlockRun R;
  R.fib.resize(nt); R.nt = R.alive = nt; R.warp_alive.assign(nt / 32, 32); R.warp_arrived.assign(nt / 32, 0); R.slots.assign(nt, 0);
  R.smem.assign(smem + 16, 0); R.bid = bid; R.bdim = block; R.gdim = grid; R.body = &body;
  t_run = &R;
  for (int t = 0; t < nt; t++) {
    Fiber& f = R.fib[t];
    f.state = RUN; f.lin = t;
    f.tid = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
#ifdef CUEMU_FAST_SWITCH
    {  // initial frame: six callee-saved slots, then the entry point `ret` jumps to, then a null return address (fiber_main never returns)
      void** top = (void**)(((uintptr_t)wk.stack(t) + Worker::STACK) & ~(uintptr_t)15);
      *--top = nullptr;
      *--top = (void*)fiber_main;
      for (int i = 0; i < 6; i++) *--top = nullptr;
      f.sp = top;
    }
#else
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = wk.stack(t); f.ctx.uc_stack.ss_size = Worker::STACK; f.ctx.uc_link = &R.sched;
    makecontext(&f.ctx, (void (*)())fiber_main, 0);
#endif
  }
  // Fibers only switch at barriers / warp collectives, so within a segment the threads of a block run one after the other in scheduler
  // order.  CUEMU_SCHED=reverse (or =shuffle) changes that order: a kernel that is missing a __syncthreads / __syncwarp between a shared-memory
  // write and another thread's read gives different results under different orders — the poor man's `compute-sanitizer --tool racecheck`.
  static const int sched_mode = [] { const char* e = getenv("CUEMU_SCHED"); return !e ? 0 : (e[0] == 'r' ? 1 : 2); }();
  std::vector<int> order(nt);
  for (int t = 0; t < nt; t++) order[t] = sched_mode == 1 ? nt - 1 - t : t;
  if (sched_mode == 2) { unsigned sd = 12345u + bid.x * 7919u + bid.y * 104729u; for (int t = nt - 1; t > 0; t--) { sd = sd * 1664525u + 1013904223u; std::swap(order[t], order[(sd >> 8) % (t + 1)]); } }
  for (;;) {
    bool progressed = false, all_done = true;
    for (int oi = 0; oi < nt; oi++) {
      const int t = order[oi];
#ifdef CUEMU_FAST_SWITCH
      if (R.fib[t].state == RUN) { R.cur = t; cuemu_switch(&R.sched, R.fib[t].sp); progressed = true; }
#else
      if (R.fib[t].state == RUN) { R.cur = t; swapcontext(&R.sched, &R.fib[t].ctx); progressed = true; }
#endif
      if (R.fib[t].state != DONE) all_done = false;
    }
    if (all_done) break;
    if (!progressed) { fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): a barrier / warp collective is not reached by every live thread\n", bid.x, bid.y, bid.z); abort(); }
  }
  t_run = nullptr;
}
template <class F> inline void launch(dim3 grid, dim3 block, size_t smem, F&& body_) {
  const int nt = (int)(block.x * block.y * block.z);
  if (nt % 32 != 0) abort();   // the warp model needs whole warps
  const std::function<void()> body = body_;
  const long long nb = (long long)grid.x * grid.y * grid.z;
  const int nw = (int)std::min<long long>(nb, std::max(1u, std::min(8u, std::thread::hardware_concurrency())));
  std::atomic<long long> next{0};
  auto work = [&]() {
    Worker wk;
    for (;;) {
      const long long b = next.fetch_add(1);
      if (b >= nb) break;
      run_block(wk, uint3{(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((long long)grid.x * grid.y))}, block, grid, smem, body);
    }
  };
  if (nw == 1) { work(); return; }
  std::vector<std::thread> th;
  for (int i = 0; i < nw; i++) th.emplace_back(work);
  for (auto& t : th) t.join();
}
inline void* dyn_smem() { return t_run->smem.data(); }
}  // namespace cuemu

#define threadIdx (cuemu::cur().tid)
#define blockIdx (cuemu::t_run->bid)
#define blockDim (cuemu::t_run->bdim)
#define gridDim (cuemu::t_run->gdim)

static inline void __syncthreads() { cuemu::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cuemu::sync_warp(); }
// every lane of the warp publishes a value, then reads what it needs (lanes that have exited keep their last published value)
template <class T, class F> static inline auto cuemu_collective(T x, F&& read) {
  static_assert(sizeof(T) <= 8, "warp slots are 8 bytes");
  cuemu::BlockRun* r = cuemu::t_run;
  const int lin = cuemu::cur().lin, lane = lin % 32;
  unsigned long long* slot = r->slots.data() + (lin / 32) * 32;
  unsigned long long raw = 0; memcpy(&raw, &x, sizeof(T));
  slot[lane] = raw;
  cuemu::sync_warp();
  auto get = [&](int l) { T v; memcpy(&v, &slot[l], sizeof(T)); return v; };
  auto res = read(get, lane);
  cuemu::sync_warp();
  return res;
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
static inline unsigned __reduce_or_sync(unsigned mask, unsigned x) {
  return cuemu_collective(x, [&](auto get, int) { unsigned m = 0; for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) m |= get(l); return m; });
}
template <class T> static inline unsigned __match_any_sync(unsigned mask, T x) {
  return cuemu_collective(x, [&](auto get, int) { unsigned m = 0; for (int l = 0; l < 32; l++) if (((mask >> l) & 1u) && get(l) == x) m |= 1u << l; return m; });
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }

template <class T> static inline void __stcs(T* p, T v) { *p = v; }
static inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }
static inline int atomicAdd(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v, std::memory_order_relaxed); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
