// fsmk_smoke.cu — the Eulerian smoke solver of fluidlab/fluidengine/simulators/smoke_field.py (abbrev. SF) for sm_100a:
// forward step (SF:95-110) and hand-written adjoint (SF:112-127), behind the C ABI of include/fluidsmoke.h.
//
// What bounds it.  The free region is a thin band (lower_y < j < higher_y: 7 layers of a 128^3 grid, 115 k cells, SF:26-27,192-194), so
// a step is dominated by the `solver_iters` (50..500) Jacobi sweeps over that band: launch latency and L2 round trips, not HBM.
// k_jacobi_tile therefore blocks the sweeps in TIME: a CTA stages a (16+2*8)^2 x H tile of the band (p, div, mask) in shared memory,
// runs up to 8 sweeps there (the region that is still exact shrinks by one cell per sweep from the tile edge; the inner 16 x 16 columns
// stay exact) and writes the inner columns back: 500 sweeps = 63 launches instead of 500, each cell update reading shared memory only.
// The arithmetic per cell is the reference's, in the reference's order, so the result equals sweep-by-sweep Jacobi bit for bit.
// The dense passes (advection writes v_tmp / q of every cell, SF:229-233) are plain coalesced HBM streams (84 MB per step at 128^3).
//
// Adjoint: every reference scatter (`.grad` of a gather through compute_location) is re-expressed as a gather over the inverse
// relation, so projection / Jacobi / divergence adjoints are deterministic and atomic-free; the Jacobi operator with its mirror
// boundary is symmetric, so the adjoint sweeps reuse k_jacobi_tile (no div term, plus the running sum that feeds div's adjoint).
// Only the semi-Lagrangian advection adjoint scatters (vector atomics into the 8 corners of 5 trilinear lookups per free cell).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include "../../include/fluidsmoke.h"
#include "fmpm_sdf.cuh"

struct FsmkHandle {
  FsmkConfig cfg;
  FsmkBuffers buf;
  int n_statics; SdfDev statics[4];
  FsmkAircon air;
  bool bound;
  char err[512];
};

struct SParams {
  int n, G, S, qd, lo_y, hi_y, j0, H;   // band rows j0 .. j0+H-1
  float dt, low_T, dx; float inj[3];
  float4* v; float4* vt; float* dv; float* p; float* q; unsigned char* fr;
  float4* gv; float4* gvt; float* gdv; float* gp; float* gq;
  int n_statics; SdfDev statics[4];
  const float *apos, *aquat, *as, *ar; float *gapos, *gaquat, *gas, *gar;
};

static SParams make_sparams(const FsmkHandle* h) {
  SParams P;
  const FsmkConfig& c = h->cfg;
  P.n = c.res; P.G = c.res * c.res * c.res; P.S = c.max_steps_local; P.qd = c.q_dim; P.lo_y = c.lower_y; P.hi_y = c.higher_y;
  P.j0 = c.lower_y + 1 < 0 ? 0 : c.lower_y + 1;
  const int j1 = c.higher_y - 1 > c.res - 1 ? c.res - 1 : c.higher_y - 1;
  P.H = j1 - P.j0 + 1 < 0 ? 0 : j1 - P.j0 + 1;
  P.dt = c.dt; P.low_T = c.low_T; P.dx = 1.0f / (float)c.res;
  for (int d = 0; d < 3; d++) P.inj[d] = c.inject_v[d];
  P.v = (float4*)h->buf.v; P.vt = (float4*)h->buf.v_tmp; P.dv = (float*)h->buf.div; P.p = (float*)h->buf.p; P.q = (float*)h->buf.q;
  P.fr = (unsigned char*)h->buf.is_free;
  P.gv = (float4*)h->buf.gv; P.gvt = (float4*)h->buf.gv_tmp; P.gdv = (float*)h->buf.gdiv; P.gp = (float*)h->buf.gp; P.gq = (float*)h->buf.gq;
  P.n_statics = h->n_statics;
  for (int i = 0; i < 4; i++) P.statics[i] = h->statics[i];
  P.apos = (const float*)h->air.pos; P.aquat = (const float*)h->air.quat; P.as = (const float*)h->air.s; P.ar = (const float*)h->air.r;
  P.gapos = (float*)h->air.gpos; P.gaquat = (float*)h->air.gquat; P.gas = (float*)h->air.gs; P.gar = (float*)h->air.gr;
  return P;
}

#define FSMK_CHECK_LAUNCH(h, name)                                                   \
  do {                                                                               \
    cudaError_t e_ = cudaGetLastError();                                             \
    if (e_ != cudaSuccess) { snprintf((h)->err, sizeof((h)->err), "%s: %s", name, cudaGetErrorString(e_)); return 1; } \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------- index helpers
__device__ __forceinline__ int cidx(const SParams& P, int i, int j, int k) { return (i * P.n + j) * P.n + k; }
__device__ __forceinline__ bool in_range(const SParams& P, int i, int j, int k) { return (unsigned)i < (unsigned)P.n && (unsigned)j < (unsigned)P.n && (unsigned)k < (unsigned)P.n; }
// is_free(), SF:312-323
__device__ __forceinline__ bool free_at(const SParams& P, const unsigned char* fr, int i, int j, int k) { return in_range(P, i, j, k) && fr[cidx(P, i, j, k)]; }
// compute_location, SF:301-310: the cell sampled for neighbour (u+du, v+dv, w+dw): clamped into the grid; the centre (u, v, w) when that
// cell is not free.  (When (u, v, w) itself is outside the grid the reference would index out of bounds; the clamped cell is used.)
__device__ __forceinline__ int loc(const SParams& P, const unsigned char* fr, int u, int v, int w, int du, int dv, int dw) {
  const int i = min(max(u + du, 0), P.n - 1), j = min(max(v + dv, 0), P.n - 1), k = min(max(w + dw, 0), P.n - 1);
  const int c = cidx(P, i, j, k);
  if (!fr[c] && in_range(P, u, v, w)) return cidx(P, u, v, w);
  return c;
}
// band thread -> cell; false when out of the band's extent
__device__ __forceinline__ bool band_cell(const SParams& P, int& i, int& j, int& k) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.n * P.H * P.n) return false;
  k = t % P.n; const int r = t / P.n; j = P.j0 + r % P.H; i = r / P.H;
  return true;
}

// ---------------------------------------------------------------------------------------------------------------- SF:190-201
__global__ void __launch_bounds__(256) k_free_space(const SParams P, const int s) {
  int i, j, k;
  if (!band_cell(P, i, j, k)) return;
  bool fr = true;
  const float pw[3] = {((float)i + 0.5f) * P.dx, ((float)j + 0.5f) * P.dx, ((float)k + 0.5f) * P.dx};
  for (int m = 0; m < P.n_statics; m++) {   // Static.is_collide, meshes/static.py:106-114
    const SdfDev& M = P.statics[m];
    float pv[3];
#pragma unroll
    for (int r = 0; r < 3; r++) pv[r] = M.T[r * 4] * pw[0] + M.T[r * 4 + 1] * pw[1] + M.T[r * 4 + 2] * pw[2] + M.T[r * 4 + 3];
    if (sdf_lookup<false>(M, pv, nullptr) <= 0.f) fr = false;
  }
  P.fr[(size_t)s * P.G + cidx(P, i, j, k)] = fr ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------- trilerp, SF:325-347
struct Tri { int idx[8]; float w[8]; float W; float wx[3][2], sg[3][2]; };
template <bool kAdj>
__device__ __forceinline__ void tri_setup(const SParams& P, const unsigned char* fr, const float* pp, Tri& t) {
  int base[3]; float pI[3];
#pragma unroll
  for (int d = 0; d < 3; d++) { pI[d] = pp[d] - 0.5f; base[d] = (int)floorf(pI[d]); }
#pragma unroll
  for (int d = 0; d < 3; d++)
#pragma unroll
    for (int o = 0; o < 2; o++) {
      const float tt = pI[d] - (float)(base[d] + o);
      t.wx[d][o] = 1.f - fabsf(tt);
      if (kAdj) t.sg[d][o] = tt > 0.f ? 1.f : (tt < 0.f ? -1.f : 0.f);
    }
  float W = 0.f;
#pragma unroll
  for (int m = 0; m < 8; m++) {
    const int oi = m >> 2, oj = (m >> 1) & 1, ok = m & 1;
    t.w[m] = t.wx[0][oi] * t.wx[1][oj] * t.wx[2][ok];
    t.idx[m] = loc(P, fr, base[0] + oi, base[1] + oj, base[2] + ok, 0, 0, 0);
    W += t.w[m];
  }
  t.W = W;
}
__device__ __forceinline__ void trilerp_v(const float4* fld, const Tri& t, float* out) {
  out[0] = out[1] = out[2] = 0.f;
#pragma unroll
  for (int m = 0; m < 8; m++) { const float4 a = fld[t.idx[m]]; out[0] += t.w[m] * a.x; out[1] += t.w[m] * a.y; out[2] += t.w[m] * a.z; }
  out[0] /= t.W; out[1] /= t.W; out[2] /= t.W;
}
__device__ __forceinline__ float trilerp_s(const float* fld, const Tri& t) {
  float o = 0.f;
#pragma unroll
  for (int m = 0; m < 8; m++) o += t.w[m] * fld[t.idx[m]];
  return o / t.W;
}
// backtrace (RK3), SF:349-360; fld = v[s]
__device__ __forceinline__ void backtrace(const SParams& P, const unsigned char* fr, const float4* vs, const float* p0, float* pe, float* v1, float* v2, float* v3,
                                          float* p1, float* p2) {
  Tri t;
  tri_setup<false>(P, fr, p0, t); trilerp_v(vs, t, v1);
#pragma unroll
  for (int d = 0; d < 3; d++) p1[d] = p0[d] - 0.5f * P.dt * v1[d];
  tri_setup<false>(P, fr, p1, t); trilerp_v(vs, t, v2);
#pragma unroll
  for (int d = 0; d < 3; d++) p2[d] = p0[d] - 0.75f * P.dt * v2[d];
  tri_setup<false>(P, fr, p2, t); trilerp_v(vs, t, v3);
#pragma unroll
  for (int d = 0; d < 3; d++) pe[d] = p0[d] - P.dt * ((float)(2.0 / 9.0) * v1[d] + (float)(1.0 / 3.0) * v2[d] + (float)(4.0 / 9.0) * v3[d]);
}
// the air conditioner's impulse at cell (i, j, k), SF:216-221
struct Impulse { float dir[3], d[3], dist, factor; };
__device__ __forceinline__ void impulse(const SParams& P, int f, int i, int j, int k, Impulse& I) {
  q_rot(P.aquat + (size_t)f * 4, P.inj, I.dir);
  const float c[3] = {(float)i, (float)j, (float)k};
  float ss = 0.f;
#pragma unroll
  for (int d = 0; d < 3; d++) { I.d[d] = c[d] - P.apos[(size_t)f * 3 + d] / P.dx; ss += I.d[d] * I.d[d]; }
  I.dist = sqrtf(ss + 1e-12f);           // norm(EPS), configs/macros.py:213
  I.factor = expf(-I.dist / P.ar[f]);
}

// ---------------------------------------------------------------------------------------------------------------- SF:203-233 (+ SF:288-289)
__global__ void __launch_bounds__(256) k_advect(const SParams P, const int s, const int f) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P.G) return;
  const size_t o0 = (size_t)s * P.G, o1 = o0 + P.G;
  const unsigned char* fr = P.fr + o0;
  if (!fr[g]) {
    P.vt[o0 + g] = make_float4(0.f, 0.f, 0.f, 0.f);
    P.v[o1 + g] = make_float4(0.f, 0.f, 0.f, 0.f);                 // subtract_gradient's else branch: v[s+1] = v_tmp[s]
    for (int a = 0; a < P.qd; a++) P.q[((size_t)(s + 1) * P.qd + a) * P.G + g] = P.q[((size_t)s * P.qd + a) * P.G + g];
    return;
  }
  const int k = g % P.n, j = (g / P.n) % P.n, i = g / (P.n * P.n);
  const float p0[3] = {(float)i + 0.5f, (float)j + 0.5f, (float)k + 0.5f};
  float pe[3], v1[3], v2[3], v3[3], p1[3], p2[3];
  const float4* vs = P.v + o0;
  backtrace(P, fr, vs, p0, pe, v1, v2, v3, p1, p2);
  Tri t; tri_setup<false>(P, fr, pe, t);
  float vf[3]; trilerp_v(vs, t, vf);
  Impulse I; impulse(P, f, i, j, k, I);
  const float sf = P.as[f];
  P.vt[o0 + g] = make_float4(vf[0] * 1.f + (I.dir[0] * sf * I.factor) * P.dt + 0.f, vf[1] * 1.f + (I.dir[1] * sf * I.factor) * P.dt + 0.f,
                             vf[2] * 1.f + (I.dir[2] * sf * I.factor) * P.dt + 0.f, 0.f);
  for (int a = 0; a < P.qd; a++) {
    const float qf = trilerp_s(P.q + ((size_t)s * P.qd + a) * P.G, t) * 1.f;
    P.q[((size_t)(s + 1) * P.qd + a) * P.G + g] = (1.f - I.factor) * qf + I.factor * P.low_T;
  }
}

// ---------------------------------------------------------------------------------------------------------------- SF:235-261
__global__ void __launch_bounds__(256) k_divergence(const SParams P, const int s) {
  int i, j, k;
  if (!band_cell(P, i, j, k)) return;
  const size_t o0 = (size_t)s * P.G;
  const unsigned char* fr = P.fr + o0;
  const int g = cidx(P, i, j, k);
  if (!fr[g]) return;
  const float4* vt = P.vt + o0;
  const float4 vc = vt[g];
  const float lx = free_at(P, fr, i - 1, j, k) ? vt[g - P.n * P.n].x : -vc.x, hx = free_at(P, fr, i + 1, j, k) ? vt[g + P.n * P.n].x : -vc.x;
  const float ly = free_at(P, fr, i, j - 1, k) ? vt[g - P.n].y : -vc.y, hy = free_at(P, fr, i, j + 1, k) ? vt[g + P.n].y : -vc.y;
  const float lz = free_at(P, fr, i, j, k - 1) ? vt[g - 1].z : -vc.z, hz = free_at(P, fr, i, j, k + 1) ? vt[g + 1].z : -vc.z;
  P.dv[o0 + g] = (hx - lx + hy - ly + hz - lz) * 0.5f;
}

// ---------------------------------------------------------------------------------------------------------------- pressure, SF:97-106,135-146
// Temporally blocked Jacobi.  Tile: (JT_OUT + 2*JT_HALO)^2 columns x H <= JT_HMAX layers; ns <= JT_HALO sweeps per launch.
//   kGrad = false: new = (pl + pr + pb + pt + pp + pq - div) / 6          (SF:146)
//   kGrad = true : acc += cur; new = (pl + pr + pb + pt + pp + pq) / 6    (the operator is symmetric: adjoint sweep = sweep without div;
//                  acc = sum over sweeps of the sweep's OUTPUT adjoint, which is what div's adjoint needs: gdiv -= acc / 6)
#define JT_OUT 16
#define JT_HALO 8
#define JT_L (JT_OUT + 2 * JT_HALO)
#define JT_HMAX 8
template <bool kGrad>
__global__ void __launch_bounds__(256) k_jacobi_tile(const SParams P, const int s, const float* __restrict__ pin, float* __restrict__ pout, float* __restrict__ acc_out,
                                                      const int ns, const int first) {
  FMPM_DYN_SMEM(float, sm);
  const int H = P.H, cells = H * JT_L * JT_L;
  float* a0 = sm; float* a1 = a0 + cells; float* aux = a1 + cells;            // aux: div (forward) or acc (adjoint)
  unsigned char* mk = (unsigned char*)(aux + cells);
  const int ti0 = blockIdx.x * JT_OUT - JT_HALO, tk0 = blockIdx.y * JT_OUT - JT_HALO;
  const unsigned char* fr = P.fr + (size_t)s * P.G;
  for (int c = threadIdx.x; c < cells; c += blockDim.x) {
    const int k = c % JT_L, i = (c / JT_L) % JT_L, jj = c / (JT_L * JT_L);
    const int gi = ti0 + i, gk = tk0 + k, gj = P.j0 + jj;
    bool m = (unsigned)gi < (unsigned)P.n && (unsigned)gk < (unsigned)P.n;
    int g = 0;
    if (m) { g = cidx(P, gi, gj, gk); m = fr[g] != 0; }
    mk[c] = m ? 1 : 0;
    a0[c] = m ? pin[g] : 0.f;
    aux[c] = kGrad ? 0.f : (m ? P.dv[(size_t)s * P.G + g] : 0.f);
  }
  __syncthreads();
  float* cur = a0; float* nxt = a1;
  for (int it = 0; it < ns; it++) {
    for (int c = threadIdx.x; c < cells; c += blockDim.x) {
      if (!mk[c]) continue;
      const int k = c % JT_L, i = (c / JT_L) % JT_L, jj = c / (JT_L * JT_L);
      const float pc = cur[c];
      const float pl = (i > 0 && mk[c - JT_L]) ? cur[c - JT_L] : pc, pr = (i < JT_L - 1 && mk[c + JT_L]) ? cur[c + JT_L] : pc;
      const float pb = (jj > 0 && mk[c - JT_L * JT_L]) ? cur[c - JT_L * JT_L] : pc, pt = (jj < H - 1 && mk[c + JT_L * JT_L]) ? cur[c + JT_L * JT_L] : pc;
      const float pp = (k > 0 && mk[c - 1]) ? cur[c - 1] : pc, pq = (k < JT_L - 1 && mk[c + 1]) ? cur[c + 1] : pc;
      if (kGrad) { aux[c] += pc; nxt[c] = (pl + pr + pb + pt + pp + pq) / 6.0f; }
      else nxt[c] = (pl + pr + pb + pt + pp + pq - aux[c]) / 6.0f;
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
  for (int c = threadIdx.x; c < H * JT_OUT * JT_OUT; c += blockDim.x) {
    const int k = c % JT_OUT + JT_HALO, i = (c / JT_OUT) % JT_OUT + JT_HALO, jj = c / (JT_OUT * JT_OUT);
    const int l = (jj * JT_L + i) * JT_L + k;
    if (!mk[l]) continue;
    const int g = cidx(P, ti0 + i, P.j0 + jj, tk0 + k);
    pout[g] = cur[l];
    if (kGrad) acc_out[g] = first ? aux[l] : acc_out[g] + aux[l];
  }
}
// one sweep per launch, any band height (fallback when H > JT_HMAX)
template <bool kGrad>
__global__ void __launch_bounds__(256) k_jacobi_simple(const SParams P, const int s, const float* __restrict__ pin, float* __restrict__ pout, float* __restrict__ acc_out,
                                                        const int sweep, const int first) {
  int i, j, k;
  if (!band_cell(P, i, j, k)) return;
  const unsigned char* fr = P.fr + (size_t)s * P.G;
  const int g = cidx(P, i, j, k);
  if (!fr[g]) return;
  if (!sweep) { pout[g] = pin[g]; if (kGrad) acc_out[g] = first ? 0.f : acc_out[g]; return; }
  const float pl = pin[loc(P, fr, i, j, k, -1, 0, 0)], pr = pin[loc(P, fr, i, j, k, 1, 0, 0)], pb = pin[loc(P, fr, i, j, k, 0, -1, 0)],
              pt = pin[loc(P, fr, i, j, k, 0, 1, 0)], pp = pin[loc(P, fr, i, j, k, 0, 0, -1)], pq = pin[loc(P, fr, i, j, k, 0, 0, 1)];
  if (kGrad) { acc_out[g] = (first ? 0.f : acc_out[g]) + pin[g]; pout[g] = (pl + pr + pb + pt + pp + pq) / 6.0f; }
  else pout[g] = (pl + pr + pb + pt + pp + pq - P.dv[(size_t)s * P.G + g]) / 6.0f;
}
// end of the adjoint solve: pressure_to_swap.grad (SF:124) and the div adjoint collected over the sweeps (SF:151)
__global__ void __launch_bounds__(256) k_pressure_grad_finish(const SParams P, const int s, const float* __restrict__ gres, const float* __restrict__ acc) {
  int i, j, k;
  if (!band_cell(P, i, j, k)) return;
  const size_t o0 = (size_t)s * P.G;
  const int g = cidx(P, i, j, k);
  if (!P.fr[o0 + g]) return;
  P.gp[o0 + g] += gres[g];
  P.gdv[o0 + g] -= acc[g] / 6.0f;
}

// ---------------------------------------------------------------------------------------------------------------- SF:275-287
__global__ void __launch_bounds__(256) k_project(const SParams P, const int s) {
  int i, j, k;
  if (!band_cell(P, i, j, k)) return;
  const size_t o0 = (size_t)s * P.G, o1 = o0 + P.G;
  const unsigned char* fr = P.fr + o0;
  const int g = cidx(P, i, j, k);
  if (!fr[g]) return;      // v[s+1] of non-free cells was written by k_advect
  const float* p1 = P.p + o1;
  const float pl = p1[loc(P, fr, i, j, k, -1, 0, 0)], pr = p1[loc(P, fr, i, j, k, 1, 0, 0)], pb = p1[loc(P, fr, i, j, k, 0, -1, 0)],
              pt = p1[loc(P, fr, i, j, k, 0, 1, 0)], pp = p1[loc(P, fr, i, j, k, 0, 0, -1)], pq = p1[loc(P, fr, i, j, k, 0, 0, 1)];
  const float4 vt = P.vt[o0 + g];
  P.v[o1 + g] = make_float4(vt.x - 0.5f * (pr - pl), vt.y - 0.5f * (pt - pb), vt.z - 0.5f * (pq - pp), 0.f);
}

// ================================================================================================================ adjoint kernels
// subtract_gradient.grad (SF:115) as a gather: p[s+1] at x is read by x - e as its "+" sample and by x + e as its "-" sample (when
// those are free), and by x itself in place of a blocked neighbour.
__global__ void __launch_bounds__(256) k_project_grad(const SParams P, const int s) {
  int i, j, k;
  if (!band_cell(P, i, j, k)) return;
  const size_t o0 = (size_t)s * P.G, o1 = o0 + P.G;
  const unsigned char* fr = P.fr + o0;
  const int g = cidx(P, i, j, k);
  if (!fr[g]) return;      // v_tmp of a non-free cell is the constant 0: its adjoint is a dead end
  const float4* gv1 = P.gv + o1;
  const float4 gc = gv1[g];
  float4 t = P.gvt[o0 + g]; t.x += gc.x; t.y += gc.y; t.z += gc.z; P.gvt[o0 + g] = t;
  const int n2 = P.n * P.n;
  float acc = 0.f;
  { const bool fm = free_at(P, fr, i - 1, j, k), fp = free_at(P, fr, i + 1, j, k);
    acc += -0.5f * ((fm ? gv1[g - n2].x : 0.f) + (fp ? 0.f : gc.x)) + 0.5f * ((fp ? gv1[g + n2].x : 0.f) + (fm ? 0.f : gc.x)); }
  { const bool fm = free_at(P, fr, i, j - 1, k), fp = free_at(P, fr, i, j + 1, k);
    acc += -0.5f * ((fm ? gv1[g - P.n].y : 0.f) + (fp ? 0.f : gc.y)) + 0.5f * ((fp ? gv1[g + P.n].y : 0.f) + (fm ? 0.f : gc.y)); }
  { const bool fm = free_at(P, fr, i, j, k - 1), fp = free_at(P, fr, i, j, k + 1);
    acc += -0.5f * ((fm ? gv1[g - 1].z : 0.f) + (fp ? 0.f : gc.z)) + 0.5f * ((fp ? gv1[g + 1].z : 0.f) + (fm ? 0.f : gc.z)); }
  P.gp[o1 + g] += acc;
}
// divergence.grad (SF:126) as a gather
__global__ void __launch_bounds__(256) k_divergence_grad(const SParams P, const int s) {
  int i, j, k;
  if (!band_cell(P, i, j, k)) return;
  const size_t o0 = (size_t)s * P.G;
  const unsigned char* fr = P.fr + o0;
  const int g = cidx(P, i, j, k);
  if (!fr[g]) return;
  const float* gd = P.gdv + o0;
  const float hc = 0.5f * gd[g];
  const int n2 = P.n * P.n;
  float4 t = P.gvt[o0 + g];
  { const bool fm = free_at(P, fr, i - 1, j, k), fp = free_at(P, fr, i + 1, j, k);
    t.x += (fm ? 0.5f * gd[g - n2] : 0.f) - (fp ? 0.5f * gd[g + n2] : 0.f) + hc * ((fm ? 0.f : 1.f) - (fp ? 0.f : 1.f)); }
  { const bool fm = free_at(P, fr, i, j - 1, k), fp = free_at(P, fr, i, j + 1, k);
    t.y += (fm ? 0.5f * gd[g - P.n] : 0.f) - (fp ? 0.5f * gd[g + P.n] : 0.f) + hc * ((fm ? 0.f : 1.f) - (fp ? 0.f : 1.f)); }
  { const bool fm = free_at(P, fr, i, j, k - 1), fp = free_at(P, fr, i, j, k + 1);
    t.z += (fm ? 0.5f * gd[g - 1] : 0.f) - (fp ? 0.5f * gd[g + 1] : 0.f) + hc * ((fm ? 0.f : 1.f) - (fp ? 0.f : 1.f)); }
  P.gvt[o0 + g] = t;
}

// adjoint of one trilinear lookup of the vector field v[s] at `pp` with output adjoint go[3]: scatter into gv[s], accumulate d/dp into gp_ (may be null)
__device__ __forceinline__ void trilerp_v_adj(const SParams& P, const unsigned char* fr, const float4* vs, float4* gvs, const float* pp, const float* out, const float* go,
                                              float* gp_) {
  Tri t; tri_setup<true>(P, fr, pp, t);
  const float gW = -(go[0] * out[0] + go[1] * out[1] + go[2] * out[2]) / t.W;
#pragma unroll
  for (int m = 0; m < 8; m++) {
    const float4 a = vs[t.idx[m]];
    const float c = t.w[m] / t.W;
    atomicAdd(&gvs[t.idx[m]].x, go[0] * c); atomicAdd(&gvs[t.idx[m]].y, go[1] * c); atomicAdd(&gvs[t.idx[m]].z, go[2] * c);
    if (gp_) {
      const float gw = gW + (go[0] * a.x + go[1] * a.y + go[2] * a.z) / t.W;
      const int oi = m >> 2, oj = (m >> 1) & 1, ok = m & 1;
      gp_[0] += gw * (-t.sg[0][oi] * t.wx[1][oj] * t.wx[2][ok]);
      gp_[1] += gw * (-t.wx[0][oi] * t.sg[1][oj] * t.wx[2][ok]);
      gp_[2] += gw * (-t.wx[0][oi] * t.wx[1][oj] * t.sg[2][ok]);
    }
  }
}
// advect_and_impulse.grad (SF:127), free cells
__global__ void __launch_bounds__(128) k_advect_grad(const SParams P, const int s, const int f) {
  int i, j, k;
  const bool inb = band_cell(P, i, j, k);
  const size_t o0 = (size_t)s * P.G;
  const unsigned char* fr = P.fr + o0;
  float red[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // pos 3, quat 4, s, r
  if (inb && fr[cidx(P, i, j, k)]) {
    const int g = cidx(P, i, j, k);
    const float4* vs = P.v + o0; float4* gvs = P.gv + o0;
    const float p0[3] = {(float)i + 0.5f, (float)j + 0.5f, (float)k + 0.5f};
    float pe[3], v1[3], v2[3], v3[3], p1[3], p2[3];
    backtrace(P, fr, vs, p0, pe, v1, v2, v3, p1, p2);
    Tri te; tri_setup<true>(P, fr, pe, te);
    float vf[3]; trilerp_v(vs, te, vf);
    Impulse I; impulse(P, f, i, j, k, I);
    const float sf = P.as[f], rf = P.ar[f];
    const float4 gvt4 = P.gvt[o0 + g];
    const float gvt[3] = {gvt4.x, gvt4.y, gvt4.z};
    float g_factor = 0.f, g_pe[3] = {0.f, 0.f, 0.f};
    // q[s+1] = (1 - factor) q_f + factor low_T
    for (int a = 0; a < P.qd; a++) {
      const float* qs = P.q + ((size_t)s * P.qd + a) * P.G; float* gqs = P.gq + ((size_t)s * P.qd + a) * P.G;
      const float go = P.gq[((size_t)(s + 1) * P.qd + a) * P.G + g];
      const float qf = trilerp_s(qs, te);
      g_factor += go * (P.low_T - qf);
      const float gqf = (1.f - I.factor) * go;
      const float gW = -(gqf * qf) / te.W;
#pragma unroll
      for (int m = 0; m < 8; m++) {
        atomicAdd(&gqs[te.idx[m]], gqf * te.w[m] / te.W);
        const float gw = gW + gqf * qs[te.idx[m]] / te.W;
        const int oi = m >> 2, oj = (m >> 1) & 1, ok = m & 1;
        g_pe[0] += gw * (-te.sg[0][oi] * te.wx[1][oj] * te.wx[2][ok]);
        g_pe[1] += gw * (-te.wx[0][oi] * te.sg[1][oj] * te.wx[2][ok]);
        g_pe[2] += gw * (-te.wx[0][oi] * te.wx[1][oj] * te.sg[2][ok]);
      }
    }
    // momentum = (dir * s * factor) * dt
    float g_dir[3], dot = 0.f;
#pragma unroll
    for (int d = 0; d < 3; d++) { g_dir[d] = gvt[d] * sf * I.factor * P.dt; dot += gvt[d] * I.dir[d]; }
    red[7] = dot * I.factor * P.dt;
    g_factor += dot * sf * P.dt;
    const float g_dist = g_factor * I.factor * (-1.f / rf);
    red[8] = g_factor * I.factor * (I.dist / (rf * rf));
#pragma unroll
    for (int d = 0; d < 3; d++) red[d] = -(g_dist * I.d[d] / I.dist) / P.dx;
    q_rot_adj_q(P.aquat + (size_t)f * 4, P.inj, g_dir, red + 3);
    // v_f = trilerp(v, pe)
    trilerp_v_adj(P, fr, vs, gvs, pe, vf, gvt, g_pe);
    // pe = p0 - dt (2/9 v1 + 1/3 v2 + 4/9 v3); p2 = p0 - 0.75 dt v2; p1 = p0 - 0.5 dt v1
    float g_v1[3], g_v2[3], g_v3[3], g_p2[3] = {0.f, 0.f, 0.f}, g_p1[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < 3; d++) { g_v1[d] = -P.dt * (float)(2.0 / 9.0) * g_pe[d]; g_v2[d] = -P.dt * (float)(1.0 / 3.0) * g_pe[d]; g_v3[d] = -P.dt * (float)(4.0 / 9.0) * g_pe[d]; }
    trilerp_v_adj(P, fr, vs, gvs, p2, v3, g_v3, g_p2);
#pragma unroll
    for (int d = 0; d < 3; d++) g_v2[d] += -0.75f * P.dt * g_p2[d];
    trilerp_v_adj(P, fr, vs, gvs, p1, v2, g_v2, g_p1);
#pragma unroll
    for (int d = 0; d < 3; d++) g_v1[d] += -0.5f * P.dt * g_p1[d];
    trilerp_v_adj(P, fr, vs, gvs, p0, v1, g_v1, nullptr);
  }
  // block reduction of the air conditioner's adjoints
  __shared__ float sred[4][9];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 9; q++) {
    float x = red[q];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
    if (lane == 0) sred[warp][q] = x;
  }
  __syncthreads();
  if (threadIdx.x < 9 && P.gapos) {
    const int q = threadIdx.x;
    const float x = sred[0][q] + sred[1][q] + sred[2][q] + sred[3][q];
    if (x != 0.f) {
      float* dst = q < 3 ? P.gapos + (size_t)f * 3 + q : (q < 7 ? P.gaquat + (size_t)f * 4 + (q - 3) : (q == 7 ? P.gas + f : P.gar + f));
      atomicAdd(dst, x);
    }
  }
}
// non-free cells: q[s+1] = q[s]  ->  gq[s] += gq[s+1]   (runs after k_advect_grad on the same stream: no atomics needed)
__global__ void __launch_bounds__(256) k_advect_grad_nonfree(const SParams P, const int s) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P.G) return;
  if (P.fr[(size_t)s * P.G + g]) return;
  for (int a = 0; a < P.qd; a++) P.gq[((size_t)s * P.qd + a) * P.G + g] += P.gq[((size_t)(s + 1) * P.qd + a) * P.G + g];
}

// ================================================================================================================ host side
static int check(FsmkHandle* h, const char* name, int s, bool grad) {
  if (!h) return 1;
  if (!h->bound) { snprintf(h->err, sizeof(h->err), "%s: buffers were not bound", name); return 1; }
  if (s < 0 || s >= h->cfg.max_steps_local) { snprintf(h->err, sizeof(h->err), "%s: step %d out of range [0,%d)", name, s, h->cfg.max_steps_local); return 1; }
  if (grad && (!h->buf.gv || !h->buf.gv_tmp || !h->buf.gdiv || !h->buf.gp || !h->buf.gq)) { snprintf(h->err, sizeof(h->err), "%s: gradient buffers were not bound", name); return 1; }
  return 0;
}
static int check_air(FsmkHandle* h, const char* name) {
  if (!h->air.pos || !h->air.quat || !h->air.s || !h->air.r) { snprintf(h->err, sizeof(h->err), "%s: the air conditioner's arrays were not set (fsmk_set_aircon)", name); return 1; }
  return 0;
}
static inline int band_blocks(const SParams& P, int threads) { const long long t = (long long)P.n * P.H * P.n; return (int)((t + threads - 1) / threads); }

extern "C" int fsmk_create(const FsmkConfig* cfg, FsmkHandle** out) {
  if (!cfg || !out) return 1;
  if (cfg->res < 2 || cfg->res > 1024 || cfg->max_steps_local < 1 || cfg->q_dim < 1 || cfg->q_dim > 4 || cfg->solver_iters < 0) return 1;
  cudaError_t e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) return 2;   // no CUDA device: the product path fails loudly (there is no CPU fallback)
  FsmkHandle* h = new FsmkHandle();
  memset(h, 0, sizeof(*h));
  h->cfg = *cfg;
  const size_t smem = (size_t)JT_HMAX * JT_L * JT_L * (3 * sizeof(float) + 1);
  if (cudaFuncSetAttribute(k_jacobi_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
      cudaFuncSetAttribute(k_jacobi_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    delete h; return 2;
  }
  *out = h;
  return 0;
}
extern "C" void fsmk_destroy(FsmkHandle* h) { delete h; }
extern "C" const char* fsmk_last_error(FsmkHandle* h) { return h ? h->err : "null handle"; }
extern "C" int fsmk_bind(FsmkHandle* h, const FsmkBuffers* b) {
  if (!h || !b) return 1;
  if (!b->v || !b->v_tmp || !b->div || !b->p || !b->q || !b->is_free || !b->tmp_a || !b->tmp_b) { snprintf(h->err, sizeof(h->err), "fsmk_bind: null state buffer"); return 1; }
  if (b->gv && !b->acc) { snprintf(h->err, sizeof(h->err), "fsmk_bind: gradients need the acc scratch buffer"); return 1; }
  h->buf = *b; h->bound = true;
  return 0;
}
static void fill_sdf_s(SdfDev& d, const FmpmSdfMesh& m) {
  d.vox = (const float*)m.voxels; d.res = m.res; d.friction = m.friction; d.softness = m.softness;
  for (int i = 0; i < 12; i++) d.T[i] = m.T_mesh_to_voxels[i];
  for (int i = 0; i < 9; i++) d.Ainv[i] = 0.f;   // normals are not needed for is_collide
}
extern "C" int fsmk_set_statics(FsmkHandle* h, int n_statics, const FmpmSdfMesh* statics) {
  if (!h) return 1;
  if (n_statics < 0 || n_statics > 4 || (n_statics > 0 && !statics)) { snprintf(h->err, sizeof(h->err), "fsmk_set_statics: at most 4 statics (got %d)", n_statics); return 1; }
  for (int i = 0; i < n_statics; i++) {
    if (!statics[i].voxels || statics[i].res < 2) { snprintf(h->err, sizeof(h->err), "fsmk_set_statics: static %d has no SDF volume", i); return 1; }
    fill_sdf_s(h->statics[i], statics[i]);
  }
  h->n_statics = n_statics;
  return 0;
}
extern "C" int fsmk_set_aircon(FsmkHandle* h, const FsmkAircon* a) {
  if (!h || !a) return 1;
  h->air = *a;
  return 0;
}

extern "C" int fsmk_free_space(FsmkHandle* h, int s, void* stream) {
  if (check(h, "fsmk_free_space", s, false)) return 1;
  SParams P = make_sparams(h);
  if (P.H == 0) return 0;
  FMPM_LAUNCH(k_free_space, band_blocks(P, 256), 256, 0,stream, P, s);
  FSMK_CHECK_LAUNCH(h, "fsmk_free_space");
  return 0;
}
extern "C" int fsmk_advect(FsmkHandle* h, int s, int f, void* stream) {
  if (check(h, "fsmk_advect", s, false) || check_air(h, "fsmk_advect")) return 1;
  SParams P = make_sparams(h);
  FMPM_LAUNCH(k_advect, (P.G + 255) / 256, 256, 0, stream, P, s, f);
  FSMK_CHECK_LAUNCH(h, "fsmk_advect");
  return 0;
}
extern "C" int fsmk_divergence(FsmkHandle* h, int s, void* stream) {
  if (check(h, "fsmk_divergence", s, false)) return 1;
  SParams P = make_sparams(h);
  if (P.H == 0) return 0;
  FMPM_LAUNCH(k_divergence, band_blocks(P, 256), 256, 0,stream, P, s);
  FSMK_CHECK_LAUNCH(h, "fsmk_divergence");
  return 0;
}
// the chain of Jacobi launches: src -> (tmp_a <-> tmp_b) -> dst
template <bool kGrad>
static int jacobi_chain(FsmkHandle* h, const SParams& P, int s, const float* src, float* dst, void* stream, const char* name) {
  const int iters = h->cfg.solver_iters;
  float* ta = (float*)h->buf.tmp_a; float* tb = (float*)h->buf.tmp_b; float* acc = (float*)h->buf.acc;
  const bool tiled = P.H <= JT_HMAX;
  const int per = tiled ? JT_HALO : 1;
  const int launches = iters == 0 ? 1 : (iters + per - 1) / per;
  const float* in = src;
  int done = 0;
  for (int l = 0; l < launches; l++) {
    const int ns = iters - done < per ? iters - done : per;
    float* out = (l == launches - 1) ? dst : ((l & 1) ? tb : ta);
    if (tiled) {
      const dim3 grid((P.n + JT_OUT - 1) / JT_OUT, (P.n + JT_OUT - 1) / JT_OUT);
      const size_t smem = (size_t)P.H * JT_L * JT_L * (3 * sizeof(float) + 1);
      FMPM_LAUNCH(k_jacobi_tile<kGrad>, grid, 256, smem, stream, P, s, in, out, acc, ns, l == 0);
    } else {
      FMPM_LAUNCH(k_jacobi_simple<kGrad>, band_blocks(P, 256), 256, 0,stream, P, s, in, out, acc, ns, l == 0);
    }
    FSMK_CHECK_LAUNCH(h, name);
    in = out; done += ns;
  }
  return 0;
}
extern "C" int fsmk_pressure(FsmkHandle* h, int s, void* stream) {
  if (check(h, "fsmk_pressure", s, false)) return 1;
  SParams P = make_sparams(h);
  if (P.H == 0) return 0;
  return jacobi_chain<false>(h, P, s, P.p + (size_t)s * P.G, P.p + (size_t)(s + 1) * P.G, stream, "fsmk_pressure");
}
extern "C" int fsmk_project(FsmkHandle* h, int s, void* stream) {
  if (check(h, "fsmk_project", s, false)) return 1;
  SParams P = make_sparams(h);
  if (P.H == 0) return 0;
  FMPM_LAUNCH(k_project, band_blocks(P, 256), 256, 0,stream, P, s);
  FSMK_CHECK_LAUNCH(h, "fsmk_project");
  return 0;
}
extern "C" int fsmk_step(FsmkHandle* h, int s, int f, void* stream) {
  if (fsmk_free_space(h, s, stream) || fsmk_advect(h, s, f, stream) || fsmk_divergence(h, s, stream) || fsmk_pressure(h, s, stream)) return 1;
  return fsmk_project(h, s, stream);
}

extern "C" int fsmk_project_grad(FsmkHandle* h, int s, void* stream) {
  if (check(h, "fsmk_project_grad", s, true)) return 1;
  SParams P = make_sparams(h);
  if (P.H == 0) return 0;
  FMPM_LAUNCH(k_project_grad, band_blocks(P, 256), 256, 0,stream, P, s);
  FSMK_CHECK_LAUNCH(h, "fsmk_project_grad");
  return 0;
}
extern "C" int fsmk_pressure_grad(FsmkHandle* h, int s, void* stream) {
  if (check(h, "fsmk_pressure_grad", s, true)) return 1;
  SParams P = make_sparams(h);
  if (P.H == 0) return 0;
  // pressure_from_swap.grad seeds the solve with gp[s+1] on the free cells (SF:118); the result lands in tmp_a or tmp_b
  const int per = P.H <= JT_HMAX ? JT_HALO : 1;
  const int launches = h->cfg.solver_iters == 0 ? 1 : (h->cfg.solver_iters + per - 1) / per;
  float* res = (float*)(((launches - 1) & 1) ? h->buf.tmp_b : h->buf.tmp_a);
  if (jacobi_chain<true>(h, P, s, P.gp + (size_t)(s + 1) * P.G, res, stream, "fsmk_pressure_grad")) return 1;
  FMPM_LAUNCH(k_pressure_grad_finish, band_blocks(P, 256), 256, 0,stream, P, s, res, (const float*)h->buf.acc);
  FSMK_CHECK_LAUNCH(h, "fsmk_pressure_grad(finish)");
  return 0;
}
extern "C" int fsmk_divergence_grad(FsmkHandle* h, int s, void* stream) {
  if (check(h, "fsmk_divergence_grad", s, true)) return 1;
  SParams P = make_sparams(h);
  if (P.H == 0) return 0;
  FMPM_LAUNCH(k_divergence_grad, band_blocks(P, 256), 256, 0,stream, P, s);
  FSMK_CHECK_LAUNCH(h, "fsmk_divergence_grad");
  return 0;
}
extern "C" int fsmk_advect_grad(FsmkHandle* h, int s, int f, void* stream) {
  if (check(h, "fsmk_advect_grad", s, true) || check_air(h, "fsmk_advect_grad")) return 1;
  SParams P = make_sparams(h);
  if (P.H > 0) {
    FMPM_LAUNCH(k_advect_grad, band_blocks(P, 128), 128, 0,stream, P, s, f);
    FSMK_CHECK_LAUNCH(h, "fsmk_advect_grad");
  }
  FMPM_LAUNCH(k_advect_grad_nonfree, (P.G + 255) / 256, 256, 0, stream, P, s);
  FSMK_CHECK_LAUNCH(h, "fsmk_advect_grad(nonfree)");
  return 0;
}
extern "C" int fsmk_step_grad(FsmkHandle* h, int s, int f, void* stream) {
  if (fsmk_free_space(h, s, stream) || fsmk_project_grad(h, s, stream) || fsmk_pressure_grad(h, s, stream) || fsmk_divergence_grad(h, s, stream)) return 1;
  return fsmk_advect_grad(h, s, f, stream);
}
