// fmpm_forward.cu — forward MLS-MPM substep kernels for sm_100a (B200).
//
// Reference semantics: fluidlab/fluidengine/simulators/mpm_simulator.py (MPM) substep(), MPM:515-533.
// Reference structure (17 Taichi launches per substep, per-particle float atomics, F_tmp/U/S/V round
// trips through global memory) is NOT followed.  B200 design:
//
//  * particles live in cell-sorted SLOT order in float4 planes -> every particle load/store is a fully
//    coalesced 128-bit access;
//  * p2g fuses F_tmp + SVD + stress + scatter + F-update (MPM:254-264, 331-378).  The scatter does not
//    issue one atomic per (particle,node,component): shared-memory float atomics are CAS loops on
//    sm_100a (ATOMS.CAST.SPIN), so instead each warp walks its sorted particles with lane = stencil
//    node (27 of 32 lanes), accumulating the node sums of the current cell in REGISTERS.  When the walk
//    moves to the next cell of the z-column the window shifts by one plane through warp shuffles and only
//    the finished 3x3 plane is flushed with one vector reduction (REDG.E.ADD.F32x4: momentum xyz + mass
//    in a single 16-byte L2 atomic).  ~1.5 vector REDs per particle instead of 108 scalar atomics.
//  * grid_op (MPM:380-398) also clears the momentum/mass accumulators for the next substep;
//  * g2p fuses advect_used / process_unused_particles / g2p / advect_kernel (MPM:304-316, 400-426, 497-505).
#include <cstdio>
#include "fmpm_common.cuh"

#define FULL_MASK 0xffffffffu
#define P2G_WARPS 4
#define P2G_ROUNDS 4
#define WSTR 33

__device__ __forceinline__ void red_add_v4(float4* addr, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct __align__(16) ScatterSmem {
  float4 uni[32 * 4];      // per particle: (q0,q1,q2,m) (B00,B01,B02,B10) (B11,B12,B20,B21) (B22,key,-,-)
  float w[27 * WSTR + 5];  // w[node*33 + particle]  (stride 33: conflict-free for both access patterns)
};

// Sliding-window register scatter shared by p2g (momentum+mass) and the g2p adjoint (v_out adjoint).
// Per lane (= stencil node (a,b,c), lane = a*9+b*3+c): acc += w * (q + B·(a,b,c)), acc.w += w*m.
struct Window {
  float4 acc; int cur_key;
  float oa, ob, oc; int c; int lane_off; bool lane_valid; int wrow;
};
__device__ __forceinline__ void window_init(Window& W, int lane, int n) {
  int L = lane < 27 ? lane : 26;
  int a = L / 9, b = (L / 3) % 3, c = L % 3;
  W.oa = (float)a; W.ob = (float)b; W.oc = (float)c; W.c = c;
  W.lane_off = (a * n + b) * n + c;
  W.lane_valid = lane < 27;
  W.wrow = L * WSTR;
  W.acc = make_float4(0.f, 0.f, 0.f, 0.f);
  W.cur_key = -1;
}
__device__ __forceinline__ void window_flush_all(Window& W, float4* __restrict__ grid) {
  if (W.cur_key >= 0 && W.lane_valid) red_add_v4(grid + W.cur_key + W.lane_off, W.acc);
  W.acc = make_float4(0.f, 0.f, 0.f, 0.f);
  W.cur_key = -1;
}
__device__ __forceinline__ void window_consume(Window& W, const ScatterSmem& S, int cnt, float4* __restrict__ grid) {
  for (int j = 0; j < cnt; j++) {
    const float4 u3 = S.uni[j * 4 + 3];
    const int key = __float_as_int(u3.y);
    if (key < 0) continue;  // warp-uniform
    if (key != W.cur_key) {
      if (W.cur_key >= 0) {
        if (key == W.cur_key + 1) {  // next cell of the same z-column: plane c=0 is complete
          if (W.lane_valid && W.c == 0) red_add_v4(grid + W.cur_key + W.lane_off, W.acc);
          float4 t;
          t.x = __shfl_down_sync(FULL_MASK, W.acc.x, 1); t.y = __shfl_down_sync(FULL_MASK, W.acc.y, 1);
          t.z = __shfl_down_sync(FULL_MASK, W.acc.z, 1); t.w = __shfl_down_sync(FULL_MASK, W.acc.w, 1);
          W.acc = (W.c == 2 || !W.lane_valid) ? make_float4(0.f, 0.f, 0.f, 0.f) : t;
        } else {
          if (W.lane_valid) red_add_v4(grid + W.cur_key + W.lane_off, W.acc);
          W.acc = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      W.cur_key = key;
    }
    const float4 u0 = S.uni[j * 4], u1 = S.uni[j * 4 + 1], u2 = S.uni[j * 4 + 2];
    const float w = S.w[W.wrow + j];
    float t0 = fmaf(u1.z, W.oc, fmaf(u1.y, W.ob, fmaf(u1.x, W.oa, u0.x)));
    float t1 = fmaf(u2.y, W.oc, fmaf(u2.x, W.ob, fmaf(u1.w, W.oa, u0.y)));
    float t2 = fmaf(u3.x, W.oc, fmaf(u2.w, W.ob, fmaf(u2.z, W.oa, u0.z)));
    W.acc.x = fmaf(w, t0, W.acc.x); W.acc.y = fmaf(w, t1, W.acc.y); W.acc.z = fmaf(w, t2, W.acc.z);
    W.acc.w = fmaf(w, u0.w, W.acc.w);
  }
}
// lane = particle: publish the scatter record (q, B, m, key, 27 weights) to the warp's shared staging area
__device__ __forceinline__ void scatter_publish(ScatterSmem& S, int lane, int key, const float* q, const float* B, float m, const float w[3][3]) {
  S.uni[lane * 4 + 0] = make_float4(q[0], q[1], q[2], m);
  S.uni[lane * 4 + 1] = make_float4(B[0], B[1], B[2], B[3]);
  S.uni[lane * 4 + 2] = make_float4(B[4], B[5], B[6], B[7]);
  S.uni[lane * 4 + 3] = make_float4(B[8], __int_as_float(key), 0.f, 0.f);
  if (key >= 0) {
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++) {
        float wab = w[a][0] * w[b][1];
#pragma unroll
        for (int c = 0; c < 3; c++) S.w[(a * 9 + b * 3 + c) * WSTR + lane] = wab * w[c][2];
      }
  }
}

// =============================================================================================
// p2g
// =============================================================================================
template <bool kWriteF>
__global__ void __launch_bounds__(P2G_WARPS * 32) k_p2g(const KParams P, const int f) {
  __shared__ ScatterSmem smem[P2G_WARPS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  ScatterSmem& S = smem[wib];
  const long long gw = (long long)blockIdx.x * P2G_WARPS + wib;
  const long long slot0 = gw * (32 * P2G_ROUNDS);
  if (slot0 >= P.N) return;
  Window W; window_init(W, lane, P.n);
#pragma unroll 1
  for (int r = 0; r < P2G_ROUNDS; r++) {
    const long long sl = slot0 + r * 32 + lane;
    const long long rem = (long long)P.N - (slot0 + r * 32);
    if (rem <= 0) break;  // warp-uniform
    const int cnt = rem < 32 ? (int)rem : 32;
    int key = -1;
    float q[3] = {0.f, 0.f, 0.f}, B[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, m = 0.f;
    float w[3][3];
    if (sl < P.N) {
      const int s = (int)sl;
      PState st; load_A(P.pa, P, f, s, st); load_F(P.pf, P.pf8, P, f, s, st.F);
      int b[3]; float fx[3];
      const bool used = st.meta & 1;
      const bool ok = used && base_fx(P, st.x, b, fx);
      if (ok) {
        const float4 mt = __ldg(P.mats + ((st.meta >> 8) & 0xff));
        Constit K; constitutive(P, st, mt.x, mt.y, mt.z, __float_as_int(mt.w), K);
        m = mt.z;
        bspline(fx, w);
        // contribution_i = w_i * (m v + A (o - fx) dx) = w_i * (q + B o),  B = A dx,  q = m v - B fx
#pragma unroll
        for (int i = 0; i < 9; i++) B[i] = K.A.m[i] * P.dx;
#pragma unroll
        for (int i = 0; i < 3; i++) q[i] = m * st.v[i] - (B[i * 3] * fx[0] + B[i * 3 + 1] * fx[1] + B[i * 3 + 2] * fx[2]);
        key = (b[0] * P.n + b[1]) * P.n + b[2];
        if (kWriteF) store_F(P.pf, P.pf8, P, f + 1, s, K.Fn);
      } else if (kWriteF) {
        store_F(P.pf, P.pf8, P, f + 1, s, st.F);  // process_unused_particles (MPM:316) / frozen out-of-grid particle
      }
    }
    scatter_publish(S, lane, key, q, B, m, w);
    __syncwarp();
    window_consume(W, S, cnt, P.grid_pm);
    __syncwarp();
  }
  window_flush_all(W, P.grid_pm);
}

// =============================================================================================
// grid_op
// =============================================================================================
__global__ void __launch_bounds__(256) k_grid_op(const KParams P, const int clear_pm) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P.G) return;
  const float4 pm = P.grid_pm[g];
  float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pm.w > FMPM_EPS) {
    const float inv_m = 1.f / pm.w;
    float v[3] = {inv_m * pm.x + P.dt * P.gx, inv_m * pm.y + P.dt * P.gy, inv_m * pm.z + P.dt * P.gz};
    const int n = P.n;
    const int i = g / (n * n), j = (g / n) % n, k = g % n;
    const float pos[3] = {(float)i * P.dx, (float)j * P.dx, (float)k * P.dx};
    float fac[3];
    boundary_v(P, pos, v, fac);
    out = make_float4(v[0], v[1], v[2], 0.f);
  }
  P.grid_v[g] = out;
  if (clear_pm && (pm.w != 0.f || pm.x != 0.f || pm.y != 0.f || pm.z != 0.f)) P.grid_pm[g] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// =============================================================================================
// g2p (+ advect_used, process_unused_particles, advect_kernel)
// =============================================================================================
__global__ void __launch_bounds__(128) k_g2p(const KParams P, const int f) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= P.N) return;
  const float4 a0 = P.pa[pa_idx(P, f, 0, s)];
  const int meta = __float_as_int(a0.w);
  const float x[3] = {a0.x, a0.y, a0.z};
  int b[3]; float fx[3];
  const bool ok = (meta & 1) && base_fx(P, x, b, fx);
  if (!ok) {  // unused (MPM:309-316) or frozen: carry the state over unchanged
    P.pa[pa_idx(P, f + 1, 0, s)] = a0;
    P.pa[pa_idx(P, f + 1, 1, s)] = P.pa[pa_idx(P, f, 1, s)];
    P.pa[pa_idx(P, f + 1, 2, s)] = P.pa[pa_idx(P, f, 2, s)];
    P.pa[pa_idx(P, f + 1, 3, s)] = P.pa[pa_idx(P, f, 3, s)];
    return;
  }
  float w[3][3]; bspline(fx, w);
  float nv[3] = {0.f, 0.f, 0.f};
  Mat3 nC = m3_zero();
  const float4* __restrict__ gv = P.grid_v + ((b[0] * P.n + b[1]) * P.n + b[2]);
  const float c4 = 4.f * P.inv_dx;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const float wij = w[i][0] * w[j][1];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float4 g = __ldg(gv + (i * P.n + j) * P.n + k);
        const float wt = wij * w[k][2];
        const float d0 = (float)i - fx[0], d1 = (float)j - fx[1], d2 = (float)k - fx[2];
        const float wg0 = wt * g.x, wg1 = wt * g.y, wg2 = wt * g.z;
        nv[0] += wg0; nv[1] += wg1; nv[2] += wg2;
        const float s0 = c4 * wg0, s1 = c4 * wg1, s2 = c4 * wg2;
        nC.m[0] = fmaf(s0, d0, nC.m[0]); nC.m[1] = fmaf(s0, d1, nC.m[1]); nC.m[2] = fmaf(s0, d2, nC.m[2]);
        nC.m[3] = fmaf(s1, d0, nC.m[3]); nC.m[4] = fmaf(s1, d1, nC.m[4]); nC.m[5] = fmaf(s1, d2, nC.m[5]);
        nC.m[6] = fmaf(s2, d0, nC.m[6]); nC.m[7] = fmaf(s2, d1, nC.m[7]); nC.m[8] = fmaf(s2, d2, nC.m[8]);
      }
    }
  const float nx[3] = {x[0] + P.dt * nv[0], x[1] + P.dt * nv[1], x[2] + P.dt * nv[2]};  // advect_kernel MPM:505
  store_A(P.pa, P, f + 1, s, nx, meta, nv, nC);
}

// =============================================================================================
// injector act (agents/agent_injector.py:30-32 -> effectors/injector.py:80-105, 240-256)
// =============================================================================================
__device__ __forceinline__ void quat_rot(const float* q, const float* v, float* o) {  // utils/geom.py:92-97
  float uv[3] = {q[2] * v[2] - q[3] * v[1], q[3] * v[0] - q[1] * v[2], q[1] * v[1] - q[2] * v[0]};
  float uuv[3] = {q[2] * uv[2] - q[3] * uv[1], q[3] * uv[0] - q[1] * uv[2], q[1] * uv[1] - q[2] * uv[0]};
#pragma unroll
  for (int k = 0; k < 3; k++) o[k] = v[k] + 2.f * (q[0] * uv[k] + uuv[k]);
}
__global__ void k_inject(const KParams P, const int f, const FmpmInjector inj, const float* __restrict__ epos,
                         const float* __restrict__ equat, const int act_id, const int rand_row, const int* __restrict__ inv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= inj.flux) return;
  const int pid = ((const int*)inj.act_range)[act_id + i];
  const int s = inv ? inv[pid] : pid;
  const float* rv = (const float*)inj.random_vector + ((size_t)rand_row * inj.flux + i) * 3;
  const float* pos = epos + f * 3;
  const float* quat = equat + f * 4;
  float x[3], v[3];
  if (inj.kind == 1) {
    float ipr[3]; quat_rot(quat, inj.inject_p, ipr);
#pragma unroll
    for (int k = 0; k < 3; k++) x[k] = (rv[k] * 2.f - 1.f) * inj.radius + pos[k] + ipr[k];
    quat_rot(quat, inj.inject_v, v);
  } else {
#pragma unroll
    for (int k = 0; k < 3; k++) { x[k] = rv[k] + pos[k]; v[k] = inj.inject_v[k]; }
  }
  float4 a0 = P.pa[pa_idx(P, f + 1, 0, s)], a1 = P.pa[pa_idx(P, f + 1, 1, s)];
  const int meta = __float_as_int(a0.w) | 1;  // used[f+1, pid] = 1
  P.pa[pa_idx(P, f + 1, 0, s)] = make_float4(x[0], x[1], x[2], __int_as_float(meta));
  P.pa[pa_idx(P, f + 1, 1, s)] = make_float4(v[0], v[1], v[2], a1.w);
}

// =============================================================================================
// host entry points
// =============================================================================================
static int check_bound(FmpmHandle* h, const char* name) {
  if (!h) return 1;
  if (!h->bound) { snprintf(h->err, sizeof(h->err), "%s: fmpm_bind() has not been called", name); return 1; }
  return 0;
}
static int check_frame(FmpmHandle* h, int f, int maxf, const char* name) {
  if (f < 0 || f > maxf) { snprintf(h->err, sizeof(h->err), "%s: frame %d out of range [0,%d]", name, f, maxf); return 1; }
  return 0;
}

extern "C" int fmpm_clear_grid(FmpmHandle* h, void* stream) {
  if (check_bound(h, "fmpm_clear_grid")) return 1;
  const size_t G = (size_t)h->cfg.n_grid * h->cfg.n_grid * h->cfg.n_grid;
  cudaError_t e = cudaMemsetAsync(h->buf.grid_pm, 0, G * sizeof(float4), (cudaStream_t)stream);
  if (e != cudaSuccess) { snprintf(h->err, sizeof(h->err), "fmpm_clear_grid: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

extern "C" int fmpm_p2g(FmpmHandle* h, int f, int write_F, void* stream) {
  if (check_bound(h, "fmpm_p2g") || check_frame(h, f, h->cfg.max_substeps_local - (write_F ? 1 : 0), "fmpm_p2g")) return 1;
  KParams P = make_kparams(h);
  if (P.N == 0) return 0;
  const long long warps = ((long long)P.N + 32 * P2G_ROUNDS - 1) / (32 * P2G_ROUNDS);
  const int blocks = (int)((warps + P2G_WARPS - 1) / P2G_WARPS);
  if (write_F) k_p2g<true><<<blocks, P2G_WARPS * 32, 0, (cudaStream_t)stream>>>(P, f);
  else k_p2g<false><<<blocks, P2G_WARPS * 32, 0, (cudaStream_t)stream>>>(P, f);
  FMPM_CHECK_LAUNCH(h, "fmpm_p2g");
  return 0;
}

extern "C" int fmpm_grid_op(FmpmHandle* h, int f, int clear_pm, void* stream) {
  (void)f;
  if (check_bound(h, "fmpm_grid_op")) return 1;
  KParams P = make_kparams(h);
  k_grid_op<<<(P.G + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, clear_pm);
  FMPM_CHECK_LAUNCH(h, "fmpm_grid_op");
  return 0;
}

extern "C" int fmpm_g2p(FmpmHandle* h, int f, void* stream) {
  if (check_bound(h, "fmpm_g2p") || check_frame(h, f, h->cfg.max_substeps_local - 1, "fmpm_g2p")) return 1;
  KParams P = make_kparams(h);
  if (P.N == 0) return 0;
  k_g2p<<<(P.N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P, f);
  FMPM_CHECK_LAUNCH(h, "fmpm_g2p");
  return 0;
}

extern "C" int fmpm_substep(FmpmHandle* h, int f, void* stream) {
  if (fmpm_p2g(h, f, 1, stream)) return 1;
  if (fmpm_grid_op(h, f, 1, stream)) return 1;
  return fmpm_g2p(h, f, stream);
}

extern "C" int fmpm_inject(FmpmHandle* h, int f, const FmpmInjector* inj, const FmpmEffector* e, int act_id, int rand_row,
                           const void* inv, void* stream) {
  if (check_bound(h, "fmpm_inject") || check_frame(h, f, h->cfg.max_substeps_local - 1, "fmpm_inject")) return 1;
  if (act_id < 0 || act_id + inj->flux > inj->n_act_range) {
    snprintf(h->err, sizeof(h->err), "fmpm_inject: too many particles added (act_id %d + flux %d > %d)", act_id, inj->flux, inj->n_act_range);
    return 2;
  }
  KParams P = make_kparams(h);
  k_inject<<<(inj->flux + 31) / 32, 32, 0, (cudaStream_t)stream>>>(P, f, *inj, (const float*)e->pos, (const float*)e->quat, act_id, rand_row, (const int*)inv);
  FMPM_CHECK_LAUNCH(h, "fmpm_inject");
  return 0;
}
