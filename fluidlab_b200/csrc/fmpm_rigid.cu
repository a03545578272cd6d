// fmpm_rigid.cu — MAT_RIGID bodies: rigidity enforcement by shape matching and its adjoint (sm_100a).
//
// Reference: fluidlab/fluidengine/simulators/mpm_simulator.py
//   forward  MPM:428-434 advect = reset_bodies_and_grad (:449-454), compute_COM (:456-462), compute_H (:464-477),
//            compute_H_svd (:479-483), compute_R (:491-495), advect_kernel (:497-505)
//   adjoint  MPM:436-447 advect_grad = advect_kernel.grad, compute_R.grad, compute_H_svd_grad (:485-489, manual SVD adjoint
//            MPM:272-292), compute_H.grad, compute_COM.grad
//
// fmpm_g2p has already written v[f+1] and the free-particle position y = x[f] + dt v[f+1] into frame f+1; for particles of a
// MAT_RIGID body this file replaces y by R (x - c0) + c1 with c0/c1 the body's centre of mass before/after and R = V U^T from
// svd(H), H = sum (x - c0)(y - c1)^T.  Rigid scenes are small (a few thousand rigid particles): the kernels sweep the slot
// range, reduce inside the warp (a warp of the cell-sorted order almost always holds one body) and finish with a few atomics
// per warp.  The body state of every ring substep stays in HBM so the adjoint does not recompute it.
#include <cstdio>
#include <cstring>
#include "fmpm_common.cuh"

namespace {

constexpr int BS = FMPM_BODY_STATE_STRIDE, BG = FMPM_BODY_GRAD_STRIDE;
// state offsets
constexpr int O_C0 = 0, O_C1 = 3, O_H = 6, O_U = 15, O_S = 24, O_V = 27, O_R = 36;
// grad offsets
constexpr int G_R = 0, G_SGX = 9, G_H = 12, G_C0 = 21, G_C1 = 24;

__device__ __forceinline__ bool rigid_slot(const int meta, const int* __restrict__ info, const int nb, int& b) {
  b = (meta >> 16) & 0xff;
  return (meta & 1) && b < nb && __ldg(info + 2 * b + 1) == FMPM_MAT_RIGID;
}

// every lane of the warp calls this; lanes with active == false contribute nothing.
template <int NV>
__device__ __forceinline__ void warp_body_add(float* __restrict__ dst, const int stride, const int b, const bool active, const float* vals) {
  const unsigned full = 0xffffffffu;
  const unsigned act = __ballot_sync(full, active);
  if (act == 0u) return;
  const int leader = __ffs(act) - 1;
  const int b0 = __shfl_sync(full, b, leader);
  const bool uniform = __all_sync(full, !active || b == b0);
  if (uniform) {
#pragma unroll
    for (int k = 0; k < NV; k++) {
      float v = active ? vals[k] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(dst + (size_t)b0 * stride + k, v);
    }
  } else if (active) {
#pragma unroll
    for (int k = 0; k < NV; k++) atomicAdd(dst + (size_t)b * stride + k, vals[k]);
  }
}

// MPM:456-462 compute_COM
__global__ void k_body_com(const KParams P, const int f, const int* __restrict__ info, float* __restrict__ st, const int nb) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  bool act = false; int b = 0; float vals[6] = {0, 0, 0, 0, 0, 0};
  if (s < P.N) {
    const float4 a0 = P.pa[pa_idx(P, f, 0, s)];
    act = rigid_slot(__float_as_int(a0.w), info, nb, b);
    if (act) {
      const float4 y = P.pa[pa_idx(P, f + 1, 0, s)];
      const float n = (float)__ldg(info + 2 * b);
      vals[0] = a0.x / n; vals[1] = a0.y / n; vals[2] = a0.z / n;
      vals[3] = y.x / n; vals[4] = y.y / n; vals[5] = y.z / n;
    }
  }
  warp_body_add<6>(st + O_C0, BS, b, act, vals);
}

// MPM:464-477 compute_H
__global__ void k_body_H(const KParams P, const int f, const int* __restrict__ info, float* __restrict__ st, const int nb) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  bool act = false; int b = 0; float vals[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (s < P.N) {
    const float4 a0 = P.pa[pa_idx(P, f, 0, s)];
    act = rigid_slot(__float_as_int(a0.w), info, nb, b);
    if (act) {
      const float4 y = P.pa[pa_idx(P, f + 1, 0, s)];
      const float* B = st + (size_t)b * BS;
      const float d0[3] = {a0.x - B[O_C0], a0.y - B[O_C0 + 1], a0.z - B[O_C0 + 2]};
      const float d1[3] = {y.x - B[O_C1], y.y - B[O_C1 + 1], y.z - B[O_C1 + 2]};
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) vals[i * 3 + j] = d0[i] * d1[j];
    }
  }
  warp_body_add<9>(st + O_H, BS, b, act, vals);
}

// MPM:479-495 compute_H_svd + compute_R
__global__ void k_body_solve(const int* __restrict__ info, float* __restrict__ st, const int nb) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb || info[2 * b + 1] != FMPM_MAT_RIGID) return;
  float* B = st + (size_t)b * BS;
  Mat3 H, U, V; float sig[3];
#pragma unroll
  for (int i = 0; i < 9; i++) H.m[i] = B[O_H + i];
  svd3(H, U, sig, V);
  const Mat3 R = m3_mul_nt(V, U);
#pragma unroll
  for (int i = 0; i < 9; i++) { B[O_U + i] = U.m[i]; B[O_V + i] = V.m[i]; B[O_R + i] = R.m[i]; }
  B[O_S] = sig[0]; B[O_S + 1] = sig[1]; B[O_S + 2] = sig[2];
}

// MPM:497-505 advect_kernel, rigid branch (the other branch was written by k_g2p)
__global__ void k_body_advect(const KParams P, const int f, const int* __restrict__ info, const float* __restrict__ st, const int nb) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= P.N) return;
  const float4 a0 = P.pa[pa_idx(P, f, 0, s)];
  int b;
  if (!rigid_slot(__float_as_int(a0.w), info, nb, b)) return;
  const float* B = st + (size_t)b * BS;
  const float d[3] = {a0.x - B[O_C0], a0.y - B[O_C0 + 1], a0.z - B[O_C0 + 2]};
  float4 o = P.pa[pa_idx(P, f + 1, 0, s)];
  o.x = B[O_R + 0] * d[0] + B[O_R + 1] * d[1] + B[O_R + 2] * d[2] + B[O_C1];
  o.y = B[O_R + 3] * d[0] + B[O_R + 4] * d[1] + B[O_R + 5] * d[2] + B[O_C1 + 1];
  o.z = B[O_R + 6] * d[0] + B[O_R + 7] * d[1] + B[O_R + 8] * d[2] + B[O_C1 + 2];
  P.pa[pa_idx(P, f + 1, 0, s)] = o;
}

// ---------------------------------------------------------------------------------------------------------------------------
// adjoint.  gin holds the adjoint of frame f+1 in the slot order of frame f.
// ---------------------------------------------------------------------------------------------------------------------------
// advect_kernel.grad (rigid branch), body side: gR += gx' (x - c0)^T ,  sum_gx += gx'
__global__ void k_body_grad_reduce(const KParams P, const int f, const int gin, const int* __restrict__ info, const float* __restrict__ st,
                                   float* __restrict__ bg, const int nb) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  bool act = false; int b = 0; float vals[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (s < P.N) {
    const float4 a0 = P.pa[pa_idx(P, f, 0, s)];
    act = rigid_slot(__float_as_int(a0.w), info, nb, b);
    if (act) {
      const float4 g = P.ga[pa_idx(P, gin, 0, s)];
      const float* B = st + (size_t)b * BS;
      const float d0[3] = {a0.x - B[O_C0], a0.y - B[O_C0 + 1], a0.z - B[O_C0 + 2]};
      const float go[3] = {g.x, g.y, g.z};
#pragma unroll
      for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) vals[i * 3 + j] = go[i] * d0[j];
        vals[9 + i] = go[i];
      }
    }
  }
  warp_body_add<12>(bg + G_R, BG, b, act, vals);
}

__device__ __forceinline__ float clamp_svd_r(float a) { return a >= 0.f ? fmaxf(a, 1e-8f) : fminf(a, -1e-8f); }  // MPM:294-302

// compute_R.grad + compute_H_svd_grad (MPM:485-489 -> backward_svd MPM:272-292 with gS = 0).  With R = V U^T:
// gU = gR^T V, gV = gR U, M = U^T gU = U^T gR^T V, and the reference formula collapses to
//   gH = U Z V^T,  Z_ij = (M_ij - M_ji) (s_j - s_i) / clamp(s_j^2 - s_i^2)   (i != j; = (M_ij - M_ji)/(s_i + s_j) when unclamped)
// evaluated in that factored form (no cancellation of the two 1/(s_j^2 - s_i^2) terms).
__global__ void k_body_grad_solve(const int* __restrict__ info, const float* __restrict__ st, float* __restrict__ bg, const int nb) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb || info[2 * b + 1] != FMPM_MAT_RIGID) return;
  const float* B = st + (size_t)b * BS;
  float* Gb = bg + (size_t)b * BG;
  Mat3 U, V, R, gR;
#pragma unroll
  for (int i = 0; i < 9; i++) { U.m[i] = B[O_U + i]; V.m[i] = B[O_V + i]; R.m[i] = B[O_R + i]; gR.m[i] = Gb[G_R + i]; }
  const float s[3] = {B[O_S], B[O_S + 1], B[O_S + 2]};
  const Mat3 M = m3_mul(m3_mul_tn(U, m3_tr(gR)), V);
  Mat3 Z = m3_zero();
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      if (i == j) continue;
      const float d = s[j] * s[j] - s[i] * s[i];
      const float inv_sum = fabsf(d) >= 1e-8f ? 1.f / (s[i] + s[j]) : (s[j] - s[i]) / clamp_svd_r(d);
      Z.m[i * 3 + j] = (M.m[i * 3 + j] - M.m[j * 3 + i]) * inv_sum;
    }
  const Mat3 gH = m3_mul_nt(m3_mul(U, Z), V);
#pragma unroll
  for (int i = 0; i < 9; i++) Gb[G_H + i] = gH.m[i];
  const float sg[3] = {Gb[G_SGX], Gb[G_SGX + 1], Gb[G_SGX + 2]};
#pragma unroll
  for (int j = 0; j < 3; j++) {
    Gb[G_C0 + j] = -(R.m[j] * sg[0] + R.m[3 + j] * sg[1] + R.m[6 + j] * sg[2]);   // -R^T sum_gx
    Gb[G_C1 + j] = sg[j];
  }
}

// particle side of advect_kernel.grad, compute_H.grad, compute_COM.grad.  Rewrites gin in place so that the unchanged
// free-particle adjoint that follows (gx[f] += gx', gv' += dt gx') yields the rigid result:
//   gy  = gH^T (x - c0) + gc1 / n                         (adjoint of y = x + dt v')
//   gx* = R^T gx' + gH (y - c1) + gc0 / n + gy            (what must reach gx[f])
//   gx' <- gx*,   gv' <- gv' + dt (gy - gx*)
// The sum(y - c1) and sum(x - c0) contributions to gc0/gc1 (compute_H.grad) vanish identically and are dropped.
// next_slot: slot of frame f+1 holding the particle of slot s of frame f (NULL: same order).
__global__ void k_body_grad_apply(const KParams P, const int f, const int gin, const int* __restrict__ next_slot, const int* __restrict__ info,
                                  const float* __restrict__ st, const float* __restrict__ bg, const int nb) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= P.N) return;
  const float4 a0 = P.pa[pa_idx(P, f, 0, s)];
  int b;
  if (!rigid_slot(__float_as_int(a0.w), info, nb, b)) return;
  const float* B = st + (size_t)b * BS;
  const float* Gb = bg + (size_t)b * BG;
  const int s1 = next_slot ? next_slot[s] : s;
  const float4 v1 = P.pa[pa_idx(P, f + 1, 1, s1)];
  const float inv_n = 1.f / (float)__ldg(info + 2 * b);
  const float x[3] = {a0.x, a0.y, a0.z};
  const float vn[3] = {v1.x, v1.y, v1.z};
  float d0[3], d1[3];
#pragma unroll
  for (int k = 0; k < 3; k++) { d0[k] = x[k] - B[O_C0 + k]; d1[k] = x[k] + P.dt * vn[k] - B[O_C1 + k]; }
  float4 g = P.ga[pa_idx(P, gin, 0, s)];
  float4 gv = P.ga[pa_idx(P, gin, 1, s)];
  const float go[3] = {g.x, g.y, g.z};
  float gy[3], gxs[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    gy[i] = Gb[G_H + i] * d0[0] + Gb[G_H + 3 + i] * d0[1] + Gb[G_H + 6 + i] * d0[2] + Gb[G_C1 + i] * inv_n;
    gxs[i] = B[O_R + i] * go[0] + B[O_R + 3 + i] * go[1] + B[O_R + 6 + i] * go[2]
           + Gb[G_H + i * 3] * d1[0] + Gb[G_H + i * 3 + 1] * d1[1] + Gb[G_H + i * 3 + 2] * d1[2] + Gb[G_C0 + i] * inv_n + gy[i];
  }
  g.x = gxs[0]; g.y = gxs[1]; g.z = gxs[2];
  gv.x += P.dt * (gy[0] - gxs[0]); gv.y += P.dt * (gy[1] - gxs[1]); gv.z += P.dt * (gy[2] - gxs[2]);
  P.ga[pa_idx(P, gin, 0, s)] = g;
  P.ga[pa_idx(P, gin, 1, s)] = gv;
}

int check(FmpmHandle* h, int f, const char* name) {
  if (!h) return 1;
  if (!h->bound) { snprintf(h->err, sizeof(h->err), "%s: buffers not bound", name); return 1; }
  if (f < 0 || f >= h->cfg.max_substeps_local) { snprintf(h->err, sizeof(h->err), "%s: frame %d out of range [0,%d)", name, f, h->cfg.max_substeps_local); return 1; }
  return 0;
}

}  // namespace

extern "C" int fmpm_set_bodies(FmpmHandle* h, const FmpmBodies* b) {
  if (!h) return 1;
  memset(&h->bodies, 0, sizeof(h->bodies));
  if (!b || b->n_bodies == 0) return 0;
  if (b->n_bodies < 0 || b->n_bodies > 256) { snprintf(h->err, sizeof(h->err), "fmpm_set_bodies: n_bodies %d not in [0,256]", b->n_bodies); return 1; }
  if (!b->info || !b->state) { snprintf(h->err, sizeof(h->err), "fmpm_set_bodies: info and state buffers are required"); return 1; }
  if (h->slab.enabled) { snprintf(h->err, sizeof(h->err), "fmpm_set_bodies: MAT_RIGID bodies are not supported in x-slab mode"); return 1; }
  h->bodies = *b;
  return 0;
}

int fmpm_advect_rigid_impl(FmpmHandle* h, int f, void* stream) {
  const FmpmBodies& b = h->bodies;
  if (b.n_bodies == 0 || h->cfg.n_particles == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  KParams P = make_kparams(h);
  float* state = (float*)b.state + (size_t)f * b.n_bodies * BS;
  const int* info = (const int*)b.info;
  if (cudaMemsetAsync(state, 0, sizeof(float) * (size_t)b.n_bodies * BS, st) != cudaSuccess) { snprintf(h->err, sizeof(h->err), "fmpm_advect_rigid: memset failed"); return 1; }
  const int blocks = (P.N + 255) / 256;
  FMPM_LAUNCH(k_body_com, blocks, 256, 0, st, P, f, info, state, b.n_bodies);
  FMPM_LAUNCH(k_body_H, blocks, 256, 0, st, P, f, info, state, b.n_bodies);
  FMPM_LAUNCH(k_body_solve, (b.n_bodies + 31) / 32, 32, 0, st, info, state, b.n_bodies);
  FMPM_LAUNCH(k_body_advect, blocks, 256, 0, st, P, f, info, state, b.n_bodies);
  FMPM_CHECK_LAUNCH(h, "fmpm_advect_rigid");
  return 0;
}

extern "C" int fmpm_advect_rigid(FmpmHandle* h, int f, void* stream) {
  if (check(h, f, "fmpm_advect_rigid")) return 1;
  return fmpm_advect_rigid_impl(h, f, stream);
}

extern "C" int fmpm_advect_rigid_grad(FmpmHandle* h, int f, int gin, const void* next_slot, void* stream) {
  if (check(h, f, "fmpm_advect_rigid_grad")) return 1;
  const FmpmBodies& b = h->bodies;
  if (b.n_bodies == 0 || h->cfg.n_particles == 0) return 0;
  if (!b.grad || !h->buf.ga || (gin != 0 && gin != 1)) { snprintf(h->err, sizeof(h->err), "fmpm_advect_rigid_grad: adjoint buffers missing or bad gin %d", gin); return 1; }
  cudaStream_t st = (cudaStream_t)stream;
  KParams P = make_kparams(h);
  const float* state = (const float*)b.state + (size_t)f * b.n_bodies * BS;
  float* bg = (float*)b.grad;
  const int* info = (const int*)b.info;
  if (cudaMemsetAsync(bg, 0, sizeof(float) * (size_t)b.n_bodies * BG, st) != cudaSuccess) { snprintf(h->err, sizeof(h->err), "fmpm_advect_rigid_grad: memset failed"); return 1; }
  const int blocks = (P.N + 255) / 256;
  FMPM_LAUNCH(k_body_grad_reduce, blocks, 256, 0, st, P, f, gin, info, state, bg, b.n_bodies);
  FMPM_LAUNCH(k_body_grad_solve, (b.n_bodies + 31) / 32, 32, 0, st, info, state, bg, b.n_bodies);
  FMPM_LAUNCH(k_body_grad_apply, blocks, 256, 0, st, P, f, gin, (const int*)next_slot, info, state, bg, b.n_bodies);
  FMPM_CHECK_LAUNCH(h, "fmpm_advect_rigid_grad");
  return 0;
}
