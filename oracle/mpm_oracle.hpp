// =====================================================================================
// oracle/mpm_oracle.hpp — TEST INFRASTRUCTURE ONLY (never linked/imported by the product).
//
// CPU restatement of the reference MLS-MPM substep (forward + reverse-mode adjoint) of
// zhouxian/FluidLab, file fluidlab/fluidengine/simulators/mpm_simulator.py (abbrev. MPM).
// Every function cites the reference lines it follows.
//
// PARITY STATUS.  Taichi (taichi==1.1.0, environment.yml:250) is not installable here and the reference ships no tests or golden
// vectors.  FORWARD: pinned to the reference itself — its unmodified kernel source runs on a NumPy emulation of the Taichi API in the
// build container (tests/golden/taichi_emu.py, make_reference_run.py) and tests/test_reference_run.py requires this oracle to reproduce
// those runs (all material classes, both boundaries, MAT_RIGID bodies, injectors, collectors, 6-DOF pose chain, Static / Dynamic SDF
// colliders at grid and particle level).  STILL UNPINNED against Taichi: the internals of `ti.svd` (convention assumed: U, V proper
// rotations, singular values sorted descending, sign carried by the smallest; the emulation assumes the same) and `kernel.grad`
// (Taichi's autodiff: branch conditions and int casts carry no gradient, min/max send the adjoint to the selected operand).  The
// hand-written adjoints are validated against central finite differences THROUGH THE REFERENCE'S OWN FORWARD KERNELS run in float64
// (tests/golden/make_reference_fd.py, tests/test_reference_run.py), against central differences of this oracle (tests/test_oracle.py)
// and against torch.autograd applied to an independent PyTorch restatement (tests/test_torch_autodiff_crosscheck.py).
//
// Templated on the scalar so the same code gives the fp32 arithmetic of the reference
// (DTYPE_TI = f32, configs/macros.py:207-211) and an fp64 ground truth.
// =====================================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

// material classes, configs/macros.py:37-41
enum { MAT_LIQUID = 200, MAT_PLASTO_ELASTIC = 201, MAT_ELASTIC = 202, MAT_RIGID = 203,
       MAT_PLASTO_ELASTIC_DEMO = 204 };

struct Config {
  int n_grid, n_particles, T, n_substeps;
  double dt, dx, inv_dx, p_vol;
  double gravity[3];
  int boundary_type;  // 0 cube (boundaries.py:95-134), 1 cylinder (boundaries.py:26-92)
  double b_lower[3], b_upper[3];  // cube lower/upper; cylinder uses [1] as y range
  double cyl_center[2], cyl_radius;
  double restitution;
  int lock_mask;  // bit d set -> v[d] = 0 (lock_dims)
};

template <class R> struct V3 {
  R a[3];
  R& operator[](int i) { return a[i]; }
  const R& operator[](int i) const { return a[i]; }
};
template <class R> struct M3 {
  R a[3][3];
  R* operator[](int i) { return a[i]; }
  const R* operator[](int i) const { return a[i]; }
};
template <class R> static inline M3<R> zero3() { M3<R> m; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = R(0); return m; }
template <class R> static inline M3<R> ident3() { M3<R> m = zero3<R>(); m[0][0] = m[1][1] = m[2][2] = R(1); return m; }
template <class R> static inline M3<R> mul(const M3<R>& A, const M3<R>& B) {
  M3<R> C; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { R s = 0; for (int k = 0; k < 3; k++) s += A[i][k] * B[k][j]; C[i][j] = s; } return C; }
template <class R> static inline M3<R> tr(const M3<R>& A) { M3<R> C; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[i][j] = A[j][i]; return C; }
template <class R> static inline M3<R> add(const M3<R>& A, const M3<R>& B) { M3<R> C; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[i][j] = A[i][j] + B[i][j]; return C; }
template <class R> static inline M3<R> sub(const M3<R>& A, const M3<R>& B) { M3<R> C; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[i][j] = A[i][j] - B[i][j]; return C; }
template <class R> static inline M3<R> scale(const M3<R>& A, R s) { M3<R> C; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C[i][j] = A[i][j] * s; return C; }
template <class R> static inline R det3(const M3<R>& A) {
  return A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
         A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
}
template <class R> static inline R trace3(const M3<R>& A) { return A[0][0] + A[1][1] + A[2][2]; }

// -------------------------------------------------------------------------------------
// 3x3 SVD with the convention of ti.svd (call sites MPM:264, MPM:483): A = U diag(s) V^T,
// det U = det V = +1, |s0| >= |s1| >= |s2|, s2 carries the sign of det A.
// One-sided (Hestenes) Jacobi: rotate column pairs of B = A V until orthogonal.
// -------------------------------------------------------------------------------------
template <class R> static void svd3(const M3<R>& A, M3<R>& U, R sig[3], M3<R>& V) {
  M3<R> B = A;
  V = ident3<R>();
  const R eps = std::is_same<R, float>::value ? R(1e-7) : R(1e-15);
  const int max_sweeps = 40;
  for (int sweep = 0; sweep < max_sweeps; sweep++) {
    bool rotated = false;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        R alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 3; k++) { alpha += B[k][p] * B[k][p]; beta += B[k][q] * B[k][q]; gamma += B[k][p] * B[k][q]; }
        if (gamma == R(0) || std::fabs(gamma) <= eps * std::sqrt(alpha * beta)) continue;
        rotated = true;
        R zeta = (beta - alpha) / (R(2) * gamma);
        R t = (zeta >= 0 ? R(1) : R(-1)) / (std::fabs(zeta) + std::sqrt(R(1) + zeta * zeta));
        R c = R(1) / std::sqrt(R(1) + t * t), s = c * t;
        for (int k = 0; k < 3; k++) {
          R bp = B[k][p], bq = B[k][q];
          B[k][p] = c * bp - s * bq; B[k][q] = s * bp + c * bq;
          R vp = V[k][p], vq = V[k][q];
          V[k][p] = c * vp - s * vq; V[k][q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  R n[3];
  for (int j = 0; j < 3; j++) n[j] = std::sqrt(B[0][j] * B[0][j] + B[1][j] * B[1][j] + B[2][j] * B[2][j]);
  // sort columns descending by norm (selection sort; each swap flips det of both B-basis and V)
  int swaps = 0;
  for (int i = 0; i < 2; i++) {
    int m = i;
    for (int j = i + 1; j < 3; j++) if (n[j] > n[m]) m = j;
    if (m != i) {
      std::swap(n[i], n[m]);
      for (int k = 0; k < 3; k++) { std::swap(B[k][i], B[k][m]); std::swap(V[k][i], V[k][m]); }
      swaps++;
    }
  }
  if (swaps & 1) for (int k = 0; k < 3; k++) { B[k][2] = -B[k][2]; V[k][2] = -V[k][2]; }  // keeps B V^T, restores det V = +1
  // U columns
  const R tiny = std::is_same<R, float>::value ? R(1e-30) : R(1e-280);
  for (int j = 0; j < 3; j++) {
    sig[j] = n[j];
    if (n[j] > tiny) for (int k = 0; k < 3; k++) U[k][j] = B[k][j] / n[j];
    else for (int k = 0; k < 3; k++) U[k][j] = R(0);
  }
  // complete degenerate columns to an orthonormal basis
  if (!(n[0] > tiny)) { U = ident3<R>(); }
  else {
    if (!(n[1] > tiny)) {  // pick any unit vector orthogonal to u0
      int m = 0; for (int k = 1; k < 3; k++) if (std::fabs(U[k][0]) < std::fabs(U[m][0])) m = k;
      R e[3] = {0, 0, 0}; e[m] = 1;
      R d = U[m][0];
      R w[3]; R nn = 0; for (int k = 0; k < 3; k++) { w[k] = e[k] - d * U[k][0]; nn += w[k] * w[k]; }
      nn = std::sqrt(nn); for (int k = 0; k < 3; k++) U[k][1] = w[k] / nn;
    }
    if (!(n[2] > tiny)) {
      U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
      U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
      U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    }
  }
  if (det3(U) < 0) { for (int k = 0; k < 3; k++) U[k][2] = -U[k][2]; sig[2] = -sig[2]; }
}

// MPM:294-302 clamp: |a| >= 1e-8 keeping the sign (a == 0 counts as positive)
template <class R> static inline R clamp_svd(R a) { if (a >= 0) a = std::max(a, R(1e-8)); else a = std::min(a, R(-1e-8)); return a; }

// MPM:272-292 backward_svd (the old PyTorch svd_backward formula), S given as its diagonal,
// gS as a full 3x3 (only the diagonal is ever non-zero for the call site MPM:270).
template <class R> static M3<R> backward_svd(const M3<R>& gU, const M3<R>& gS, const M3<R>& gV, const M3<R>& U, const R sig[3], const M3<R>& V) {
  M3<R> vt = tr(V), ut = tr(U);
  M3<R> S = zero3<R>(); for (int d = 0; d < 3; d++) S[d][d] = sig[d];
  M3<R> S_term = mul(U, mul(gS, vt));
  R s2[3] = {sig[0] * sig[0], sig[1] * sig[1], sig[2] * sig[2]};
  M3<R> Fm;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Fm[i][j] = (i == j) ? R(0) : R(1) / clamp_svd<R>(s2[j] - s2[i]);
  M3<R> a = sub(mul(ut, gU), mul(tr(gU), U));
  M3<R> b = sub(mul(vt, gV), mul(tr(gV), V));
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { a[i][j] *= Fm[i][j]; b[i][j] *= Fm[i][j]; }
  M3<R> u_term = mul(U, mul(mul(a, S), vt));
  M3<R> v_term = mul(U, mul(S, mul(b, vt)));
  return add(add(u_term, v_term), S_term);
}

// -------------------------------------------------------------------------------------
// Effector pose chain (effectors/effector.py) — only what carries dLoss/dAction.
// -------------------------------------------------------------------------------------
struct EffectorCfg {
  int type;         // 0 plain, 1 Injector (injector.py:80-105), 2 BallInjector (injector.py:240-256)
  int action_dim;   // 0, 3 or 6
  double scale_v[6], scale_p[6];
  // own boundary (effector.py:63 setup_boundary)
  int boundary_type; double b_lower[3], b_upper[3]; double cyl_center[2], cyl_radius;
  // injector
  double radius; int flux; double inject_v[3], inject_p[3]; int locally_random; int random_length; int randomize_inject_v;
  int max_action_steps;
};

template <class R> struct Effector {
  EffectorCfg cfg;
  int T;
  std::vector<R> pos, quat, v, w, gpos, gquat, gv, gw;  // [(T+1)*3|4]
  std::vector<R> act, gact;                             // [max_action_steps * action_dim]
  std::vector<R> act_p, gact_p;                         // [action_dim]
  std::vector<R> random_vector;                         // [random_length * flux * 3]
  std::vector<int> act_range, act_id;                   // act_id[(T+1)]
  void init(const EffectorCfg& c, int T_) {
    cfg = c; T = T_;
    pos.assign((T + 1) * 3, 0); quat.assign((T + 1) * 4, 0); v.assign((T + 1) * 3, 0); w.assign((T + 1) * 3, 0);
    gpos = pos; gquat = quat; gv = v; gw = w;
    int ad = std::max(cfg.action_dim, 1);
    act.assign((size_t)cfg.max_action_steps * ad, 0); gact = act; act_p.assign(ad, 0); gact_p = act_p;
    act_id.assign(T + 1, 0);
  }
};


// -------------------------------------------------------------------------------------
// SDF mesh colliders: meshes/static.py:26-104 (Static) and meshes/dynamic.py:29-121 (Dynamic).
// The baked volume is `voxels[res^3]` (utils/mesh.py:63-87) with T_mesh_to_voxels already
// multiplied by inv(T_init) (meshes/mesh.py:121-127).
// -------------------------------------------------------------------------------------
template <class R> struct SdfMesh {
  int res = 0;
  std::vector<R> vox;
  R T[16];       // T_mesh_to_voxels (row-major 4x4)
  R Ainv[9];     // inverse of T[:3,:3]  (R_voxels_to_mesh)
  R friction = 0, softness = 0;
  int has_dynamics = 1;
  // sdf_ : trilinear lookup, 1.0 outside the volume (static.py:35-48).  grad (optional) = d sdf / d pos_voxels.
  R sdf_(const R* pv, R* grad) const {
    int b[3]; for (int d = 0; d < 3; d++) b[d] = (int)std::floor(pv[d]);
    if (grad) grad[0] = grad[1] = grad[2] = 0;
    for (int d = 0; d < 3; d++) if (b[d] >= res - 1 || b[d] < 0) return R(1);
    R sd = 0;
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
      int vp[3] = {b[0] + i, b[1] + j, b[2] + k};
      R w[3], dwd[3];
      for (int d = 0; d < 3; d++) { R t = pv[d] - (R)vp[d]; w[d] = R(1) - std::fabs(t); dwd[d] = t > 0 ? R(-1) : (t < 0 ? R(1) : R(0)); }
      R val = vox[((size_t)vp[0] * res + vp[1]) * res + vp[2]];
      sd += w[0] * w[1] * w[2] * val;
      if (grad) { grad[0] += dwd[0] * w[1] * w[2] * val; grad[1] += w[0] * dwd[1] * w[2] * val; grad[2] += w[0] * w[1] * dwd[2] * val; }
    }
    return sd;
  }
  void to_voxels(const R* pm, R* pv) const { for (int r = 0; r < 3; r++) pv[r] = T[r * 4] * pm[0] + T[r * 4 + 1] * pm[1] + T[r * 4 + 2] * pm[2] + T[r * 4 + 3]; }
  // normal_ in voxel space: central differences with delta = 1e-2 voxel, normalised with eps (static.py:66-79)
  void normal_vox(const R* pv, R* nv, R* graw, R& gn) const {
    const R delta = R(1e-2);
    for (int i = 0; i < 3; i++) {
      R inc[3] = {pv[0], pv[1], pv[2]}, dec[3] = {pv[0], pv[1], pv[2]};
      inc[i] += delta; dec[i] -= delta;
      graw[i] = (sdf_(inc, nullptr) - sdf_(dec, nullptr)) / (R(2) * delta);
    }
    gn = std::sqrt(graw[0] * graw[0] + graw[1] * graw[1] + graw[2] * graw[2] + R(1e-12));
    for (int i = 0; i < 3; i++) nv[i] = graw[i] / gn;
  }
};

template <class R> static inline void quat_rot_t(const R* q, const R* v, R* o) {  // utils/geom.py:92-97
  R uv[3] = {q[2] * v[2] - q[3] * v[1], q[3] * v[0] - q[1] * v[2], q[1] * v[1] - q[2] * v[0]};
  R uuv[3] = {q[2] * uv[2] - q[3] * uv[1], q[3] * uv[0] - q[1] * uv[2], q[1] * uv[1] - q[2] * uv[0]};
  for (int k = 0; k < 3; k++) o[k] = v[k] + R(2) * (q[0] * uv[k] + uuv[k]);
}
template <class R> static inline void quat_inv_t(const R* q, R* qi) {  // utils/geom.py:30-32 (normalised conjugate)
  R n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  qi[0] = q[0] / n; qi[1] = -q[1] / n; qi[2] = -q[2] / n; qi[3] = -q[3] / n;
}

// adjoint of o = quat_rot_t(q, v) (the polynomial v + 2 (q0 (qv x v) + qv x (qv x v)), differentiated as written): accumulates into gq[4], gv[3]
template <class R> static inline void quat_rot_adj(const R* q, const R* v, const R* go, R* gq, R* gv) {
  const R uv[3] = {q[2] * v[2] - q[3] * v[1], q[3] * v[0] - q[1] * v[2], q[1] * v[1] - q[2] * v[0]};
  // g_uv = 2 q0 go + 2 (go x qv)
  const R guv[3] = {R(2) * (q[0] * go[0] + go[1] * q[3] - go[2] * q[2]), R(2) * (q[0] * go[1] + go[2] * q[1] - go[0] * q[3]),
                    R(2) * (q[0] * go[2] + go[0] * q[2] - go[1] * q[1])};
  if (gq) {
    gq[0] += R(2) * (go[0] * uv[0] + go[1] * uv[1] + go[2] * uv[2]);
    // g_qv = 2 (uv x go) + v x g_uv
    gq[1] += R(2) * (uv[1] * go[2] - uv[2] * go[1]) + (v[1] * guv[2] - v[2] * guv[1]);
    gq[2] += R(2) * (uv[2] * go[0] - uv[0] * go[2]) + (v[2] * guv[0] - v[0] * guv[2]);
    gq[3] += R(2) * (uv[0] * go[1] - uv[1] * go[0]) + (v[0] * guv[1] - v[1] * guv[0]);
  }
  if (gv) {  // gv = go + g_uv x qv   (uv = qv x v)
    gv[0] += go[0] + (guv[1] * q[3] - guv[2] * q[2]);
    gv[1] += go[1] + (guv[2] * q[1] - guv[0] * q[3]);
    gv[2] += go[2] + (guv[0] * q[2] - guv[1] * q[1]);
  }
}
// adjoint of qi = quat_inv_t(q): accumulates into gq
template <class R> static inline void quat_inv_adj(const R* q, const R* qi, const R* gqi, R* gq) {
  const R n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const R d = qi[0] * gqi[0] + qi[1] * gqi[1] + qi[2] * gqi[2] + qi[3] * gqi[3];
  const R sgn[4] = {R(1), R(-1), R(-1), R(-1)};
  for (int k = 0; k < 4; k++) gq[k] += sgn[k] * (gqi[k] - qi[k] * d) / n;
}

// One collide evaluation (static: pos0 = pos1 = 0, quats identity, dynamic = false).  Forward value in `out`; when
// gout != nullptr the adjoints of (v, p, pos0, pos1) are ACCUMULATED into gv, gp, gpos0, gpos1, and those of the pose quaternions
// into gq0 / gq1 when these are given (6-DOF Rigid effectors: agent_pouring.yaml; quat is constant for action_dim = 3).
// Follows meshes/dynamic.py:93-121 (Dynamic.collide) / meshes/static.py:82-104 (Static.collide).
template <class R> static void sdf_collide(const SdfMesh<R>& M, bool dynamic, const R* pos0, const R* q0, const R* pos1, const R* q1, R dt,
                                           const R* p, const R* v, R* out, const R* gout, R* gv, R* gp, R* gpos0, R* gpos1,
                                           R* gq0 = nullptr, R* gq1 = nullptr) {
  for (int k = 0; k < 3; k++) out[k] = v[k];
  if (!M.has_dynamics) { if (gout) for (int k = 0; k < 3; k++) gv[k] += gout[k]; return; }
  R qi[4], d0[3], pm[3], pv[3];
  quat_inv_t(q0, qi);
  for (int k = 0; k < 3; k++) d0[k] = p[k] - pos0[k];
  quat_rot_t(qi, d0, pm);
  M.to_voxels(pm, pv);
  R gsd[3];
  const R sd = M.sdf_(pv, gsd);
  const R infl = dynamic ? std::min((R)std::exp(-sd * M.softness), R(1)) : R(1);
  const bool hit = dynamic ? (sd <= 0 || (M.softness > 0 && infl > R(0.1))) : (sd <= 0);
  if (!hit) { if (gout) for (int k = 0; k < 3; k++) gv[k] += gout[k]; return; }
  R cv[3] = {0, 0, 0};
  if (dynamic) {
    R pw1[3]; quat_rot_t(q1, pm, pw1);
    for (int k = 0; k < 3; k++) cv[k] = (pw1[k] + pos1[k] - p[k]) / dt;   // collider_v, dynamic.py:86-91
  }
  const bool sticky = dynamic && (M.friction > R(10));
  R rel[3], nvx[3], graw[3], gnorm = 1, u[3], un = 1, n[3] = {0, 0, 0}, vn = 0, m = 0, rt[3], rtn = 0, g = 0, rt2[3], nm[3] = {0, 0, 0};
  bool flag = false;
  if (sticky) { for (int k = 0; k < 3; k++) out[k] = cv[k]; }
  else {
    for (int k = 0; k < 3; k++) rel[k] = v[k] - cv[k];
    M.normal_vox(pv, nvx, graw, gnorm);
    for (int r = 0; r < 3; r++) nm[r] = M.Ainv[r * 3] * nvx[0] + M.Ainv[r * 3 + 1] * nvx[1] + M.Ainv[r * 3 + 2] * nvx[2];
    quat_rot_t(q0, nm, u);
    un = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2] + R(1e-12));
    for (int k = 0; k < 3; k++) n[k] = u[k] / un;
    vn = rel[0] * n[0] + rel[1] * n[1] + rel[2] * n[2];
    m = std::min(vn, R(0));
    for (int k = 0; k < 3; k++) rt[k] = rel[k] - m * n[k];
    rtn = std::sqrt(rt[0] * rt[0] + rt[1] * rt[1] + rt[2] * rt[2]);
    g = std::max(R(0), rtn + vn * M.friction);
    flag = (vn < 0) && (rtn > R(1e-12));
    for (int k = 0; k < 3; k++) rt2[k] = flag ? rt[k] / rtn * g : rt[k];
    for (int k = 0; k < 3; k++) out[k] = cv[k] + rt2[k] * infl + rel[k] * (R(1) - infl);
  }
  if (!gout) return;
  // ------------------------------------------------------------------ adjoint
  R gcv[3] = {gout[0], gout[1], gout[2]}, gpv[3] = {0, 0, 0}, gsdv = 0;
  if (!sticky) {
    R grt2[3], grel[3], ginfl = 0;
    for (int k = 0; k < 3; k++) { grt2[k] = infl * gout[k]; ginfl += gout[k] * (rt2[k] - rel[k]); grel[k] = (R(1) - infl) * gout[k]; }
    R grt[3] = {0, 0, 0}, gvn = 0, gn[3] = {0, 0, 0};
    if (flag) {
      const R sc = g / rtn;
      R sbar = 0;
      for (int k = 0; k < 3; k++) { grt[k] += sc * grt2[k]; sbar += rt[k] * grt2[k]; }
      R grtn = -sbar * g / (rtn * rtn);
      if (rtn + vn * M.friction > 0) { grtn += sbar / rtn; gvn += sbar / rtn * M.friction; }   // max(0, a): adjoint to a iff 0 < a
      for (int k = 0; k < 3; k++) grt[k] += grtn * rt[k] / rtn;
    } else {
      for (int k = 0; k < 3; k++) grt[k] += grt2[k];
    }
    R gm = 0;
    for (int k = 0; k < 3; k++) { grel[k] += grt[k]; gm -= grt[k] * n[k]; gn[k] -= m * grt[k]; }
    if (vn < 0) gvn += gm;                                                                       // min(vn, 0): adjoint to vn iff vn < 0
    for (int k = 0; k < 3; k++) { grel[k] += gvn * n[k]; gn[k] += gvn * rel[k]; }
    for (int k = 0; k < 3; k++) { gv[k] += grel[k]; gcv[k] -= grel[k]; }
    if (dynamic) {
      // n = u / |u|_eps, u = R0 nm, nm = Ainv nvx, nvx = graw / |graw|_eps, graw = FD gradient of sdf_ at pv
      R nd = n[0] * gn[0] + n[1] * gn[1] + n[2] * gn[2];
      R gu[3]; for (int k = 0; k < 3; k++) gu[k] = (gn[k] - n[k] * nd) / un;
      R gnm[3]; quat_rot_t(qi, gu, gnm);
      if (gq0) quat_rot_adj(q0, nm, gu, gq0, (R*)nullptr);
      R gnvx[3]; for (int c = 0; c < 3; c++) gnvx[c] = M.Ainv[c] * gnm[0] + M.Ainv[3 + c] * gnm[1] + M.Ainv[6 + c] * gnm[2];
      R nd2 = nvx[0] * gnvx[0] + nvx[1] * gnvx[1] + nvx[2] * gnvx[2];
      R ggraw[3]; for (int k = 0; k < 3; k++) ggraw[k] = (gnvx[k] - nvx[k] * nd2) / gnorm;
      const R delta = R(1e-2);
      for (int i = 0; i < 3; i++) {
        R inc[3] = {pv[0], pv[1], pv[2]}, dec[3] = {pv[0], pv[1], pv[2]}, gi[3], gd[3];
        inc[i] += delta; dec[i] -= delta;
        M.sdf_(inc, gi); M.sdf_(dec, gd);
        for (int k = 0; k < 3; k++) gpv[k] += ggraw[i] * (gi[k] - gd[k]) / (R(2) * delta);
      }
      if (std::exp(-sd * M.softness) < R(1)) gsdv += ginfl * (-M.softness) * infl;               // min(e, 1): adjoint to e iff e < 1
    }
  }
  if (dynamic) {
    for (int k = 0; k < 3; k++) gpv[k] += gsdv * gsd[k];
    // cv = (R1 pm + pos1 - p) / dt
    R q1i[4]; quat_inv_t(q1, q1i);
    R t1[3] = {gcv[0] / dt, gcv[1] / dt, gcv[2] / dt}, gpm[3];
    quat_rot_t(q1i, t1, gpm);
    if (gq1) quat_rot_adj(q1, pm, t1, gq1, (R*)nullptr);                                               // pw1 = R1 pm
    for (int k = 0; k < 3; k++) { gpos1[k] += t1[k]; gp[k] -= t1[k]; }
    for (int c = 0; c < 3; c++) gpm[c] += M.T[c] * gpv[0] + M.T[4 + c] * gpv[1] + M.T[8 + c] * gpv[2];   // pv = A pm + t
    R gd0[3]; quat_rot_t(q0, gpm, gd0);                                                                // pm = R0^T (p - pos0)
    if (gq0) { R gqi[4] = {0, 0, 0, 0}; quat_rot_adj(qi, d0, gpm, gqi, (R*)nullptr); quat_inv_adj(q0, qi, gqi, gq0); }
    for (int k = 0; k < 3; k++) { gp[k] += gd0[k]; gpos0[k] -= gd0[k]; }
  }
}

template <class R> struct Sim {
  Config c;
  int N, G, T;
  R dt, dx, inv_dx, p_vol, k_stress;  // Python-double constants cast once to R, like Taichi does
  // frame ring, MPM:75-88,106-107 (x,v,C,F,used) — F_tmp/U/S/V are recomputed scratch here
  std::vector<R> x, v, C, F, gx, gv, gC, gF;
  std::vector<int> used;
  // particle info, MPM:96-103
  std::vector<R> mu, lam, mass;
  std::vector<int> mat, cls;
  // scratch for the frame being processed
  std::vector<R> Ftmp, Usv, Ssv, Vsv;           // [N*9], [N*9], [N*3], [N*9]
  std::vector<R> g_vin, g_m, g_vout;            // grid, MPM:112-117
  std::vector<R> gg_vin, gg_m, gg_vout;         // grid grads
  // agent
  int agent_type = 0;     // 0 none, 1 AgentRigid (agents/agent_rigid.py), 2 AgentInjector (agents/agent_injector.py)
  int collide_type = 0;   // 0 particle, 1 grid, 2 both (agents/agent.py:17)
  int rigid_idx = 0, inj_idx = 0;   // effector indices (AgentIceCreamDynamic: injector 0, rigid 1)
  int inject_till = 1 << 30;        // agents/agent_icecreamdynamic.py:24-27
  std::vector<Effector<R>> eff;
  std::vector<SdfMesh<R>> statics;   // Statics (meshes/statics.py), collided in grid_op in order (MPM:388-390)
  SdfMesh<R> rigid_mesh;             // mesh of the single Rigid effector (effectors/rigid.py:21-26)
  bool has_rigid = false;
  bool has_injector = false;
  R collide_y_min = R(-1e30);        // AgentIceCreamDynamic.collide only acts above y = 0.25 (agents/agent_icecreamdynamic.py:38-43)
  // collector of AgentPouring / AgentJetBot (agents/agent_pouring.py:31-41, agents/agent_jetbot.py:30-40); type -1 = none
  int collector_type = -1;           // 0 cube, 1 cylinder (boundaries/boundaries.py:81-93,128-134 is_out)
  R col_lo[3] = {0, 0, 0}, col_hi[3] = {1, 1, 1}, col_c[2] = {R(0.5), R(0.5)}, col_r = R(1);
  int collector_mat = -1;            // -1: every material (AgentPouring); otherwise only that material (AgentJetBot: WATER)
  R nowhere[3] = {R(-100), R(-100), R(-100)};   // configs/macros.py:216
  inline bool collector_is_out(const R* xp) const {
    bool out = false;
    if (collector_type == 0) { for (int k = 0; k < 3; k++) if (xp[k] > col_hi[k] || xp[k] < col_lo[k]) out = true; }
    else {
      if (xp[1] > col_hi[1] || xp[1] < col_lo[1]) out = true;
      R rx = xp[0] - col_c[0], rz = xp[2] - col_c[1];
      if (std::sqrt(rx * rx + rz * rz + R(1e-12)) > col_r) out = true;
    }
    return out;
  }
  void collector_act(int f) {   // collect out-of-boundary particles: they leave the simulation at frame f already
    if (collector_type < 0) return;
    for (int p = 0; p < N; p++) if (used[pi(f, p)] && (collector_mat < 0 || mat[p] == collector_mat)) {
      if (collector_is_out(&x[pi(f, p) * 3])) {
        used[pi(f + 1, p)] = 0;
        for (int k = 0; k < 3; k++) x[pi(f + 1, p) * 3 + k] = nowhere[k];
        used[pi(f, p)] = 0;
      }
    }
  }
  // bodies (MPM:177-201): rigidity enforcement by shape matching for MAT_RIGID bodies
  int n_bodies = 0;
  std::vector<int> body_id, body_n, body_cls;
  struct BodyState { R c0[3], c1[3]; M3<R> H, U, V, Rm; R S[3]; };
  std::vector<BodyState> bodies;
  bool any_rigid_body() const { for (int b = 0; b < n_bodies; b++) if (body_cls[b] == MAT_RIGID) return true; return false; }
  void set_bodies(const int* bid, int nb) {  // MPM:177-201 init_bodies: n_particles counts every slot of the body, mat_cls = its first particle's
    n_bodies = nb; body_id.assign(bid, bid + N); body_n.assign(nb, 0); body_cls.assign(nb, 0); bodies.resize(nb);
    std::vector<char> seen(nb, 0);
    for (int p = 0; p < N; p++) { int b = bid[p]; body_n[b]++; if (!seen[b]) { seen[b] = 1; body_cls[b] = cls[p]; } }
  }
  inline bool rigid_p(int f, int p) const { return n_bodies > 0 && used[pi(f, p)] && cls[p] == MAT_RIGID; }

  // agent.collide(f, pos, v, dt) for AgentRigid -> Rigid.collide -> Dynamic.collide; adjoint accumulates into effector gpos
  inline void agent_collide(int f, const R* p, const R* vin, R* out, const R* gout, R* gvv, R* gpp) {
    if (agent_type != 1 || !has_rigid || !(p[1] > collide_y_min)) { for (int k = 0; k < 3; k++) out[k] = vin[k]; if (gout) for (int k = 0; k < 3; k++) gvv[k] += gout[k]; return; }
    Effector<R>& e = eff[rigid_idx];
    R g0[3] = {0, 0, 0}, g1[3] = {0, 0, 0}, gq0[4] = {0, 0, 0, 0}, gq1[4] = {0, 0, 0, 0};
    sdf_collide(rigid_mesh, true, &e.pos[f * 3], &e.quat[f * 4], &e.pos[(f + 1) * 3], &e.quat[(f + 1) * 4], dt, p, vin, out, gout, gvv, gpp, g0, g1, gq0, gq1);
    if (gout) {
      for (int k = 0; k < 3; k++) {
#pragma omp atomic
        e.gpos[f * 3 + k] += g0[k];
#pragma omp atomic
        e.gpos[(f + 1) * 3 + k] += g1[k];
      }
      for (int k = 0; k < 4; k++) {
#pragma omp atomic
        e.gquat[f * 4 + k] += gq0[k];
#pragma omp atomic
        e.gquat[(f + 1) * 4 + k] += gq1[k];
      }
    }
  }

  explicit Sim(const Config& cfg) : c(cfg) {
    N = c.n_particles; T = c.T; G = c.n_grid * c.n_grid * c.n_grid;
    dt = (R)c.dt; dx = (R)c.dx; inv_dx = (R)c.inv_dx; p_vol = (R)c.p_vol;
    k_stress = (R)(-c.dt * c.p_vol * 4 * c.inv_dx * c.inv_dx);  // MPM:343
    size_t F1 = (size_t)(T + 1) * N;
    x.assign(F1 * 3, 0); v.assign(F1 * 3, 0); C.assign(F1 * 9, 0); F.assign(F1 * 9, 0);
    gx = x; gv = v; gC = C; gF = F;
    used.assign(F1, 0);
    mu.assign(N, 0); lam.assign(N, 0); mass.assign(N, 0); mat.assign(N, 0); cls.assign(N, MAT_LIQUID);
    Ftmp.assign((size_t)N * 9, 0); Usv = Ftmp; Vsv = Ftmp; Ssv.assign((size_t)N * 3, 0);
    g_vin.assign((size_t)G * 3, 0); g_m.assign(G, 0); g_vout.assign((size_t)G * 3, 0);
    gg_vin = g_vin; gg_m = g_m; gg_vout = g_vout;
  }

  inline size_t pi(int f, int p) const { return (size_t)f * N + p; }
  inline M3<R> getM(const std::vector<R>& a, size_t i) const { M3<R> m; std::memcpy(&m, &a[i * 9], sizeof(R) * 9); return m; }
  inline void setM(std::vector<R>& a, size_t i, const M3<R>& m) { std::memcpy(&a[i * 9], &m, sizeof(R) * 9); }
  inline void addM(std::vector<R>& a, size_t i, const M3<R>& m) { for (int k = 0; k < 9; k++) a[i * 9 + k] += (&m.a[0][0])[k]; }

  // ---------------------------------------------------------------- boundaries
  // CubeBoundary.impose_x_v (boundaries.py:106-120) / CylinderBoundary.impose_x_v (:39-63); returns v only.
  // flags (optional) records the factor applied per component for the adjoint.
  inline void boundary_v(const R pos[3], R vv[3], R fac[3]) const {
    fac[0] = fac[1] = fac[2] = R(1);
    const R rest = (R)c.restitution;
    if (c.boundary_type == 0) {
      for (int i = 0; i < 3; i++) {
        if (pos[i] >= (R)c.b_upper[i] && vv[i] >= 0) fac[i] *= -rest;
        else if (pos[i] <= (R)c.b_lower[i] && vv[i] <= 0) fac[i] *= -rest;
      }
    } else {
      if (pos[1] > (R)c.b_upper[1] && vv[1] > R(0)) fac[1] *= -rest;
      else if (pos[1] < (R)c.b_lower[1] && vv[1] < R(0)) fac[1] *= -rest;
      R rx = pos[0] - (R)c.cyl_center[0], rz = pos[2] - (R)c.cyl_center[1];
      R rn = std::sqrt(rx * rx + rz * rz + R(1e-12));  // norm(EPS), macros.py:213
      if (rn > (R)c.cyl_radius) { fac[0] = 0; fac[2] = 0; }
    }
    for (int i = 0; i < 3; i++) if (c.lock_mask & (1 << i)) fac[i] = 0;
    for (int i = 0; i < 3; i++) vv[i] = (fac[i] == R(0)) ? R(0) : vv[i] * fac[i];
  }

  // Boundary.impose_x for the effector's own boundary (boundaries.py:65-78,122-125).
  // jac (3x3, row-major d out / d in) is filled for the adjoint using Taichi's min/max tie rule.
  static inline void effector_impose_x(const EffectorCfg& e, const R in[3], R out[3], R jac[9]) {
    for (int k = 0; k < 9; k++) jac[k] = 0;
    R lo[3], hi[3];
    if (e.boundary_type == 0) { for (int i = 0; i < 3; i++) { lo[i] = (R)e.b_lower[i]; hi[i] = (R)e.b_upper[i]; } }
    else { lo[0] = 0; hi[0] = 1; lo[2] = 0; hi[2] = 1; lo[1] = (R)e.b_lower[1]; hi[1] = (R)e.b_upper[1]; }
    R y[3]; R d[3];
    for (int i = 0; i < 3; i++) {
      R m = std::min(in[i], hi[i]); bool pass = in[i] < hi[i];  // min(a,b): grad to a iff a < b
      R mm = std::max(m, lo[i]); pass = pass && (lo[i] < m);     // max(a,b): grad to a iff b < a
      y[i] = mm; d[i] = pass ? R(1) : R(0);
    }
    out[0] = y[0]; out[1] = y[1]; out[2] = y[2];
    jac[0] = d[0]; jac[4] = d[1]; jac[8] = d[2];
    if (e.boundary_type == 1) {
      R rx = in[0] - (R)e.cyl_center[0], rz = in[2] - (R)e.cyl_center[1];
      R rn = std::sqrt(rx * rx + rz * rz + R(1e-12));
      if (rn > (R)e.cyl_radius) {
        R Rr = (R)e.cyl_radius;
        out[0] = rx / rn * Rr + (R)e.cyl_center[0];
        out[2] = rz / rn * Rr + (R)e.cyl_center[1];
        // d(r/|r| * R)/dr = R (I/|r| - r r^T/|r|^3)
        R i3 = R(1) / (rn * rn * rn);
        jac[0] = Rr * (R(1) / rn - rx * rx * i3); jac[2] = Rr * (-rx * rz * i3);
        jac[6] = Rr * (-rz * rx * i3);            jac[8] = Rr * (R(1) / rn - rz * rz * i3);
      }
    }
  }

  // ---------------------------------------------------------------- forward kernels
  void reset_grid() {  // MPM:219-223
    std::fill(g_vin.begin(), g_vin.end(), R(0)); std::fill(g_m.begin(), g_m.end(), R(0)); std::fill(g_vout.begin(), g_vout.end(), R(0));
    std::fill(gg_vin.begin(), gg_vin.end(), R(0)); std::fill(gg_m.begin(), gg_m.end(), R(0)); std::fill(gg_vout.begin(), gg_vout.end(), R(0));
  }
  void advect_used(int f) {  // MPM:304-307
    for (int p = 0; p < N; p++) used[pi(f + 1, p)] = used[pi(f, p)];
  }
  void process_unused(int f) {  // MPM:309-316
#pragma omp parallel for
    for (int p = 0; p < N; p++) if (used[pi(f, p)] == 0) {
      size_t a = pi(f, p), b = pi(f + 1, p);
      for (int k = 0; k < 3; k++) { v[b * 3 + k] = v[a * 3 + k]; x[b * 3 + k] = x[a * 3 + k]; }
      for (int k = 0; k < 9; k++) { C[b * 9 + k] = C[a * 9 + k]; F[b * 9 + k] = F[a * 9 + k]; }
    }
  }
  void compute_F_tmp_svd(int f) {  // MPM:254-264
#pragma omp parallel for
    for (int p = 0; p < N; p++) if (used[pi(f, p)]) {
      M3<R> Cm = getM(C, pi(f, p)), Fm = getM(F, pi(f, p));
      M3<R> A = add(ident3<R>(), scale(Cm, dt));
      M3<R> Ft = mul(A, Fm);
      M3<R> U, V; R s[3];
      svd3(Ft, U, s, V);
      setM(Ftmp, p, Ft); setM(Usv, p, U); setM(Vsv, p, V);
      Ssv[(size_t)p * 3] = s[0]; Ssv[(size_t)p * 3 + 1] = s[1]; Ssv[(size_t)p * 3 + 2] = s[2];
    }
  }
  static inline void weights(const R fx[3], R w[3][3]) {  // MPM:337
    for (int d = 0; d < 3; d++) {
      w[0][d] = R(0.5) * (R(1.5) - fx[d]) * (R(1.5) - fx[d]);
      w[1][d] = R(0.75) - (fx[d] - R(1)) * (fx[d] - R(1));
      w[2][d] = R(0.5) * (fx[d] - R(0.5)) * (fx[d] - R(0.5));
    }
  }
  static inline void dweights(const R fx[3], R dw[3][3]) {  // derivative of MPM:337
    for (int d = 0; d < 3; d++) {
      dw[0][d] = -(R(1.5) - fx[d]);
      dw[1][d] = R(-2) * (fx[d] - R(1));
      dw[2][d] = fx[d] - R(0.5);
    }
  }
  inline void base_fx(const R* xp, int base[3], R fx[3]) const {  // MPM:335-336 (cast(int) truncates toward zero)
    for (int d = 0; d < 3; d++) { base[d] = (int)(xp[d] * inv_dx - R(0.5)); fx[d] = xp[d] * inv_dx - (R)base[d]; }
  }
  inline size_t gidx(int i, int j, int k) const { return ((size_t)i * c.n_grid + j) * c.n_grid + k; }

  // stress/affine of MPM:339-344 from the scratch SVD
  inline M3<R> affine_of(int f, int p, R& J) const {
    M3<R> Ft = getM(Ftmp, p), U = getM(Usv, p), V = getM(Vsv, p);
    const R* s = &Ssv[(size_t)p * 3];
    J = s[0] * s[1] * s[2];  // determinant of the diagonal S
    M3<R> r = mul(U, tr(V));
    M3<R> stress = add(scale(mul(sub(Ft, r), tr(Ft)), R(2) * mu[p]), scale(ident3<R>(), lam[p] * J * (J - R(1))));
    stress = scale(stress, k_stress);
    return add(stress, scale(getM(C, pi(f, p)), mass[p]));
  }

  void p2g(int f, bool write_F) {  // MPM:331-378
#pragma omp parallel for
    for (int p = 0; p < N; p++) if (used[pi(f, p)]) {
      const R* xp = &x[pi(f, p) * 3];
      const R* vp = &v[pi(f, p) * 3];
      int base[3]; R fx[3]; base_fx(xp, base, fx);
      R w[3][3]; weights(fx, w);
      R J; M3<R> affine = affine_of(f, p, J);
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
        R dpos[3] = {((R)i - fx[0]) * dx, ((R)j - fx[1]) * dx, ((R)k - fx[2]) * dx};
        R weight = R(1); weight *= w[i][0]; weight *= w[j][1]; weight *= w[k][2];
        size_t g = gidx(base[0] + i, base[1] + j, base[2] + k);
        for (int a = 0; a < 3; a++) {
          R val = weight * (mass[p] * vp[a] + (affine[a][0] * dpos[0] + affine[a][1] * dpos[1] + affine[a][2] * dpos[2]));
#pragma omp atomic
          g_vin[g * 3 + a] += val;
        }
        R mval = weight * mass[p];
#pragma omp atomic
        g_m[g] += mval;
      }
      if (write_F) {
        M3<R> Fn = zero3<R>();
        if (cls[p] == MAT_LIQUID) Fn = scale(ident3<R>(), (R)std::pow(J, R(1.0 / 3)));            // MPM:358-359
        else if (cls[p] == MAT_ELASTIC || cls[p] == MAT_RIGID) Fn = getM(Ftmp, p);                  // MPM:361-365
        else if (cls[p] == MAT_PLASTO_ELASTIC || cls[p] == MAT_PLASTO_ELASTIC_DEMO) {              // MPM:367-376
          M3<R> Sn = zero3<R>();
          for (int d = 0; d < 3; d++) Sn[d][d] = std::min(std::max(Ssv[(size_t)p * 3 + d], R(1 - 2e-3)), R(1 + 3e-3));
          Fn = mul(getM(Usv, p), mul(Sn, tr(getM(Vsv, p))));
        }
        setM(F, pi(f + 1, p), Fn);
      }
    }
  }

  void grid_op(int f) {  // MPM:380-398 (no statics / grid-level agent collide in this oracle yet)
    const int n = c.n_grid;
#pragma omp parallel for
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) {
      size_t g = gidx(i, j, k);
      if (g_m[g] > R(1e-12)) {
        R inv_m = R(1) / g_m[g];
        R vv[3];
        for (int a = 0; a < 3; a++) { vv[a] = inv_m * g_vin[g * 3 + a]; vv[a] += dt * (R)c.gravity[a]; }
        R pos[3] = {(R)i * dx, (R)j * dx, (R)k * dx};
        const R zero3v[3] = {0, 0, 0}, idq[4] = {1, 0, 0, 0};
        for (size_t si = 0; si < statics.size(); si++) {   // MPM:388-390
          R o[3]; sdf_collide<R>(statics[si], false, zero3v, idq, zero3v, idq, dt, pos, vv, o, nullptr, nullptr, nullptr, nullptr, nullptr);
          for (int a = 0; a < 3; a++) vv[a] = o[a];
        }
        if (collide_type >= 1) { R o[3]; agent_collide(f, pos, vv, o, nullptr, nullptr, nullptr); for (int a = 0; a < 3; a++) vv[a] = o[a]; }  // MPM:393-395
        R fac[3];
        boundary_v(pos, vv, fac);
        for (int a = 0; a < 3; a++) g_vout[g * 3 + a] = vv[a];
      }
    }
  }

  void g2p(int f) {  // MPM:400-426 (+ advect_kernel MPM:497-505 for non-rigid particles)
#pragma omp parallel for
    for (int p = 0; p < N; p++) if (used[pi(f, p)]) {
      const R* xp = &x[pi(f, p) * 3];
      int base[3]; R fx[3]; base_fx(xp, base, fx);
      R w[3][3]; weights(fx, w);
      R nv[3] = {0, 0, 0}; M3<R> nC = zero3<R>();
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
        R dpos[3] = {(R)i - fx[0], (R)j - fx[1], (R)k - fx[2]};
        const R* gv_ = &g_vout[gidx(base[0] + i, base[1] + j, base[2] + k) * 3];
        R weight = R(1); weight *= w[i][0]; weight *= w[j][1]; weight *= w[k][2];
        for (int a = 0; a < 3; a++) {
          nv[a] += weight * gv_[a];
          for (int b = 0; b < 3; b++) nC[a][b] += R(4) * inv_dx * weight * gv_[a] * dpos[b];
        }
      }
      if (collide_type == 0 || collide_type == 2) {   // MPM:419-422
        R xt[3] = {xp[0] + dt * nv[0], xp[1] + dt * nv[1], xp[2] + dt * nv[2]}, o[3];
        agent_collide(f, xt, nv, o, nullptr, nullptr, nullptr);
        for (int a = 0; a < 3; a++) nv[a] = o[a];
      }
      size_t q = pi(f + 1, p);
      for (int a = 0; a < 3; a++) v[q * 3 + a] = nv[a];
      setM(C, q, nC);
    }
  }
  // MPM:449-495: reset_bodies_and_grad, compute_COM, compute_H, compute_H_svd, compute_R (serial sums: deterministic)
  void body_forward(int f) {
    for (int b = 0; b < n_bodies; b++) if (body_cls[b] == MAT_RIGID) {
      BodyState& B = bodies[b];
      for (int k = 0; k < 3; k++) B.c0[k] = B.c1[k] = R(0);
      B.H = zero3<R>();
    }
    for (int p = 0; p < N; p++) if (rigid_p(f, p)) {   // MPM:456-462
      BodyState& B = bodies[body_id[p]]; const R n = (R)body_n[body_id[p]];
      size_t a = pi(f, p), b1 = pi(f + 1, p);
      for (int k = 0; k < 3; k++) { B.c0[k] += x[a * 3 + k] / n; B.c1[k] += (x[a * 3 + k] + dt * v[b1 * 3 + k]) / n; }
    }
    for (int p = 0; p < N; p++) if (rigid_p(f, p)) {   // MPM:464-477
      BodyState& B = bodies[body_id[p]];
      size_t a = pi(f, p), b1 = pi(f + 1, p);
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        B.H[i][j] += (x[a * 3 + i] - B.c0[i]) * (x[a * 3 + j] + dt * v[b1 * 3 + j] - B.c1[j]);
    }
    for (int b = 0; b < n_bodies; b++) if (body_cls[b] == MAT_RIGID) {   // MPM:479-495
      BodyState& B = bodies[b];
      svd3(B.H, B.U, B.S, B.V);
      B.Rm = mul(B.V, tr(B.U));
    }
  }
  void advect(int f) {  // MPM:428-434,497-505
    const bool rb = any_rigid_body();
    if (rb) body_forward(f);
#pragma omp parallel for
    for (int p = 0; p < N; p++) if (used[pi(f, p)]) {
      size_t a = pi(f, p), b = pi(f + 1, p);
      if (rb && cls[p] == MAT_RIGID) {
        const BodyState& B = bodies[body_id[p]];
        for (int k = 0; k < 3; k++) {
          R acc = B.c1[k];
          for (int j = 0; j < 3; j++) acc += B.Rm[k][j] * (x[a * 3 + j] - B.c0[j]);
          x[b * 3 + k] = acc;
        }
      } else {
        for (int k = 0; k < 3; k++) x[b * 3 + k] = x[a * 3 + k] + dt * v[b * 3 + k];
      }
    }
  }
  // MPM:436-447 advect_grad for the MAT_RIGID particles: advect_kernel.grad, compute_R.grad, compute_H_svd_grad (manual, MPM:485-489),
  // compute_H.grad, compute_COM.grad.  Adds into gx[f] and gv[f+1]; gx[f+1] of rigid particles is consumed here.
  void body_advect_grad(int f) {
    body_forward(f);   // the reference recomputes the body state too (MPM:437-441)
    struct BG { M3<R> gR, gH; R gc0[3], gc1[3]; };
    std::vector<BG> g(n_bodies);
    for (auto& q : g) { q.gR = zero3<R>(); q.gH = zero3<R>(); for (int k = 0; k < 3; k++) q.gc0[k] = q.gc1[k] = R(0); }
    for (int p = 0; p < N; p++) if (rigid_p(f, p)) {   // advect_kernel.grad, rigid branch
      const BodyState& B = bodies[body_id[p]]; BG& q = g[body_id[p]];
      size_t a = pi(f, p), b1 = pi(f + 1, p);
      const R* go = &gx[b1 * 3];
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) q.gR[i][j] += go[i] * (x[a * 3 + j] - B.c0[j]);
      for (int j = 0; j < 3; j++) {
        R t = 0; for (int i = 0; i < 3; i++) t += B.Rm[i][j] * go[i];
        gx[a * 3 + j] += t; q.gc0[j] -= t; q.gc1[j] += go[j];
      }
    }
    for (int b = 0; b < n_bodies; b++) if (body_cls[b] == MAT_RIGID) {   // compute_R.grad + compute_H_svd_grad
      const BodyState& B = bodies[b];
      M3<R> gV = mul(g[b].gR, B.U), gU = mul(tr(g[b].gR), B.V);
      g[b].gH = backward_svd(gU, zero3<R>(), gV, B.U, B.S, B.V);
    }
    for (int p = 0; p < N; p++) if (rigid_p(f, p)) {   // compute_H.grad
      const BodyState& B = bodies[body_id[p]]; BG& q = g[body_id[p]];
      size_t a = pi(f, p), b1 = pi(f + 1, p);
      R d0[3], d1[3];
      for (int k = 0; k < 3; k++) { d0[k] = x[a * 3 + k] - B.c0[k]; d1[k] = x[a * 3 + k] + dt * v[b1 * 3 + k] - B.c1[k]; }
      for (int i = 0; i < 3; i++) {
        R t0 = 0, t1 = 0;
        for (int j = 0; j < 3; j++) { t0 += q.gH[i][j] * d1[j]; t1 += q.gH[j][i] * d0[j]; }
        gx[a * 3 + i] += t0 + t1; gv[b1 * 3 + i] += dt * t1;
        q.gc0[i] -= t0; q.gc1[i] -= t1;
      }
    }
    for (int p = 0; p < N; p++) if (rigid_p(f, p)) {   // compute_COM.grad
      const BG& q = g[body_id[p]]; const R n = (R)body_n[body_id[p]];
      size_t a = pi(f, p), b1 = pi(f + 1, p);
      for (int k = 0; k < 3; k++) { gx[a * 3 + k] += q.gc0[k] / n + q.gc1[k] / n; gv[b1 * 3 + k] += dt * q.gc1[k] / n; }
    }
  }

  // ---------------------------------------------------------------- agent
  static inline void quat_rotate(const R q[4], const R vv[3], R out[3]) {  // utils/geom.py:92-97
    R uv[3] = {q[2] * vv[2] - q[3] * vv[1], q[3] * vv[0] - q[1] * vv[2], q[1] * vv[1] - q[2] * vv[0]};
    R uuv[3] = {q[2] * uv[2] - q[3] * uv[1], q[3] * uv[0] - q[1] * uv[2], q[1] * uv[1] - q[2] * uv[0]};
    for (int k = 0; k < 3; k++) out[k] = vv[k] + R(2) * (q[0] * uv[k] + uuv[k]);
  }
  void agent_act(int f, int f_global) {  // agents/agent_injector.py:23-32 -> effectors/injector.py:80-105,240-256
    if (!has_injector) return;
    Effector<R>& e = eff[inj_idx];
    const EffectorCfg& ec = e.cfg;
    if (f_global >= inject_till) { e.act_id[f + 1] = e.act_id[f]; return; }
    for (int i = 0; i < ec.flux; i++) {
      int pid = e.act_range[e.act_id[f] + i];
      int ridx = ec.locally_random ? f : f_global;
      const R* rv = &e.random_vector[((size_t)ridx * ec.flux + i) * 3];
      size_t q = pi(f + 1, pid);
      if (ec.type == 1) {
        R ip[3] = {(R)ec.inject_p[0], (R)ec.inject_p[1], (R)ec.inject_p[2]}, ipr[3];
        quat_rotate(&e.quat[f * 4], ip, ipr);
        for (int k = 0; k < 3; k++) x[q * 3 + k] = (rv[k] * R(2) - R(1)) * (R)ec.radius + e.pos[f * 3 + k] + ipr[k];
        R iv[3] = {(R)ec.inject_v[0], (R)ec.inject_v[1], (R)ec.inject_v[2]}, ivr[3];
        quat_rotate(&e.quat[f * 4], iv, ivr);
        for (int k = 0; k < 3; k++) v[q * 3 + k] = ivr[k];
        if (ec.randomize_inject_v) {   // injector.py:96-97: + (random_vector * 2 - 1) * inject_v.norm() * 2.0 (no pose dependence: agent_act_grad is unchanged)
          const R nv = std::sqrt(iv[0] * iv[0] + iv[1] * iv[1] + iv[2] * iv[2]);
          for (int k = 0; k < 3; k++) v[q * 3 + k] = ivr[k] + (rv[k] * R(2) - R(1)) * nv * R(2);
        }
      } else {
        for (int k = 0; k < 3; k++) { x[q * 3 + k] = rv[k] + e.pos[f * 3 + k]; v[q * 3 + k] = (R)ec.inject_v[k]; }
      }
      used[q] = 1;
    }
    e.act_id[f + 1] = e.act_id[f] + ec.flux;
  }
  void agent_act_grad(int f, int f_global) {  // adjoint of the above w.r.t. pos[f] and quat[f]
    if (!has_injector || f_global >= inject_till) return;
    Effector<R>& e = eff[inj_idx];
    const EffectorCfg& ec = e.cfg;
    for (int i = 0; i < ec.flux; i++) {
      int pid = e.act_range[e.act_id[f] + i];
      size_t q = pi(f + 1, pid);
      for (int k = 0; k < 3; k++) e.gpos[f * 3 + k] += gx[q * 3 + k];
      if (ec.type == 1) {   // Injector rotates inject_p / inject_v by quat[f] (injector.py:93-96); BallInjector does not (:240-256)
        R ip[3] = {(R)ec.inject_p[0], (R)ec.inject_p[1], (R)ec.inject_p[2]}, iv[3] = {(R)ec.inject_v[0], (R)ec.inject_v[1], (R)ec.inject_v[2]};
        quat_rot_adj(&e.quat[f * 4], ip, &gx[q * 3], &e.gquat[f * 4], (R*)nullptr);
        quat_rot_adj(&e.quat[f * 4], iv, &gv[q * 3], &e.gquat[f * 4], (R*)nullptr);
      }
    }
  }
  void agent_move(int f) {  // effectors/effector.py:157-161 move_kernel
    for (auto& e : eff) {
      R in[3], out[3], jac[9];
      for (int k = 0; k < 3; k++) in[k] = e.pos[f * 3 + k] + e.v[f * 3 + k];
      effector_impose_x(e.cfg, in, out, jac);
      for (int k = 0; k < 3; k++) e.pos[(f + 1) * 3 + k] = out[k];
      // quat[f+1] = qmul(w2quat(w[f]), quat[f])   utils/geom.py:7-28
      R wv[3] = {e.w[f * 3], e.w[f * 3 + 1], e.w[f * 3 + 2]};
      R wn = std::sqrt(wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2] + R(1e-12));
      R q[4] = {std::cos(wn / 2), wv[0] / wn * std::sin(wn / 2), wv[1] / wn * std::sin(wn / 2), wv[2] / wn * std::sin(wn / 2)};
      const R* r = &e.quat[f * 4];
      R o[4] = {r[0] * q[0] - r[1] * q[1] - r[2] * q[2] - r[3] * q[3],
                r[0] * q[1] + r[1] * q[0] - r[2] * q[3] + r[3] * q[2],
                r[0] * q[2] + r[1] * q[3] + r[2] * q[0] - r[3] * q[1],
                r[0] * q[3] - r[1] * q[2] + r[2] * q[1] + r[3] * q[0]};
      R on = std::sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
      for (int k = 0; k < 4; k++) e.quat[(f + 1) * 4 + k] = o[k] / on;
    }
  }
  void agent_move_grad(int f) {  // move_kernel.grad (effector.py:155): position and orientation
    for (auto& e : eff) {
      R in[3], out[3], jac[9];
      for (int k = 0; k < 3; k++) in[k] = e.pos[f * 3 + k] + e.v[f * 3 + k];
      effector_impose_x(e.cfg, in, out, jac);
      for (int a = 0; a < 3; a++) {
        R g = 0;
        for (int b = 0; b < 3; b++) g += jac[b * 3 + a] * e.gpos[(f + 1) * 3 + b];
        e.gpos[f * 3 + a] += g; e.gv[f * 3 + a] += g;
      }
      // quat[f+1] = normalize(qmul_raw(w2quat(w[f]), quat[f]))   utils/geom.py:7-28
      const R* wv = &e.w[f * 3]; const R* r = &e.quat[f * 4]; const R* gq1 = &e.gquat[(f + 1) * 4];
      const R wn = std::sqrt(wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2] + R(1e-12));
      const R sh = std::sin(wn / 2), ch = std::cos(wn / 2);
      const R q[4] = {ch, wv[0] / wn * sh, wv[1] / wn * sh, wv[2] / wn * sh};
      const R o[4] = {r[0] * q[0] - r[1] * q[1] - r[2] * q[2] - r[3] * q[3], r[0] * q[1] + r[1] * q[0] - r[2] * q[3] + r[3] * q[2],
                      r[0] * q[2] + r[1] * q[3] + r[2] * q[0] - r[3] * q[1], r[0] * q[3] - r[1] * q[2] + r[2] * q[1] + r[3] * q[0]};
      const R on = std::sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
      R dd = 0; for (int k = 0; k < 4; k++) dd += o[k] / on * gq1[k];
      R go[4]; for (int k = 0; k < 4; k++) go[k] = (gq1[k] - o[k] / on * dd) / on;
      const R gr[4] = {go[0] * q[0] + go[1] * q[1] + go[2] * q[2] + go[3] * q[3], -go[0] * q[1] + go[1] * q[0] + go[2] * q[3] - go[3] * q[2],
                       -go[0] * q[2] - go[1] * q[3] + go[2] * q[0] + go[3] * q[1], -go[0] * q[3] + go[1] * q[2] - go[2] * q[1] + go[3] * q[0]};
      const R gq[4] = {go[0] * r[0] + go[1] * r[1] + go[2] * r[2] + go[3] * r[3], -go[0] * r[1] + go[1] * r[0] - go[2] * r[3] + go[3] * r[2],
                       -go[0] * r[2] + go[1] * r[3] + go[2] * r[0] - go[3] * r[1], -go[0] * r[3] - go[1] * r[2] + go[2] * r[1] + go[3] * r[0]};
      for (int k = 0; k < 4; k++) e.gquat[f * 4 + k] += gr[k];
      // w2quat: q0 = cos(wn/2), q_k = w_k / wn * sin(wn/2), wn = |w|_eps
      R gwn = -sh / 2 * gq[0];
      for (int k = 0; k < 3; k++) gwn += gq[k + 1] * wv[k] * (ch / 2 * wn - sh) / (wn * wn);
      for (int k = 0; k < 3; k++) e.gw[f * 3 + k] += gq[k + 1] * sh / wn + gwn * wv[k] / wn;
    }
  }
  void set_action(int ei, int s, int s_global, const R* action) {  // effector.py:218-221,252-260
    Effector<R>& e = eff[ei];
    const int ad = e.cfg.action_dim, ns = c.n_substeps;
    for (int j = 0; j < ad; j++) e.act[(size_t)s_global * ad + j] = action[j];
    for (int j = s * ns; j < (s + 1) * ns; j++) {
      for (int k = 0; k < 3; k++) e.v[j * 3 + k] = e.act[(size_t)s_global * ad + k] * (R)e.cfg.scale_v[k] / (R)ns;
      if (ad > 3) for (int k = 0; k < 3; k++) e.w[j * 3 + k] = e.act[(size_t)s_global * ad + k + 3] * (R)e.cfg.scale_v[k + 3] / (R)ns;
    }
  }
  void set_action_grad(int ei, int s, int s_global) {  // set_velocity.grad, effector.py:270-274
    Effector<R>& e = eff[ei];
    const int ad = e.cfg.action_dim, ns = c.n_substeps;
    for (int j = s * ns; j < (s + 1) * ns; j++) {
      for (int k = 0; k < 3; k++) e.gact[(size_t)s_global * ad + k] += e.gv[j * 3 + k] * (R)e.cfg.scale_v[k] / (R)ns;
      if (ad > 3) for (int k = 0; k < 3; k++) e.gact[(size_t)s_global * ad + k + 3] += e.gw[j * 3 + k] * (R)e.cfg.scale_v[k + 3] / (R)ns;
    }
  }
  void apply_action_p(int ei, const R* ap) {  // effector.py:223-231
    Effector<R>& e = eff[ei];
    for (int j = 0; j < e.cfg.action_dim; j++) e.act_p[j] = ap[j];
    R in[3], out[3], jac[9];
    for (int k = 0; k < 3; k++) in[k] = e.act_p[k] * (R)e.cfg.scale_p[k];
    effector_impose_x(e.cfg, in, out, jac);
    for (int k = 0; k < 3; k++) e.pos[k] = out[k];
  }
  void apply_action_p_grad(int ei) {  // effector.py:233-234
    Effector<R>& e = eff[ei];
    R in[3], out[3], jac[9];
    for (int k = 0; k < 3; k++) in[k] = e.act_p[k] * (R)e.cfg.scale_p[k];
    effector_impose_x(e.cfg, in, out, jac);
    for (int a = 0; a < 3; a++) {
      R g = 0; for (int b = 0; b < 3; b++) g += jac[b * 3 + a] * e.gpos[b];
      e.gact_p[a] += g * (R)e.cfg.scale_p[a];
    }
  }

  // ---------------------------------------------------------------- substep, MPM:515-533
  void substep(int f, int f_global, bool none_action) {
    reset_grid(); advect_used(f); process_unused(f);
    if (!none_action) { agent_act(f, f_global); collector_act(f); }
    compute_F_tmp_svd(f); p2g(f, true);
    if (!none_action) agent_move(f);
    grid_op(f); g2p(f); advect(f);
  }

  // ---------------------------------------------------------------- adjoints
  void g2p_advect_grad(int f) {  // advect_grad (MPM:436-447) then g2p.grad (MPM:538)
    const bool rb = any_rigid_body();
    if (rb) body_advect_grad(f);
#pragma omp parallel for
    for (int p = 0; p < N; p++) if (used[pi(f, p)]) {
      size_t a = pi(f, p), b = pi(f + 1, p);
      if (!(rb && cls[p] == MAT_RIGID))
        for (int k = 0; k < 3; k++) { gx[a * 3 + k] += gx[b * 3 + k]; gv[b * 3 + k] += dt * gx[b * 3 + k]; }
      const R* xp = &x[a * 3];
      int base[3]; R fx[3]; base_fx(xp, base, fx);
      R w[3][3], dw[3][3]; weights(fx, w); dweights(fx, dw);
      R gvn_buf[3] = {gv[b * 3], gv[b * 3 + 1], gv[b * 3 + 2]};
      if ((collide_type == 0 || collide_type == 2) && agent_type == 1 && has_rigid) {
        // recompute the pre-collision v' and back-propagate through agent.collide(f, x + dt v', v', dt)  (MPM:419-422)
        R nv0[3] = {0, 0, 0};
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
          const R* gg = &g_vout[gidx(base[0] + i, base[1] + j, base[2] + k) * 3];
          R weight = w[i][0] * w[j][1] * w[k][2];
          for (int r = 0; r < 3; r++) nv0[r] += weight * gg[r];
        }
        R xt[3] = {xp[0] + dt * nv0[0], xp[1] + dt * nv0[1], xp[2] + dt * nv0[2]}, o[3], gvpre[3] = {0, 0, 0}, gxt[3] = {0, 0, 0};
        agent_collide(f, xt, nv0, o, gvn_buf, gvpre, gxt);
        for (int r = 0; r < 3; r++) { gx[a * 3 + r] += gxt[r]; gvn_buf[r] = gvpre[r] + dt * gxt[r]; }
      }
      const R* gvn = gvn_buf;
      M3<R> gCn = getM(gC, b);
      R gfx[3] = {0, 0, 0};
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
        R del[3] = {(R)i - fx[0], (R)j - fx[1], (R)k - fx[2]};
        size_t g = gidx(base[0] + i, base[1] + j, base[2] + k);
        const R* gg = &g_vout[g * 3];
        R weight = w[i][0] * w[j][1] * w[k][2];
        R Cd[3], Ctg[3];
        for (int r = 0; r < 3; r++) {
          Cd[r] = gCn[r][0] * del[0] + gCn[r][1] * del[1] + gCn[r][2] * del[2];
          Ctg[r] = gCn[0][r] * gg[0] + gCn[1][r] * gg[1] + gCn[2][r] * gg[2];
        }
        R wbar = 0;
        for (int r = 0; r < 3; r++) {
          R val = weight * gvn[r] + R(4) * inv_dx * weight * Cd[r];
#pragma omp atomic
          gg_vout[g * 3 + r] += val;
          wbar += gg[r] * gvn[r] + R(4) * inv_dx * gg[r] * Cd[r];
        }
        R gw[3] = {dw[i][0] * w[j][1] * w[k][2], w[i][0] * dw[j][1] * w[k][2], w[i][0] * w[j][1] * dw[k][2]};
        for (int r = 0; r < 3; r++) gfx[r] += -(R(4) * inv_dx * weight * Ctg[r]) + wbar * gw[r];
      }
      for (int k = 0; k < 3; k++) gx[a * 3 + k] += inv_dx * gfx[k];
    }
  }
  void grid_op_grad(int f) {  // grid_op.grad, MPM:539
    const int n = c.n_grid;
#pragma omp parallel for
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) {
      size_t g = gidx(i, j, k);
      if (g_m[g] > R(1e-12)) {
        R inv_m = R(1) / g_m[g];
        R vv[3];
        for (int a = 0; a < 3; a++) { vv[a] = inv_m * g_vin[g * 3 + a]; vv[a] += dt * (R)c.gravity[a]; }
        R pos[3] = {(R)i * dx, (R)j * dx, (R)k * dx};
        const R zero3v[3] = {0, 0, 0}, idq[4] = {1, 0, 0, 0};
        // forward chain with the intermediate velocities kept
        std::vector<V3<R>> chain;
        chain.push_back(V3<R>{{vv[0], vv[1], vv[2]}});
        for (size_t si = 0; si < statics.size(); si++) {
          R o[3]; sdf_collide<R>(statics[si], false, zero3v, idq, zero3v, idq, dt, pos, chain.back().a, o, nullptr, nullptr, nullptr, nullptr, nullptr);
          chain.push_back(V3<R>{{o[0], o[1], o[2]}});
        }
        if (collide_type >= 1) { R o[3]; agent_collide(f, pos, chain.back().a, o, nullptr, nullptr, nullptr); chain.push_back(V3<R>{{o[0], o[1], o[2]}}); }
        R vlast[3] = {chain.back()[0], chain.back()[1], chain.back()[2]};
        R fac[3];
        boundary_v(pos, vlast, fac);
        R vb[3];
        for (int a = 0; a < 3; a++) vb[a] = gg_vout[g * 3 + a] * fac[a];
        int ci = (int)chain.size() - 1;
        if (collide_type >= 1) {
          R o[3], gvv[3] = {0, 0, 0}, gpp[3] = {0, 0, 0};
          agent_collide(f, pos, chain[ci - 1].a, o, vb, gvv, gpp);   // node positions are constants: gpp is dropped
          for (int a = 0; a < 3; a++) vb[a] = gvv[a];
          ci--;
        }
        for (int si = (int)statics.size() - 1; si >= 0; si--) {
          R o[3], gvv[3] = {0, 0, 0}, gpp[3] = {0, 0, 0}, g0[3] = {0, 0, 0}, g1[3] = {0, 0, 0};
          sdf_collide<R>(statics[si], false, zero3v, idq, zero3v, idq, dt, pos, chain[ci - 1].a, o, vb, gvv, gpp, g0, g1);
          for (int a = 0; a < 3; a++) vb[a] = gvv[a];
          ci--;
        }
        R mbar = 0;
        for (int a = 0; a < 3; a++) {
          gg_vin[g * 3 + a] += vb[a] * inv_m;
          mbar += -(g_vin[g * 3 + a] * vb[a]) * inv_m * inv_m;
        }
        gg_m[g] += mbar;
      }
    }
  }
  void p2g_grad(int f) {  // p2g.grad (MPM:544) + svd_grad (MPM:545) + compute_F_tmp.grad (MPM:546)
#pragma omp parallel for
    for (int p = 0; p < N; p++) if (used[pi(f, p)]) {
      size_t a = pi(f, p), b = pi(f + 1, p);
      const R* xp = &x[a * 3];
      const R* vp = &v[a * 3];
      int base[3]; R fx[3]; base_fx(xp, base, fx);
      R w[3][3], dw[3][3]; weights(fx, w); dweights(fx, dw);
      R J; M3<R> A = affine_of(f, p, J);
      M3<R> Ft = getM(Ftmp, p), U = getM(Usv, p), V = getM(Vsv, p);
      const R* s = &Ssv[(size_t)p * 3];
      const R m = mass[p];
      R gvp[3] = {0, 0, 0}, gfx[3] = {0, 0, 0};
      M3<R> gA = zero3<R>();
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
        R d[3] = {((R)i - fx[0]) * dx, ((R)j - fx[1]) * dx, ((R)k - fx[2]) * dx};
        size_t g = gidx(base[0] + i, base[1] + j, base[2] + k);
        const R* gvi = &gg_vin[g * 3];
        R gmi = gg_m[g];
        R weight = w[i][0] * w[j][1] * w[k][2];
        R wbar = m * gmi;
        R Atg[3];
        for (int r = 0; r < 3; r++) {
          gvp[r] += m * weight * gvi[r];
          R Ad = A[r][0] * d[0] + A[r][1] * d[1] + A[r][2] * d[2];
          wbar += gvi[r] * (m * vp[r] + Ad);
          for (int q = 0; q < 3; q++) gA[r][q] += weight * gvi[r] * d[q];
          Atg[r] = A[0][r] * gvi[0] + A[1][r] * gvi[1] + A[2][r] * gvi[2];
        }
        R gw[3] = {dw[i][0] * w[j][1] * w[k][2], w[i][0] * dw[j][1] * w[k][2], w[i][0] * w[j][1] * dw[k][2]};
        for (int r = 0; r < 3; r++) gfx[r] += -dx * weight * Atg[r] + wbar * gw[r];
      }
      for (int k = 0; k < 3; k++) { gv[a * 3 + k] += gvp[k]; gx[a * 3 + k] += inv_dx * gfx[k]; }
      addM(gC, a, scale(gA, m));
      M3<R> gP = scale(gA, k_stress);
      M3<R> R_ = mul(U, tr(V));
      M3<R> Mm = sub(Ft, R_);
      M3<R> gM = scale(mul(gP, Ft), R(2) * mu[p]);
      M3<R> gFt = add(gM, scale(mul(tr(gP), Mm), R(2) * mu[p]));
      M3<R> gR = scale(gM, R(-1));
      R gJ = lam[p] * (R(2) * J - R(1)) * trace3(gP);
      M3<R> gU = mul(gR, V);
      M3<R> gV = mul(tr(gR), U);
      M3<R> gS = zero3<R>();
      M3<R> gFn = getM(gF, b);
      if (cls[p] == MAT_LIQUID) {
        // F_new = I * pow(J, 1/3): dJ += (1/3) J^(1/3 - 1) tr(gFn)
        gJ += R(1.0 / 3) * (R)std::pow(J, R(1.0 / 3) - R(1)) * trace3(gFn);
      } else if (cls[p] == MAT_ELASTIC || cls[p] == MAT_RIGID) {
        gFt = add(gFt, gFn);
      } else if (cls[p] == MAT_PLASTO_ELASTIC || cls[p] == MAT_PLASTO_ELASTIC_DEMO) {
        M3<R> Sn = zero3<R>(); R pass[3];
        for (int d = 0; d < 3; d++) {
          R sv = s[d];
          R mx = std::max(sv, R(1 - 2e-3)); bool p1 = R(1 - 2e-3) < sv;   // max(a,b) -> a iff b < a
          R mn = std::min(mx, R(1 + 3e-3)); bool p2 = mx < R(1 + 3e-3);   // min(a,b) -> a iff a < b
          Sn[d][d] = mn; pass[d] = (p1 && p2) ? R(1) : R(0);
        }
        gU = add(gU, mul(gFn, mul(V, Sn)));
        gV = add(gV, mul(tr(gFn), mul(U, Sn)));
        M3<R> t = mul(tr(U), mul(gFn, V));
        for (int d = 0; d < 3; d++) gS[d][d] += t[d][d] * pass[d];
      }
      gS[0][0] += gJ * s[1] * s[2]; gS[1][1] += gJ * s[0] * s[2]; gS[2][2] += gJ * s[0] * s[1];
      gFt = add(gFt, backward_svd(gU, gS, gV, U, s, V));
      // compute_F_tmp.grad, MPM:254-258
      M3<R> Cm = getM(C, a), Fm = getM(F, a);
      addM(gC, a, scale(mul(gFt, tr(Fm)), dt));
      addM(gF, a, mul(tr(add(ident3<R>(), scale(Cm, dt))), gFt));
    }
  }
  void process_unused_grad(int f) {  // MPM:551
#pragma omp parallel for
    for (int p = 0; p < N; p++) if (used[pi(f, p)] == 0) {
      size_t a = pi(f, p), b = pi(f + 1, p);
      for (int k = 0; k < 3; k++) { gv[a * 3 + k] += gv[b * 3 + k]; gx[a * 3 + k] += gx[b * 3 + k]; }
      for (int k = 0; k < 9; k++) { gC[a * 9 + k] += gC[b * 9 + k]; gF[a * 9 + k] += gF[b * 9 + k]; }
    }
  }
  void substep_grad(int f, int f_global, bool none_action) {  // MPM:535-552
    // recompute the forward scratch of frame f (the reference keeps it per frame instead)
    reset_grid(); compute_F_tmp_svd(f); p2g(f, false); grid_op(f);
    g2p_advect_grad(f);
    grid_op_grad(f);
    if (!none_action) agent_move_grad(f);
    p2g_grad(f);
    if (!none_action) agent_act_grad(f, f_global);
    process_unused_grad(f);
  }

  // ---------------------------------------------------------------- losses/shapematching_loss.py:80-93
  double loss_value(int f, int matching_mat, double weight, const double* tgt) const {
    R acc = 0;
    for (int p = 0; p < N; p++) if (used[pi(f, p)] && mat[p] == matching_mat)
      for (int k = 0; k < 3; k++) { R d = x[pi(f, p) * 3 + k] - (R)tgt[(size_t)p * 3 + k]; acc += d * d; }
    return (double)(acc * (R)weight);
  }
  void loss_seed(int f, int matching_mat, double weight, const double* tgt) {
    for (int p = 0; p < N; p++) if (used[pi(f, p)] && mat[p] == matching_mat)
      for (int k = 0; k < 3; k++) gx[pi(f, p) * 3 + k] += R(2) * (R)weight * (x[pi(f, p) * 3 + k] - (R)tgt[(size_t)p * 3 + k]);
  }

  // ---------------------------------------------------------------- ring helpers MPM:588-609
  void copy_frame(int s, int t) {
    for (int p = 0; p < N; p++) {
      size_t a = pi(s, p), b = pi(t, p);
      for (int k = 0; k < 3; k++) { x[b * 3 + k] = x[a * 3 + k]; v[b * 3 + k] = v[a * 3 + k]; }
      for (int k = 0; k < 9; k++) { C[b * 9 + k] = C[a * 9 + k]; F[b * 9 + k] = F[a * 9 + k]; }
      used[b] = used[a];
    }
    for (auto& e : eff) {
      for (int k = 0; k < 3; k++) { e.pos[t * 3 + k] = e.pos[s * 3 + k]; e.v[t * 3 + k] = e.v[s * 3 + k]; e.w[t * 3 + k] = e.w[s * 3 + k]; }
      for (int k = 0; k < 4; k++) e.quat[t * 4 + k] = e.quat[s * 4 + k];
      e.act_id[t] = e.act_id[s];
    }
  }
  void copy_grad(int s, int t) {
    for (int p = 0; p < N; p++) {
      size_t a = pi(s, p), b = pi(t, p);
      for (int k = 0; k < 3; k++) { gx[b * 3 + k] = gx[a * 3 + k]; gv[b * 3 + k] = gv[a * 3 + k]; }
      for (int k = 0; k < 9; k++) { gC[b * 9 + k] = gC[a * 9 + k]; gF[b * 9 + k] = gF[a * 9 + k]; }
      used[b] = used[a];
    }
    for (auto& e : eff) {
      for (int k = 0; k < 3; k++) { e.gpos[t * 3 + k] = e.gpos[s * 3 + k]; e.gv[t * 3 + k] = e.gv[s * 3 + k]; e.gw[t * 3 + k] = e.gw[s * 3 + k]; }
      for (int k = 0; k < 4; k++) e.gquat[t * 4 + k] = e.gquat[s * 4 + k];
    }
  }
  void reset_grad_till(int f) {
    std::fill(gx.begin(), gx.begin() + (size_t)f * N * 3, R(0)); std::fill(gv.begin(), gv.begin() + (size_t)f * N * 3, R(0));
    std::fill(gC.begin(), gC.begin() + (size_t)f * N * 9, R(0)); std::fill(gF.begin(), gF.begin() + (size_t)f * N * 9, R(0));
    for (auto& e : eff) {
      std::fill(e.gpos.begin(), e.gpos.begin() + f * 3, R(0)); std::fill(e.gv.begin(), e.gv.begin() + f * 3, R(0));
      std::fill(e.gw.begin(), e.gw.begin() + f * 3, R(0)); std::fill(e.gquat.begin(), e.gquat.begin() + f * 4, R(0));
    }
  }
  void reset_grad() {
    std::fill(gx.begin(), gx.end(), R(0)); std::fill(gv.begin(), gv.end(), R(0));
    std::fill(gC.begin(), gC.end(), R(0)); std::fill(gF.begin(), gF.end(), R(0));
    for (auto& e : eff) {
      std::fill(e.gpos.begin(), e.gpos.end(), R(0)); std::fill(e.gv.begin(), e.gv.end(), R(0));
      std::fill(e.gw.begin(), e.gw.end(), R(0)); std::fill(e.gquat.begin(), e.gquat.end(), R(0));
      std::fill(e.gact.begin(), e.gact.end(), R(0)); std::fill(e.gact_p.begin(), e.gact_p.end(), R(0));
    }
  }
};

}  // namespace orc
