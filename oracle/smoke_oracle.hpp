// =====================================================================================
// oracle/smoke_oracle.hpp — TEST INFRASTRUCTURE ONLY (never linked/imported by the product).
//
// CPU restatement of the reference's Eulerian smoke solver, fluidlab/fluidengine/simulators/smoke_field.py (abbrev. SF):
// one step = free-space mask, RK3 semi-Lagrangian advection + air-conditioner impulse, divergence, `solver_iters` Jacobi
// sweeps of the pressure Poisson problem, projection (SF:95-110); and its reverse-mode adjoint (SF:112-127, where every
// `kernel.grad` is Taichi autodiff: the hand-derived adjoints below follow the same conventions as oracle/mpm_oracle.hpp —
// integer casts, floor and branch conditions carry no gradient, abs'(0) = 0).
//
// PARITY STATUS.  Forward: pinned to runs of the reference's own kernel source on the NumPy emulation of the Taichi API
// (tests/golden/make_reference_smoke.py -> tests/golden/reference_smoke.npz, tests/test_smoke_oracle.py).  Adjoint: pinned to central
// finite differences of this oracle in float64 and through the reference's own forward kernels run in float64 on the emulation.
// Taichi itself is not installable here, so float32 rounding of compile-time constants (e.g. 0.75 * dt) is only matched to
// tolerance, not bit for bit.
//
// One deliberate deviation: SF:301-310 `compute_location` falls back to the UNCLAMPED index when the clamped cell is not free; for a
// sample point within half a cell of the domain edge next to a blocked edge cell that index is out of range (an out-of-bounds field
// read in Taichi).  Here the clamped index is used in that case.
// =====================================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "mpm_oracle.hpp"

namespace orc {

struct SmokeConfig {
  int n;              // res (SF:14)
  int S;              // max_steps_local: S + 1 step frames (SF:74)
  int q_dim;          // SF:14
  int solver_iters;   // SF:14
  double dt;          // SF:14 (0.03)
  int lower_y, higher_y;   // SF:26-27: free band lower_y < j < higher_y
  double high_T, low_T;    // SF:24-25
  int T_sub;          // max_substeps_local of the MPM ring: the air conditioner's arrays are indexed by substep f (SF:217-221)
  double inject_v[3]; // effectors/aircon.py:14
  double mpm_dx;      // not used by the live code path (the MPM coupling is commented out, SF:224-227)
};

template <class R> struct Smoke {
  SmokeConfig c;
  int n; size_t G;
  R dx;
  std::vector<R> v, v_tmp, dv, p, q;            // [S+1][G][3|3|1|1|q_dim]   (dv = div)
  std::vector<R> gv, gv_tmp, gdv, gp, gq;       // grads
  std::vector<int> is_free;                     // [S+1][G]
  std::vector<R> cur, nxt, gcur, gnxt;          // p_swap (SF:77-80)
  std::vector<SdfMesh<R>> statics;              // room etc. (SF:197-200)
  // air conditioner (effectors/aircon.py): per-substep pose, strength s and radius r, and their adjoints
  std::vector<R> a_pos, a_quat, a_s, a_r, ga_pos, ga_quat, ga_s, ga_r;

  explicit Smoke(const SmokeConfig& cfg) : c(cfg) {
    n = c.n; G = (size_t)n * n * n; dx = R(1) / (R)n;
    const size_t F = (size_t)(c.S + 1) * G;
    v.assign(F * 3, 0); v_tmp.assign(F * 3, 0); dv.assign(F, 0); p.assign(F, 0); q.assign(F * c.q_dim, 0);
    gv.assign(F * 3, 0); gv_tmp.assign(F * 3, 0); gdv.assign(F, 0); gp.assign(F, 0); gq.assign(F * c.q_dim, 0);
    is_free.assign(F, 0);
    cur.assign(G, 0); nxt.assign(G, 0); gcur.assign(G, 0); gnxt.assign(G, 0);
    const size_t T1 = (size_t)c.T_sub + 1;
    a_pos.assign(T1 * 3, 0); a_quat.assign(T1 * 4, 0); a_s.assign(T1, 0); a_r.assign(T1, 0);
    ga_pos.assign(T1 * 3, 0); ga_quat.assign(T1 * 4, 0); ga_s.assign(T1, 0); ga_r.assign(T1, 0);
    for (size_t f = 0; f < T1; f++) a_quat[f * 4] = 1;
    // init_fields, SF:87-93: q[0] = high_T inside the band (first component only: ti.Vector([high_T]) has one entry)
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++)
      if (c.lower_y < j && j < c.higher_y) q[cell(0, i, j, k) * c.q_dim] = (R)c.high_T;
  }

  inline size_t cell(int s, int i, int j, int k) const { return (size_t)s * G + ((size_t)i * n + j) * n + k; }
  inline bool in_range(int i, int j, int k) const { return i >= 0 && j >= 0 && k >= 0 && i < n && j < n && k < n; }

  // ---------------------------------------------------------------- SF:190-201
  void compute_free_space(int s) {
#pragma omp parallel for
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) {
      int fr = (c.lower_y < j && j < c.higher_y) ? 1 : 0;
      if (fr) {
        R pw[3] = {((R)i + R(0.5)) * dx, ((R)j + R(0.5)) * dx, ((R)k + R(0.5)) * dx};
        for (const auto& M : statics) {   // Static.is_collide, meshes/static.py:106-114
          if (!M.has_dynamics) continue;
          R pv[3]; M.to_voxels(pw, pv);
          if (M.sdf_(pv, nullptr) <= R(0)) fr = 0;
        }
      }
      is_free[cell(s, i, j, k)] = fr;
    }
  }

  // compute_location, SF:301-310: index of the sample used for the neighbour (u+du, v+dv, w+dw)
  inline void locate(int s, int u, int v_, int w, int du, int dv_, int dw, int* I) const {
    I[0] = std::max(0, std::min(n - 1, u + du)); I[1] = std::max(0, std::min(n - 1, v_ + dv_)); I[2] = std::max(0, std::min(n - 1, w + dw));
    if (!is_free[cell(s, I[0], I[1], I[2])]) {
      if (in_range(u, v_, w)) { I[0] = u; I[1] = v_; I[2] = w; }   // else: keep the clamped cell (see the header note)
    }
  }
  inline size_t loc(int s, int u, int v_, int w, int du, int dv_, int dw) const { int I[3]; locate(s, u, v_, w, du, dv_, dw, I); return cell(s, I[0], I[1], I[2]); }
  // is_free(), SF:312-323
  inline bool free_at(int s, int u, int v_, int w, int du, int dv_, int dw) const {
    const int i = u + du, j = v_ + dv_, k = w + dw;
    return in_range(i, j, k) && is_free[cell(s, i, j, k)];
  }

  // trilerp, SF:325-347.  nc components of field `fld` (frame s).  Optional outputs for the adjoint.
  struct Tri { size_t idx[8]; R w[8]; R W; R dwdp[8][3]; };
  inline void trilerp(int s, const std::vector<R>& fld, int nc, const R* pp, R* out, Tri* t) const {
    int base[3]; R pI[3];
    for (int d = 0; d < 3; d++) { base[d] = (int)std::floor(pp[d] - R(0.5)); pI[d] = pp[d] - R(0.5); }
    for (int a = 0; a < nc; a++) out[a] = 0;
    R W = 0; int m = 0;
    for (int oi = 0; oi < 2; oi++) for (int oj = 0; oj < 2; oj++) for (int ok = 0; ok < 2; ok++, m++) {
      const int g[3] = {base[0] + oi, base[1] + oj, base[2] + ok};
      R wx[3], sg[3];
      for (int d = 0; d < 3; d++) { const R t_ = pI[d] - (R)g[d]; wx[d] = R(1) - std::fabs(t_); sg[d] = t_ > 0 ? R(1) : (t_ < 0 ? R(-1) : R(0)); }
      const R w = wx[0] * wx[1] * wx[2];
      const size_t id = loc(s, g[0], g[1], g[2], 0, 0, 0);
      for (int a = 0; a < nc; a++) out[a] += w * fld[id * nc + a];
      W += w;
      if (t) { t->idx[m] = id; t->w[m] = w; t->dwdp[m][0] = -sg[0] * wx[1] * wx[2]; t->dwdp[m][1] = -wx[0] * sg[1] * wx[2]; t->dwdp[m][2] = -wx[0] * wx[1] * sg[2]; }
    }
    for (int a = 0; a < nc; a++) out[a] /= W;
    if (t) t->W = W;
  }
  // adjoint of trilerp: given g_out, accumulate into the field adjoint (Tri indices are absolute: they include the frame offset) and into g_p (3)
  inline void trilerp_adj(const std::vector<R>& fld, std::vector<R>& gfld, int nc, const Tri& t, const R* out, const R* g_out, R* g_p) {
    R gW = 0; for (int a = 0; a < nc; a++) gW -= g_out[a] * out[a]; gW /= t.W;
    for (int m = 0; m < 8; m++) {
      R gw = gW;
      for (int a = 0; a < nc; a++) {
        gw += g_out[a] * fld[t.idx[m] * nc + a] / t.W;
        const R add = g_out[a] * t.w[m] / t.W;
#pragma omp atomic
        gfld[t.idx[m] * nc + a] += add;
      }
      if (g_p) for (int d = 0; d < 3; d++) g_p[d] += gw * t.dwdp[m][d];
    }
  }

  // ---------------------------------------------------------------- SF:203-233
  struct Impulse { R imp_dir[3], d[3], dist, factor; };
  inline void impulse(int f, int i, int j, int k, Impulse& I) const {
    const R inj[3] = {(R)c.inject_v[0], (R)c.inject_v[1], (R)c.inject_v[2]};
    quat_rot_t(&a_quat[(size_t)f * 4], inj, I.imp_dir);
    const R ijk[3] = {(R)i, (R)j, (R)k};
    R ss = 0;
    for (int d = 0; d < 3; d++) { I.d[d] = ijk[d] - a_pos[(size_t)f * 3 + d] / dx; ss += I.d[d] * I.d[d]; }
    I.dist = std::sqrt(ss + R(1e-12));   // norm(EPS), configs/macros.py:213
    I.factor = std::exp(-I.dist / a_r[f]);
  }
  void advect_and_impulse(int s, int f) {
    const R dt = (R)c.dt;
    const int qd = c.q_dim;
#pragma omp parallel for
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) {
      const size_t g = cell(s, i, j, k), g1 = cell(s + 1, i, j, k);
      if (is_free[g]) {
        R p0[3] = {(R)i + R(0.5), (R)j + R(0.5), (R)k + R(0.5)}, pe[3];
        backtrace(s, p0, pe, nullptr);
        R vf[3], qf[8];
        trilerp(s, v, 3, pe, vf, nullptr); trilerp(s, q, qd, pe, qf, nullptr);
        Impulse I; impulse(f, i, j, k, I);
        for (int d = 0; d < 3; d++) v_tmp[g * 3 + d] = vf[d] * R(1) + (I.imp_dir[d] * a_s[f] * I.factor) * dt + R(0);
        for (int a = 0; a < qd; a++) q[g1 * qd + a] = (R(1) - I.factor) * (qf[a] * R(1)) + I.factor * (R)c.low_T;
      } else {
        for (int d = 0; d < 3; d++) v_tmp[g * 3 + d] = 0;
        for (int a = 0; a < qd; a++) q[g1 * qd + a] = q[g * qd + a];
      }
    }
  }
  // backtrace (RK3), SF:349-360
  struct Trace { Tri t1, t2, t3; R v1[3], v2[3], v3[3], p1[3], p2[3]; };
  inline void backtrace(int s, const R* p0, R* pe, Trace* tr) const {
    const R dt = (R)c.dt;
    R v1[3], v2[3], v3[3], p1[3], p2[3];
    trilerp(s, v, 3, p0, v1, tr ? &tr->t1 : nullptr);
    for (int d = 0; d < 3; d++) p1[d] = p0[d] - R(0.5) * dt * v1[d];
    trilerp(s, v, 3, p1, v2, tr ? &tr->t2 : nullptr);
    for (int d = 0; d < 3; d++) p2[d] = p0[d] - R(0.75) * dt * v2[d];
    trilerp(s, v, 3, p2, v3, tr ? &tr->t3 : nullptr);
    for (int d = 0; d < 3; d++) pe[d] = p0[d] - dt * (R(2.0 / 9.0) * v1[d] + R(1.0 / 3.0) * v2[d] + R(4.0 / 9.0) * v3[d]);
    if (tr) for (int d = 0; d < 3; d++) { tr->v1[d] = v1[d]; tr->v2[d] = v2[d]; tr->v3[d] = v3[d]; tr->p1[d] = p1[d]; tr->p2[d] = p2[d]; }
  }

  // ---------------------------------------------------------------- SF:235-261
  void divergence(int s) {
#pragma omp parallel for
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) {
      const size_t g = cell(s, i, j, k);
      if (!is_free[g]) continue;
      const R* vc = &v_tmp[g * 3];
      R lo[3], hi[3];
      for (int d = 0; d < 3; d++) {
        const int e[3] = {d == 0, d == 1, d == 2};
        lo[d] = free_at(s, i, j, k, -e[0], -e[1], -e[2]) ? v_tmp[loc(s, i, j, k, -e[0], -e[1], -e[2]) * 3 + d] : -vc[d];
        hi[d] = free_at(s, i, j, k, e[0], e[1], e[2]) ? v_tmp[loc(s, i, j, k, e[0], e[1], e[2]) * 3 + d] : -vc[d];
      }
      dv[g] = (hi[0] - lo[0] + hi[1] - lo[1] + hi[2] - lo[2]) * R(0.5);
    }
  }

  // ---------------------------------------------------------------- SF:97-106,129-146,263-273
  void reset_swap_and_grad() { std::fill(cur.begin(), cur.end(), R(0)); std::fill(nxt.begin(), nxt.end(), R(0)); std::fill(gcur.begin(), gcur.end(), R(0)); std::fill(gnxt.begin(), gnxt.end(), R(0)); }
  void pressure_solve(int s) {
    reset_swap_and_grad();
#pragma omp parallel for
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) { const size_t g = cell(s, i, j, k); if (is_free[g]) cur[g - (size_t)s * G] = p[g]; }
    for (int it = 0; it < c.solver_iters; it++) {
#pragma omp parallel for
      for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) {
        const size_t g = cell(s, i, j, k);
        if (!is_free[g]) continue;
        const size_t o = (size_t)s * G;
        const R pl = cur[loc(s, i, j, k, -1, 0, 0) - o], pr = cur[loc(s, i, j, k, 1, 0, 0) - o], pb = cur[loc(s, i, j, k, 0, -1, 0) - o],
                pt = cur[loc(s, i, j, k, 0, 1, 0) - o], pp = cur[loc(s, i, j, k, 0, 0, -1) - o], pq = cur[loc(s, i, j, k, 0, 0, 1) - o];
        nxt[g - o] = (pl + pr + pb + pt + pp + pq - dv[g]) / R(6.0);
      }
      cur.swap(nxt);
    }
#pragma omp parallel for
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) { const size_t g = cell(s, i, j, k); if (is_free[g]) p[g + G] = cur[g - (size_t)s * G]; }
    reset_swap_and_grad();
  }

  // ---------------------------------------------------------------- SF:275-289
  void subtract_gradient(int s) {
#pragma omp parallel for
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) {
      const size_t g = cell(s, i, j, k), g1 = g + G;
      if (is_free[g]) {
        const size_t o = G;   // p[s+1]
        const R pl = p[loc(s, i, j, k, -1, 0, 0) + o], pr = p[loc(s, i, j, k, 1, 0, 0) + o], pb = p[loc(s, i, j, k, 0, -1, 0) + o],
                pt = p[loc(s, i, j, k, 0, 1, 0) + o], pp = p[loc(s, i, j, k, 0, 0, -1) + o], pq = p[loc(s, i, j, k, 0, 0, 1) + o];
        v[g1 * 3 + 0] = v_tmp[g * 3 + 0] - R(0.5) * (pr - pl);
        v[g1 * 3 + 1] = v_tmp[g * 3 + 1] - R(0.5) * (pt - pb);
        v[g1 * 3 + 2] = v_tmp[g * 3 + 2] - R(0.5) * (pq - pp);
      } else {
        for (int d = 0; d < 3; d++) v[g1 * 3 + d] = v_tmp[g * 3 + d];
      }
    }
  }

  void step(int s, int f) {  // SF:95-110 (colorize is renderer-only)
    compute_free_space(s);
    advect_and_impulse(s, f);
    divergence(s);
    pressure_solve(s);
    subtract_gradient(s);
  }

  // ================================================================ adjoint, SF:112-127
  void subtract_gradient_grad(int s) {
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) {
      const size_t g = cell(s, i, j, k), g1 = g + G;
      for (int d = 0; d < 3; d++) gv_tmp[g * 3 + d] += gv[g1 * 3 + d];
      if (!is_free[g]) continue;
      const size_t o = G;
      for (int d = 0; d < 3; d++) {
        const int e[3] = {d == 0, d == 1, d == 2};
        const R h = R(0.5) * gv[g1 * 3 + d];
        gp[loc(s, i, j, k, e[0], e[1], e[2]) + o] -= h;
        gp[loc(s, i, j, k, -e[0], -e[1], -e[2]) + o] += h;
      }
    }
  }
  void pressure_solve_grad(int s) {
    const size_t o = (size_t)s * G;
    reset_swap_and_grad();
    for (size_t g = 0; g < G; g++) if (is_free[o + g]) gcur[g] += gp[o + G + g];   // pressure_from_swap.grad
    for (int it = c.solver_iters - 1; it >= 0; it--) {
      gcur.swap(gnxt);
      std::fill(gcur.begin(), gcur.end(), R(0));
      for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) {   // pressure_jacobi.grad
        const size_t g = cell(s, i, j, k);
        if (!is_free[g]) continue;
        const R gn = gnxt[g - o] / R(6.0);
        gdv[g] -= gn;
        gcur[loc(s, i, j, k, -1, 0, 0) - o] += gn; gcur[loc(s, i, j, k, 1, 0, 0) - o] += gn; gcur[loc(s, i, j, k, 0, -1, 0) - o] += gn;
        gcur[loc(s, i, j, k, 0, 1, 0) - o] += gn; gcur[loc(s, i, j, k, 0, 0, -1) - o] += gn; gcur[loc(s, i, j, k, 0, 0, 1) - o] += gn;
      }
    }
    for (size_t g = 0; g < G; g++) if (is_free[o + g]) gp[o + g] += gcur[g];              // pressure_to_swap.grad
    reset_swap_and_grad();
  }
  void divergence_grad(int s) {
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) {
      const size_t g = cell(s, i, j, k);
      if (!is_free[g]) continue;
      const R h = gdv[g] * R(0.5);
      for (int d = 0; d < 3; d++) {
        const int e[3] = {d == 0, d == 1, d == 2};
        if (free_at(s, i, j, k, e[0], e[1], e[2])) gv_tmp[loc(s, i, j, k, e[0], e[1], e[2]) * 3 + d] += h; else gv_tmp[g * 3 + d] -= h;
        if (free_at(s, i, j, k, -e[0], -e[1], -e[2])) gv_tmp[loc(s, i, j, k, -e[0], -e[1], -e[2]) * 3 + d] -= h; else gv_tmp[g * 3 + d] += h;
      }
    }
  }
  void advect_and_impulse_grad(int s, int f) {
    const R dt = (R)c.dt;
    const int qd = c.q_dim;
    R acc_pos[3] = {0, 0, 0}, acc_quat[4] = {0, 0, 0, 0}, acc_s = 0, acc_r = 0;
#pragma omp parallel for reduction(+ : acc_pos[:3], acc_quat[:4], acc_s, acc_r)
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) {
      const size_t g = cell(s, i, j, k), g1 = cell(s + 1, i, j, k);
      if (!is_free[g]) {
        for (int a = 0; a < qd; a++) {
#pragma omp atomic
          gq[g * qd + a] += gq[g1 * qd + a];
        }
        continue;
      }
      // forward replay
      R p0[3] = {(R)i + R(0.5), (R)j + R(0.5), (R)k + R(0.5)}, pe[3];
      Trace tr; backtrace(s, p0, pe, &tr);
      R vf[3], qf[8]; Tri tv, tq;
      trilerp(s, v, 3, pe, vf, &tv); trilerp(s, q, qd, pe, qf, &tq);
      Impulse I; impulse(f, i, j, k, I);
      // outputs' adjoints
      const R* gvt = &gv_tmp[g * 3];
      R g_factor = 0, g_qf[8];
      for (int a = 0; a < qd; a++) { const R go = gq[g1 * qd + a]; g_qf[a] = (R(1) - I.factor) * go; g_factor += go * ((R)c.low_T - qf[a]); }
      R g_dir[3], dot = 0;
      for (int d = 0; d < 3; d++) { g_dir[d] = gvt[d] * a_s[f] * I.factor * dt; dot += gvt[d] * I.imp_dir[d]; }
      acc_s += dot * I.factor * dt;
      g_factor += dot * a_s[f] * dt;
      // factor = exp(-dist / r)
      const R g_dist = g_factor * I.factor * (-R(1) / a_r[f]);
      acc_r += g_factor * I.factor * (I.dist / (a_r[f] * a_r[f]));
      for (int d = 0; d < 3; d++) acc_pos[d] += -(g_dist * I.d[d] / I.dist) / dx;
      const R inj[3] = {(R)c.inject_v[0], (R)c.inject_v[1], (R)c.inject_v[2]};
      R gq4[4] = {0, 0, 0, 0};
      quat_rot_adj(&a_quat[(size_t)f * 4], inj, g_dir, gq4, (R*)nullptr);
      for (int d = 0; d < 4; d++) acc_quat[d] += gq4[d];
      // the two lookups at the end point
      R g_pe[3] = {0, 0, 0};
      trilerp_adj(v, gv, 3, tv, vf, gvt, g_pe);
      trilerp_adj(q, gq, qd, tq, qf, g_qf, g_pe);
      // backtrace adjoint
      R g_v1[3], g_v2[3], g_v3[3], g_p2[3] = {0, 0, 0}, g_p1[3] = {0, 0, 0};
      for (int d = 0; d < 3; d++) { g_v1[d] = -dt * R(2.0 / 9.0) * g_pe[d]; g_v2[d] = -dt * R(1.0 / 3.0) * g_pe[d]; g_v3[d] = -dt * R(4.0 / 9.0) * g_pe[d]; }
      trilerp_adj(v, gv, 3, tr.t3, tr.v3, g_v3, g_p2);
      for (int d = 0; d < 3; d++) g_v2[d] += -R(0.75) * dt * g_p2[d];
      trilerp_adj(v, gv, 3, tr.t2, tr.v2, g_v2, g_p1);
      for (int d = 0; d < 3; d++) g_v1[d] += -R(0.5) * dt * g_p1[d];
      trilerp_adj(v, gv, 3, tr.t1, tr.v1, g_v1, nullptr);
    }
    for (int d = 0; d < 3; d++) ga_pos[(size_t)f * 3 + d] += acc_pos[d];
    for (int d = 0; d < 4; d++) ga_quat[(size_t)f * 4 + d] += acc_quat[d];
    ga_s[f] += acc_s; ga_r[f] += acc_r;
  }

  void step_grad(int s, int f) {
    compute_free_space(s);
    subtract_gradient_grad(s);
    pressure_solve_grad(s);
    divergence_grad(s);
    advect_and_impulse_grad(s, f);
  }

  void reset_grad() {  // SF:180-183
    std::fill(gv.begin(), gv.end(), R(0)); std::fill(gv_tmp.begin(), gv_tmp.end(), R(0)); std::fill(gdv.begin(), gdv.end(), R(0));
    std::fill(gp.begin(), gp.end(), R(0)); std::fill(gq.begin(), gq.end(), R(0));
    std::fill(gcur.begin(), gcur.end(), R(0)); std::fill(gnxt.begin(), gnxt.end(), R(0));
    std::fill(ga_pos.begin(), ga_pos.end(), R(0)); std::fill(ga_quat.begin(), ga_quat.end(), R(0)); std::fill(ga_s.begin(), ga_s.end(), R(0)); std::fill(ga_r.begin(), ga_r.end(), R(0));
  }
  void copy_frame(int src, int dst) {  // SF:162-169
    auto cp = [&](std::vector<R>& a, int nc) { std::copy(a.begin() + (size_t)src * G * nc, a.begin() + (size_t)(src + 1) * G * nc, a.begin() + (size_t)dst * G * nc); };
    cp(v, 3); cp(v_tmp, 3); cp(dv, 1); cp(p, 1); cp(q, c.q_dim);
  }
  void copy_grad(int src, int dst) {  // SF:171-178
    auto cp = [&](std::vector<R>& a, int nc) { std::copy(a.begin() + (size_t)src * G * nc, a.begin() + (size_t)(src + 1) * G * nc, a.begin() + (size_t)dst * G * nc); };
    cp(gv, 3); cp(gv_tmp, 3); cp(gdv, 1); cp(gp, 1); cp(gq, c.q_dim);
  }
  void reset_grad_till_frame(int s) {  // SF:185-188
    auto z = [&](std::vector<R>& a, int nc) { std::fill(a.begin(), a.begin() + (size_t)s * G * nc, R(0)); };
    z(gv, 3); z(gv_tmp, 3); z(gdv, 1); z(gp, 1); z(gq, c.q_dim);
  }
};

}  // namespace orc
