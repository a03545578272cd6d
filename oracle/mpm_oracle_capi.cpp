// oracle/mpm_oracle_capi.cpp — TEST INFRASTRUCTURE ONLY. C ABI over mpm_oracle.hpp so pytest (ctypes)
// and bench.py's cpu_baseline leg can drive the CPU restatement. All arrays cross the ABI as
// double (fp32 values are exactly representable); `precision` picks the internal arithmetic.
#include "mpm_oracle.hpp"
#include <omp.h>

using namespace orc;

struct Handle {
  int precision;
  Sim<float>* f32 = nullptr;
  Sim<double>* f64 = nullptr;
};

#define DISPATCH(h, expr)                         \
  do {                                            \
    if ((h)->precision == 32) { auto& S = *(h)->f32; expr; } \
    else { auto& S = *(h)->f64; expr; }           \
  } while (0)

template <class R> static void set_frame_t(Sim<R>& S, int f, const double* x, const double* v, const double* C, const double* F, const int* used) {
  size_t o = (size_t)f * S.N;
  for (size_t i = 0; i < (size_t)S.N * 3; i++) { S.x[o * 3 + i] = (R)x[i]; S.v[o * 3 + i] = (R)v[i]; }
  for (size_t i = 0; i < (size_t)S.N * 9; i++) { S.C[o * 9 + i] = (R)C[i]; S.F[o * 9 + i] = (R)F[i]; }
  for (int i = 0; i < S.N; i++) S.used[o + i] = used[i];
}
template <class R> static void get_frame_t(Sim<R>& S, int f, double* x, double* v, double* C, double* F, int* used) {
  size_t o = (size_t)f * S.N;
  for (size_t i = 0; i < (size_t)S.N * 3; i++) { x[i] = S.x[o * 3 + i]; v[i] = S.v[o * 3 + i]; }
  for (size_t i = 0; i < (size_t)S.N * 9; i++) { C[i] = S.C[o * 9 + i]; F[i] = S.F[o * 9 + i]; }
  for (int i = 0; i < S.N; i++) used[i] = S.used[o + i];
}
template <class R> static void set_grad_t(Sim<R>& S, int f, const double* x, const double* v, const double* C, const double* F) {
  size_t o = (size_t)f * S.N;
  for (size_t i = 0; i < (size_t)S.N * 3; i++) { S.gx[o * 3 + i] = (R)x[i]; S.gv[o * 3 + i] = (R)v[i]; }
  for (size_t i = 0; i < (size_t)S.N * 9; i++) { S.gC[o * 9 + i] = (R)C[i]; S.gF[o * 9 + i] = (R)F[i]; }
}
template <class R> static void get_grad_t(Sim<R>& S, int f, double* x, double* v, double* C, double* F) {
  size_t o = (size_t)f * S.N;
  for (size_t i = 0; i < (size_t)S.N * 3; i++) { x[i] = S.gx[o * 3 + i]; v[i] = S.gv[o * 3 + i]; }
  for (size_t i = 0; i < (size_t)S.N * 9; i++) { C[i] = S.gC[o * 9 + i]; F[i] = S.gF[o * 9 + i]; }
}
template <class R> static void get_grid_t(Sim<R>& S, double* vin, double* m, double* vout) {
  for (size_t i = 0; i < (size_t)S.G * 3; i++) { vin[i] = S.g_vin[i]; vout[i] = S.g_vout[i]; }
  for (size_t i = 0; i < (size_t)S.G; i++) m[i] = S.g_m[i];
}
template <class R> static void set_grid_t(Sim<R>& S, const double* vin, const double* m) {
  for (size_t i = 0; i < (size_t)S.G * 3; i++) S.g_vin[i] = (R)vin[i];
  for (size_t i = 0; i < (size_t)S.G; i++) S.g_m[i] = (R)m[i];
}
template <class R> static void get_grid_grad_t(Sim<R>& S, double* vin, double* m, double* vout) {
  for (size_t i = 0; i < (size_t)S.G * 3; i++) { vin[i] = S.gg_vin[i]; vout[i] = S.gg_vout[i]; }
  for (size_t i = 0; i < (size_t)S.G; i++) m[i] = S.gg_m[i];
}
template <class R> static void set_grid_grad_t(Sim<R>& S, const double* vin, const double* m, const double* vout) {
  for (size_t i = 0; i < (size_t)S.G * 3; i++) { S.gg_vin[i] = (R)vin[i]; S.gg_vout[i] = (R)vout[i]; }
  for (size_t i = 0; i < (size_t)S.G; i++) S.gg_m[i] = (R)m[i];
}
template <class R> static void set_info_t(Sim<R>& S, const int* mat, const int* cls, const double* mu, const double* lam, const double* mass) {
  for (int i = 0; i < S.N; i++) { S.mat[i] = mat[i]; S.cls[i] = cls[i]; S.mu[i] = (R)mu[i]; S.lam[i] = (R)lam[i]; S.mass[i] = (R)mass[i]; }
}
template <class R> static int add_effector_t(Sim<R>& S, const EffectorCfg* cfg, const double* random_vector, const int* act_range, int n_act_range) {
  Effector<R> e; e.init(*cfg, S.T);
  if (cfg->type != 0) {
    size_t n = (size_t)cfg->random_length * cfg->flux * 3;
    e.random_vector.resize(n);
    for (size_t i = 0; i < n; i++) e.random_vector[i] = (R)random_vector[i];
    e.act_range.assign(act_range, act_range + n_act_range);
    e.act_id[0] = 0;
  }
  S.eff.push_back(e);
  return (int)S.eff.size() - 1;
}
template <class R> static void eff_set_state_t(Sim<R>& S, int ei, int f, const double* st) {
  auto& e = S.eff[ei];
  for (int k = 0; k < 3; k++) e.pos[f * 3 + k] = (R)st[k];
  for (int k = 0; k < 4; k++) e.quat[f * 4 + k] = (R)st[3 + k];
  if (e.cfg.type != 0) e.act_id[f] = (int)st[7];
}
template <class R> static void eff_get_state_t(Sim<R>& S, int ei, int f, double* st) {
  auto& e = S.eff[ei];
  for (int k = 0; k < 3; k++) st[k] = e.pos[f * 3 + k];
  for (int k = 0; k < 4; k++) st[3 + k] = e.quat[f * 4 + k];
  st[7] = (double)e.act_id[f];
}
template <class R> static void eff_get_grad_t(Sim<R>& S, int ei, int n, double* out) {  // effector.py:276-283
  auto& e = S.eff[ei];
  const int ad = e.cfg.action_dim;
  for (int i = 0; i < n; i++) for (int j = 0; j < ad; j++) out[(size_t)i * ad + j] = e.gact[(size_t)i * ad + j];
  for (int j = 0; j < ad; j++) out[(size_t)n * ad + j] = e.gact_p[j];
}
template <class R> static void eff_get_pose_grad_t(Sim<R>& S, int ei, int f, double* out) {
  auto& e = S.eff[ei];
  for (int k = 0; k < 3; k++) { out[k] = e.gpos[f * 3 + k]; out[3 + k] = e.gv[f * 3 + k]; }
}
template <class R> static void set_action_t(Sim<R>& S, int ei, int s, int sg, const double* a) {
  R aa[6]; for (int k = 0; k < S.eff[ei].cfg.action_dim; k++) aa[k] = (R)a[k];
  S.set_action(ei, s, sg, aa);
}
template <class R> static void apply_action_p_t(Sim<R>& S, int ei, const double* a) {
  R aa[6]; for (int k = 0; k < S.eff[ei].cfg.action_dim; k++) aa[k] = (R)a[k];
  S.apply_action_p(ei, aa);
}
template <class R> static void fill_mesh(SdfMesh<R>& M, int res, const double* vox, const double* T, double friction, double softness) {
  M.res = res; M.vox.resize((size_t)res * res * res);
  for (size_t i = 0; i < M.vox.size(); i++) M.vox[i] = (R)vox[i];
  for (int i = 0; i < 16; i++) M.T[i] = (R)T[i];
  // inverse of the 3x3 block, evaluated in R like `T[:3,:3].inverse()` in the kernels (static.py:59)
  M3<R> A; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) A[r][c] = M.T[r * 4 + c];
  R d = det3(A);
  R inv[9] = {(A[1][1] * A[2][2] - A[1][2] * A[2][1]) / d, (A[0][2] * A[2][1] - A[0][1] * A[2][2]) / d, (A[0][1] * A[1][2] - A[0][2] * A[1][1]) / d,
              (A[1][2] * A[2][0] - A[1][0] * A[2][2]) / d, (A[0][0] * A[2][2] - A[0][2] * A[2][0]) / d, (A[0][2] * A[1][0] - A[0][0] * A[1][2]) / d,
              (A[1][0] * A[2][1] - A[1][1] * A[2][0]) / d, (A[0][1] * A[2][0] - A[0][0] * A[2][1]) / d, (A[0][0] * A[1][1] - A[0][1] * A[1][0]) / d};
  for (int i = 0; i < 9; i++) M.Ainv[i] = inv[i];
  M.friction = (R)friction; M.softness = (R)softness; M.has_dynamics = 1;
}
template <class R> static void add_static_t(Sim<R>& S, int res, const double* vox, const double* T, double friction) {
  SdfMesh<R> M; fill_mesh(M, res, vox, T, friction, 0.0); S.statics.push_back(M);
}
template <class R> static void set_rigid_t(Sim<R>& S, int res, const double* vox, const double* T, double friction, double softness, int collide_type) {
  fill_mesh(S.rigid_mesh, res, vox, T, friction, softness); S.has_rigid = true; S.agent_type = 1; S.collide_type = collide_type;
}
template <class R> static void svd_t(const double* A, double* U, double* s, double* V) {
  M3<R> a, u, v; R sg[3];
  for (int i = 0; i < 9; i++) (&a.a[0][0])[i] = (R)A[i];
  svd3(a, u, sg, v);
  for (int i = 0; i < 9; i++) { U[i] = (&u.a[0][0])[i]; V[i] = (&v.a[0][0])[i]; }
  for (int i = 0; i < 3; i++) s[i] = sg[i];
}

extern "C" {

void* orc_create(const Config* cfg, int precision) {
  Handle* h = new Handle();
  h->precision = precision;
  if (precision == 32) h->f32 = new Sim<float>(*cfg); else h->f64 = new Sim<double>(*cfg);
  return h;
}
void orc_destroy(void* hp) { Handle* h = (Handle*)hp; delete h->f32; delete h->f64; delete h; }
void orc_set_threads(int n) { omp_set_num_threads(n); }
int orc_get_max_threads() { return omp_get_max_threads(); }

void orc_set_collector(void* hp, int type, const double* lo, const double* hi, const double* c, double r, int mat) {
  Handle* h = (Handle*)hp;
  DISPATCH(h, (S.collector_type = type, S.collector_mat = mat, S.col_r = (decltype(S.dt))r, S.col_c[0] = (decltype(S.dt))c[0], S.col_c[1] = (decltype(S.dt))c[1]));
  for (int k = 0; k < 3; k++) DISPATCH(h, (S.col_lo[k] = (decltype(S.dt))lo[k], S.col_hi[k] = (decltype(S.dt))hi[k]));
}
void orc_set_bodies(void* hp, const int* body_id, int n_bodies) { Handle* h = (Handle*)hp; DISPATCH(h, S.set_bodies(body_id, n_bodies)); }
void orc_set_particle_info(void* hp, const int* mat, const int* cls, const double* mu, const double* lam, const double* mass) {
  Handle* h = (Handle*)hp; DISPATCH(h, set_info_t(S, mat, cls, mu, lam, mass)); }
void orc_set_frame(void* hp, int f, const double* x, const double* v, const double* C, const double* F, const int* used) {
  Handle* h = (Handle*)hp; DISPATCH(h, set_frame_t(S, f, x, v, C, F, used)); }
void orc_get_frame(void* hp, int f, double* x, double* v, double* C, double* F, int* used) {
  Handle* h = (Handle*)hp; DISPATCH(h, get_frame_t(S, f, x, v, C, F, used)); }
void orc_set_grad_frame(void* hp, int f, const double* x, const double* v, const double* C, const double* F) {
  Handle* h = (Handle*)hp; DISPATCH(h, set_grad_t(S, f, x, v, C, F)); }
void orc_get_grad_frame(void* hp, int f, double* x, double* v, double* C, double* F) {
  Handle* h = (Handle*)hp; DISPATCH(h, get_grad_t(S, f, x, v, C, F)); }
void orc_get_grid(void* hp, double* vin, double* m, double* vout) { Handle* h = (Handle*)hp; DISPATCH(h, get_grid_t(S, vin, m, vout)); }
void orc_set_grid(void* hp, const double* vin, const double* m) { Handle* h = (Handle*)hp; DISPATCH(h, set_grid_t(S, vin, m)); }
void orc_get_grid_grad(void* hp, double* vin, double* m, double* vout) { Handle* h = (Handle*)hp; DISPATCH(h, get_grid_grad_t(S, vin, m, vout)); }
void orc_set_grid_grad(void* hp, const double* vin, const double* m, const double* vout) { Handle* h = (Handle*)hp; DISPATCH(h, set_grid_grad_t(S, vin, m, vout)); }

void orc_substep(void* hp, int f, int f_global, int none_action) { Handle* h = (Handle*)hp; DISPATCH(h, S.substep(f, f_global, none_action != 0)); }
void orc_substep_grad(void* hp, int f, int f_global, int none_action) { Handle* h = (Handle*)hp; DISPATCH(h, S.substep_grad(f, f_global, none_action != 0)); }
// phase-level entry points (parity tests drive the CUDA phases one by one)
void orc_phase_reset_grid(void* hp) { Handle* h = (Handle*)hp; DISPATCH(h, S.reset_grid()); }
void orc_phase_p2g(void* hp, int f, int write_F) { Handle* h = (Handle*)hp; DISPATCH(h, (S.compute_F_tmp_svd(f), S.p2g(f, write_F != 0))); }
void orc_phase_grid_op(void* hp, int f) { Handle* h = (Handle*)hp; DISPATCH(h, S.grid_op(f)); }
void orc_phase_g2p(void* hp, int f) { Handle* h = (Handle*)hp; DISPATCH(h, (S.advect_used(f), S.process_unused(f), S.g2p(f), S.advect(f))); }
void orc_phase_g2p_grad(void* hp, int f) { Handle* h = (Handle*)hp; DISPATCH(h, S.g2p_advect_grad(f)); }
void orc_phase_grid_op_grad(void* hp, int f) { Handle* h = (Handle*)hp; DISPATCH(h, S.grid_op_grad(f)); }
void orc_phase_unused_grad(void* hp, int f) { Handle* h = (Handle*)hp; DISPATCH(h, S.process_unused_grad(f)); }
void orc_phase_p2g_grad(void* hp, int f) { Handle* h = (Handle*)hp; DISPATCH(h, S.p2g_grad(f)); }

void orc_copy_frame(void* hp, int s, int t) { Handle* h = (Handle*)hp; DISPATCH(h, S.copy_frame(s, t)); }
void orc_copy_grad(void* hp, int s, int t) { Handle* h = (Handle*)hp; DISPATCH(h, S.copy_grad(s, t)); }
void orc_reset_grad_till(void* hp, int f) { Handle* h = (Handle*)hp; DISPATCH(h, S.reset_grad_till(f)); }
void orc_reset_grad(void* hp) { Handle* h = (Handle*)hp; DISPATCH(h, S.reset_grad()); }

double orc_loss_value(void* hp, int f, int mat, double w, const double* tgt) { Handle* h = (Handle*)hp; double r = 0; DISPATCH(h, r = S.loss_value(f, mat, w, tgt)); return r; }
void orc_loss_seed(void* hp, int f, int mat, double w, const double* tgt) { Handle* h = (Handle*)hp; DISPATCH(h, S.loss_seed(f, mat, w, tgt)); }

void orc_add_static(void* hp, int res, const double* vox, const double* T, double friction) { Handle* h = (Handle*)hp; DISPATCH(h, add_static_t(S, res, vox, T, friction)); }
void orc_set_rigid_mesh(void* hp, int res, const double* vox, const double* T, double friction, double softness, int collide_type) {
  Handle* h = (Handle*)hp; DISPATCH(h, set_rigid_t(S, res, vox, T, friction, softness, collide_type)); }
// unit access to one collide evaluation (tests): io = [p(3) v(3) pos0(3) pos1(3)] -> out(3); with gout != null also the adjoints [gp gv gpos0 gpos1]
void orc_sdf_collide_eval(int res, const double* vox, const double* T, double friction, double softness, int dynamic, double dt,
                          const double* io, double* out, const double* gout, double* gio) {
  SdfMesh<double> M; fill_mesh(M, res, vox, T, friction, softness);
  const double q[4] = {1, 0, 0, 0};
  double gp[3] = {0, 0, 0}, gv[3] = {0, 0, 0}, g0[3] = {0, 0, 0}, g1[3] = {0, 0, 0};
  sdf_collide<double>(M, dynamic != 0, io + 6, q, io + 9, q, dt, io, io + 3, out, gout, gv, gp, g0, g1);
  if (gout) for (int k = 0; k < 3; k++) { gio[k] = gp[k]; gio[3 + k] = gv[k]; gio[6 + k] = g0[k]; gio[9 + k] = g1[k]; }
}
// same with free pose quaternions: io = [p(3) v(3) pos0(3) pos1(3) q0(4) q1(4)]; adjoints in the same layout
void orc_sdf_collide_eval_q(int res, const double* vox, const double* T, double friction, double softness, double dt,
                            const double* io, double* out, const double* gout, double* gio) {
  SdfMesh<double> M; fill_mesh(M, res, vox, T, friction, softness);
  double gp[3] = {0, 0, 0}, gv[3] = {0, 0, 0}, g0[3] = {0, 0, 0}, g1[3] = {0, 0, 0}, gq0[4] = {0, 0, 0, 0}, gq1[4] = {0, 0, 0, 0};
  sdf_collide<double>(M, true, io + 6, io + 12, io + 9, io + 16, dt, io, io + 3, out, gout, gv, gp, g0, g1, gq0, gq1);
  if (gout) {
    for (int k = 0; k < 3; k++) { gio[k] = gp[k]; gio[3 + k] = gv[k]; gio[6 + k] = g0[k]; gio[9 + k] = g1[k]; }
    for (int k = 0; k < 4; k++) { gio[12 + k] = gq0[k]; gio[16 + k] = gq1[k]; }
  }
}
void orc_set_collide_y_min(void* hp, double y) { Handle* h = (Handle*)hp; DISPATCH(h, S.collide_y_min = (decltype(S.dt))y); }
void orc_set_agent(void* hp, int agent_type) { Handle* h = (Handle*)hp; DISPATCH(h, (S.agent_type = agent_type, S.has_injector = (agent_type == 2))); }
void orc_set_agent_layout(void* hp, int inj_idx, int rigid_idx, int inject_till, int has_injector) {
  Handle* h = (Handle*)hp; DISPATCH(h, (S.inj_idx = inj_idx, S.rigid_idx = rigid_idx, S.inject_till = inject_till, S.has_injector = has_injector != 0)); }
int orc_add_effector(void* hp, const EffectorCfg* cfg, const double* random_vector, const int* act_range, int n_act_range) {
  Handle* h = (Handle*)hp; int r = -1; DISPATCH(h, r = add_effector_t(S, cfg, random_vector, act_range, n_act_range)); return r; }
void orc_effector_set_state(void* hp, int ei, int f, const double* st) { Handle* h = (Handle*)hp; DISPATCH(h, eff_set_state_t(S, ei, f, st)); }
void orc_effector_get_state(void* hp, int ei, int f, double* st) { Handle* h = (Handle*)hp; DISPATCH(h, eff_get_state_t(S, ei, f, st)); }
void orc_effector_set_action(void* hp, int ei, int s, int sg, const double* a) { Handle* h = (Handle*)hp; DISPATCH(h, set_action_t(S, ei, s, sg, a)); }
void orc_effector_set_action_grad(void* hp, int ei, int s, int sg) { Handle* h = (Handle*)hp; DISPATCH(h, S.set_action_grad(ei, s, sg)); }
void orc_effector_apply_action_p(void* hp, int ei, const double* a) { Handle* h = (Handle*)hp; DISPATCH(h, apply_action_p_t(S, ei, a)); }
void orc_effector_apply_action_p_grad(void* hp, int ei) { Handle* h = (Handle*)hp; DISPATCH(h, S.apply_action_p_grad(ei)); }
void orc_effector_get_action_grad(void* hp, int ei, int n, double* out) { Handle* h = (Handle*)hp; DISPATCH(h, eff_get_grad_t(S, ei, n, out)); }
void orc_effector_get_pose_grad(void* hp, int ei, int f, double* out) { Handle* h = (Handle*)hp; DISPATCH(h, eff_get_pose_grad_t(S, ei, f, out)); }

// unit access to the manual SVD adjoint (MPM:272-292): gU, gS (diagonal matrix), gV, U, sig[3], V -> gA (row-major 3x3), double precision
void orc_backward_svd(const double* gU, const double* gS, const double* gV, const double* U, const double* sig, const double* V, double* out) {
  M3<double> a, b, c, u, v;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { a[i][j] = gU[i * 3 + j]; b[i][j] = gS[i * 3 + j]; c[i][j] = gV[i * 3 + j]; u[i][j] = U[i * 3 + j]; v[i][j] = V[i * 3 + j]; }
  M3<double> r = backward_svd<double>(a, b, c, u, sig, v);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[i * 3 + j] = r[i][j];
}
void orc_svd3(const double* A, double* U, double* s, double* V, int precision) {
  if (precision == 32) svd_t<float>(A, U, s, V); else svd_t<double>(A, U, s, V);
}

}  // extern "C"
