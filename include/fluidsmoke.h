/* fluidsmoke.h — C ABI of the B200 smoke solver (same shared library as fluidmpm.h: libfluidmpm.so).
 *
 * Replaces the Taichi kernels of fluidlab/fluidengine/simulators/smoke_field.py (abbrev. SF) — SURVEY.md §8(f) rank 3: the Eulerian
 * solver that `MPMSimulator.step_` / `step_grad` run once per STEP (mpm_simulator.py:744-747, 765-767) for the air-circulation task
 * (envs/circulation_env.py).  One step = free-space mask (SF:190-201), RK3 semi-Lagrangian advection + air-conditioner impulse
 * (SF:203-233), divergence (SF:235-261), `solver_iters` Jacobi sweeps of the pressure problem (SF:97-106, 135-146), projection
 * (SF:275-289); fsmk_step_grad is the hand-written adjoint of SF:112-127 (where the reference calls Taichi autodiff).
 *
 * Conventions as in fluidmpm.h: plain pointers and sizes, no torch types; every pointer is DEVICE memory owned by the caller
 * (the Python host uses torch tensors as the allocator); calls are asynchronous on `stream`; 0 = success, otherwise
 * fsmk_last_error(); one handle per GPU, not thread-safe per handle.
 *
 * Layouts (n = res, G = n^3, cell (i, j, k) -> (i*n + j)*n + k, S + 1 step frames):
 *   v, v_tmp   float4[S+1][G]   (x, y, z, unused)          div, p   float[S+1][G]
 *   q          float[S+1][q_dim][G]                        is_free  unsigned char[S+1][G]
 *   gradients  same shapes (gv, gv_tmp, gdiv, gp, gq);   tmp_a, tmp_b, acc: float[G] scratch of the pressure solver.
 */
#ifndef FLUIDSMOKE_H
#define FLUIDSMOKE_H
#include "fluidmpm.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct FsmkHandle FsmkHandle;

typedef struct {
  int res;              /* SF:14 (128) */
  int max_steps_local;  /* S: frames 0..S (SF:74) */
  int q_dim;            /* SF:14; 1..4 */
  int solver_iters;     /* SF:14 */
  float dt;             /* SF:14 (0.03) */
  int lower_y, higher_y;/* SF:26-27: only cells with lower_y < j < higher_y can be free */
  float low_T;          /* SF:25: temperature blown by the air conditioner */
  float inject_v[3];    /* effectors/aircon.py:14 */
  int device;
} FsmkConfig;

typedef struct {
  void *v, *v_tmp, *div, *p, *q, *is_free;
  void *gv, *gv_tmp, *gdiv, *gp, *gq;     /* NULL: forward only */
  void *tmp_a, *tmp_b, *acc;
} FsmkBuffers;

/* the air conditioner (effectors/aircon.py:18-26): per-SUBSTEP arrays of the MPM ring, indexed by f (SF:217-221) */
typedef struct {
  const void *pos, *quat;   /* float[(T+1)*3], float[(T+1)*4] */
  const void *s, *r;        /* float[T+1] strength and radius */
  void *gpos, *gquat, *gs, *gr;   /* adjoints (accumulated); NULL when gradients are never taken */
} FsmkAircon;

int  fsmk_create(const FsmkConfig* cfg, FsmkHandle** out);
void fsmk_destroy(FsmkHandle* h);
int  fsmk_bind(FsmkHandle* h, const FsmkBuffers* b);
int  fsmk_set_statics(FsmkHandle* h, int n_statics, const FmpmSdfMesh* statics);   /* Static.is_collide, meshes/static.py:106-114; at most 4 */
int  fsmk_set_aircon(FsmkHandle* h, const FsmkAircon* a);
const char* fsmk_last_error(FsmkHandle* h);

int fsmk_step(FsmkHandle* h, int s, int f, void* stream);        /* SF:95-110 (colorize is renderer-only) */
int fsmk_step_grad(FsmkHandle* h, int s, int f, void* stream);   /* SF:112-127 */

/* phase-level entry points (parity tests, ncu) */
int fsmk_free_space(FsmkHandle* h, int s, void* stream);               /* SF:190-201 */
int fsmk_advect(FsmkHandle* h, int s, int f, void* stream);            /* SF:203-233; also writes v[s+1] of the non-free cells (SF:288-289) */
int fsmk_divergence(FsmkHandle* h, int s, void* stream);               /* SF:235-261 */
int fsmk_pressure(FsmkHandle* h, int s, void* stream);                 /* SF:97-106: p[s] -> solver_iters Jacobi sweeps -> p[s+1] */
int fsmk_project(FsmkHandle* h, int s, void* stream);                  /* SF:275-287 */
int fsmk_project_grad(FsmkHandle* h, int s, void* stream);
int fsmk_pressure_grad(FsmkHandle* h, int s, void* stream);
int fsmk_divergence_grad(FsmkHandle* h, int s, void* stream);
int fsmk_advect_grad(FsmkHandle* h, int s, int f, void* stream);

#ifdef __cplusplus
}
#endif
#endif
