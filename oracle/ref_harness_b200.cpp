/*============================================================================
 * oracle/ref_harness_b200.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * Drives integration/cs_sles_b200.cpp (the reference-side binding of
 * libcsb200.so) through the reference's REAL solver registry:
 *   cs_sles_b200_define -> cs_sles_define     (src/alge/cs_sles.cpp:1385-1425)
 *   cs_sles_solve                             (src/alge/cs_sles.cpp:1726-1860)
 *   cs_sles_free / cs_sles_finalize           (:1883, :981)
 * on a cs_matrix_t the reference built itself (csref_system_create in
 * ref_harness.cpp: cs_matrix_create + cs_matrix_transfer_coefficients_msr +
 * cs_matrix_set_mesh_association).  Linked with the same reference objects into
 * oracle/_ref/libcsref_b200.so, which also depends on libcsb200.so -- it is
 * only loaded by the `-m gpu` test tests/test_gpu_shim.py.
 *============================================================================*/

#include "base/cs_defs.h"

#include <cstring>

#include "alge/cs_matrix.h"
#include "alge/cs_sles.h"

#include "cs_sles_b200.h"

extern "C" void *csref_system_matrix(void *sys);        /* ref_harness.cpp */

extern "C" {

/* Solve sys through the registry with a B200 solver of the given type, n_solves times
   (zero start each time unless use_vx_ini); between solves `touch` != 0 re-defines nothing
   but calls cs_sles_free, so the next solve goes through setup again.
   out: [0] n_iter, [1] residual, [2] state, [3] uploads, [4] value updates, [5] n_calls */
int
csref_b200_sles_solve(void *sys, int type, int n_max_iter, double precision, double r_norm,
                      const double *rhs, double *vx, int use_vx_ini, int n_solves,
                      int free_between, double *out)
{
  const cs_matrix_t *a = static_cast<const cs_matrix_t *>(csref_system_matrix(sys));
  static int n_defs = 0;
  char name[64];
  snprintf(name, sizeof(name), "b200_system_%d", n_defs++);
  cs_sles_t *sc = cs_sles_b200_define(-1, name, (cs_sles_b200_type_t)type, n_max_iter);
  /* what any caller of the registry does next: find the system by name, solve it */
  cs_sles_t *found = cs_sles_find(-1, name);
  if (found != sc) return -100;
  const cs_lnum_t n_cols = cs_matrix_get_n_columns(a);
  int n_iter = 0;
  double residual = -1.;
  cs_sles_convergence_state_t state = CS_SLES_ITERATING;
  for (int i = 0; i < n_solves; i++) {
    if (!use_vx_ini) memset(vx, 0, sizeof(double)*n_cols);
    state = cs_sles_solve(found, a, precision, r_norm, use_vx_ini ? false : true, &n_iter,
                          &residual, rhs, vx, 0, nullptr);
    if (free_between && i + 1 < n_solves) cs_sles_free(found);
  }
  int counts[2] = {0, 0};
  cs_sles_b200_get_transfer_counts(cs_sles_get_context(found), counts);
  out[0] = n_iter; out[1] = residual; out[2] = (double)state;
  out[3] = counts[0]; out[4] = counts[1];
  cs_sles_free(found);
  return (int)state;
}

/* cs_sles_finalize: destroys every registered system through its destroy callback */
void
csref_b200_finalize(void)
{
  cs_sles_finalize();
}

}
