#ifndef CS_SLES_B200_H
#define CS_SLES_B200_H

/*============================================================================
 * cs_sles_b200.h -- reference-side binding of libcsb200.so (the file a
 * code_saturne maintainer adds as src/alge/cs_sles_b200.h).
 *
 * Registers a B200 solver for one linear system through the reference's own
 * plug-in interface, cs_sles_define (src/alge/cs_sles.cpp:1385-1425), exactly
 * like cs_sles_amgx / cs_sles_hypre / cs_sles_petsc do
 * (src/alge/cs_sles_amgx.cpp:592-612).  Callers (cs_sles_solve,
 * cs_sles_solve_native, the physics modules) are unchanged.
 *============================================================================*/

#include "alge/cs_sles.h"

typedef enum {
  CS_SLES_B200_MULTIGRID,   /* cs_multigrid V-cycle in solver mode (cs_multigrid_define) */
  CS_SLES_B200_FCG_MG,      /* cs_sles_it FCG + cs_multigrid_pc: the default pressure solver */
  CS_SLES_B200_PCG,         /* cs_sles_it PCG + Jacobi */
  CS_SLES_B200_FCG          /* cs_sles_it FCG + Jacobi */
} cs_sles_b200_type_t;

/* Define the solver of system (f_id, name); n_max_iter: Krylov iterations (cycles for
   CS_SLES_B200_MULTIGRID).  Returns the cs_sles_t the registry holds.  Call from
   cs_user_linear_solvers(), or install for every system with
   cs_sles_set_default_define (src/alge/cs_sles.cpp:2002). */
cs_sles_t *
cs_sles_b200_define(int                  f_id,
                    const char          *name,
                    cs_sles_b200_type_t  type,
                    int                  n_max_iter);

/* Number of matrix uploads / value-only updates done so far by a context (tests). */
void
cs_sles_b200_get_transfer_counts(const void *context, int counts[2]);

#endif
