/*============================================================================
 * oracle/cs_oracle_rhs.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the right-hand-side operator around the pressure solve
 * (SURVEY.md section 8f item 2): cs_diffusion_potential
 * (src/alge/cs_convection_diffusion.cpp:8743-8990) WITHOUT flux reconstruction
 * (eqp->ircflu = 0, :8869-8889) -- the divergence of the diffusive flux of a cell
 * variable: interior faces in face order, then boundary faces in face order with
 * flux_diff = inc*af + bf*pvar (cs_boundary_conditions_set_coeffs.cpp:172-185).
 * On an orthogonal mesh (II' = JJ' = 0) the reconstructed branch (:8895-8972) adds
 * exact zeros, so this is also its result there.  The gradient reconstruction of
 * non-orthogonal meshes (cs_gradient.cpp) is not restated.
 *
 * Pinned bit for bit against oracle/_ref/libcsref.so (csref_diffusion_potential
 * drives the reference's real function), tests/test_oracle.py.
 *============================================================================*/

#include "cs_oracle.h"

void
ora_diffusion_potential(int n_cells, int n_i_faces, const int *i_face_cells, int n_b_faces,
                        const int *b_face_cells, int init, int inc, const double *pvar,
                        const double *cofaf, const double *cofbf, const double *i_visc,
                        const double *b_visc, double *diverg)
{
  if (init >= 1)
    for (int i = 0; i < n_cells; i++) diverg[i] = 0.;
  for (int f = 0; f < n_i_faces; f++) {
    const int c0 = i_face_cells[2*f], c1 = i_face_cells[2*f+1];
    const double flux = i_visc[f]*(pvar[c0] - pvar[c1]);
    if (c0 < n_cells) diverg[c0] += flux;
    if (c1 < n_cells) diverg[c1] += -flux;
  }
  for (int f = 0; f < n_b_faces; f++) {
    const int c0 = b_face_cells[f];
    const double flux_d = inc*cofaf[f] + cofbf[f]*pvar[c0];
    diverg[c0] += b_visc[f]*flux_d;
  }
}
