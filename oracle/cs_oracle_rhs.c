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
 * exact zeros, so this is also its result there.
 *
 * ora_gradient_lsq / ora_diffusion_potential_rc: the reconstructed branch and the cell
 * gradient it needs -- least squares over the face neighbours (imrgra = 1, CS_GRADIENT_LSQ,
 * standard neighbourhood, no weighting, no limiter): _lsq_scalar_gradient
 * (src/alge/cs_gradient.cpp:2695-2900) with the cocg matrices of _compute_cell_cocg_lsq
 * (:2501-2680), which the reference keeps in SINGLE precision (cs_cocg_t = float,
 * cs_gradient_priv.h:54) and inverts with the float instance of
 * cs_math_matrix_gauss_inverse (src/base/cs_math.h:542-590): restated as such.
 *
 * The *_b forms take II' of the boundary faces (cs_mesh_quantities_t::diipb): the value at I'
 * of every boundary face is reconstructed by cs_gradient_boundary_iprime_lsq_s
 * (src/alge/cs_gradient_boundary.cpp:197-640; no limiter, no hydrostatic term, no weighting),
 * the face values follow _compute_bc_flux_grad<true>
 * (src/base/cs_boundary_conditions_set_coeffs.cpp:125-195: val_f = inc*a + b*val_ip,
 * flux_diff = inc*af + bf*val_ip), and the boundary directions of the cell gradient become
 * n + II'/b_dist (cs_gradient.cpp:2632-2648, :2875-2900).  With diipb NULL they are the plain
 * forms (II' = 0: val_ip = pvar of the cell).
 *
 * Pinned bit for bit against oracle/_ref/libcsref.so (csref_diffusion_potential
 * drives the reference's real function), tests/test_oracle.py.
 *============================================================================*/

#include <math.h>
#include <stdlib.h>

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

/* cs_math_6_gauss_inverse_in_place<float> (cs_math.h:598-615, :542-590) */
static void
_gauss_inverse_6f(float a[6])
{
  float t[3][6] = {{a[0], a[3], a[5], 1, 0, 0},
                   {a[3], a[1], a[4], 0, 1, 0},
                   {a[5], a[4], a[2], 0, 0, 1}};
  float *p_a[3] = {t[0], t[1], t[2]};
  for (int i = 0; i < 3; i++) {
    for (int j = i+1; j < 3; j++) {
      if (fabsf(p_a[i][i]) < fabsf(p_a[j][i])) {
        float *tmp = p_a[j]; p_a[j] = p_a[i]; p_a[i] = tmp;
      }
    }
    for (int j = 0; j < 3; j++) {
      if (j != i) {
        const float s = p_a[j][i] / p_a[i][i];
        for (int k = 0; k < 6; k++) {
          const float prod = p_a[i][k] * s;
          p_a[j][k] = p_a[j][k] - prod;
        }
      }
    }
  }
  float b[3][3];
  for (int i = 0; i < 3; i++) {
    const float s = 1.f / p_a[i][i];
    for (int j = 0; j < 3; j++) b[i][j] = p_a[i][3+j] * s;
  }
  a[0] = b[0][0]; a[1] = b[1][1]; a[2] = b[2][2];
  a[3] = b[0][1]; a[4] = b[1][2]; a[5] = b[0][2];
}

/* simple limiter of the value at I' (cs_gradient_boundary.cpp:612-628): kept between the
   smallest and largest of the cell's, its neighbours' and its boundary faces' unreconstructed
   values, the bounds stretched around the cell value by clip_coeff; none if clip_coeff < 0
   (b_rc_clip_factor, on with eqp->imligr >= 0) */
static double
_clip_ip(double var_ip, double var_i, double var_min, double var_max, double clip_coeff)
{
  if (clip_coeff >= 0) {
    if (clip_coeff < 1. || clip_coeff > 1.) {
      const double d_min = var_min - var_i;
      const double d_max = var_max - var_i;
      var_min = var_i + d_min*clip_coeff;
      var_max = var_i + d_max*clip_coeff;
    }
    if (var_ip < var_min)
      var_ip = var_min;
    if (var_ip > var_max)
      var_ip = var_max;
  }
  return var_ip;
}

/* direction of a boundary face in the least-squares system, B_DIRECTION_LSQ = CS_IPRIME_F_LSQ
   (cs_gradient.h:55): n + II'/b_dist (cs_gradient.cpp:2632-2636, :2880-2885) */
static void
_b_direction(int f, const double *b_face_u_normal, const double *b_dist, const double *diipb,
             double d[3])
{
  if (diipb == NULL) {
    for (int ll = 0; ll < 3; ll++) d[ll] = b_face_u_normal[3*(size_t)f+ll];
    return;
  }
  const double ddif = 1. / b_dist[f];
  for (int ll = 0; ll < 3; ll++)
    d[ll] = b_face_u_normal[3*(size_t)f+ll] + ddif*diipb[3*(size_t)f+ll];
}

/* cs_gradient_boundary_iprime_lsq_s (src/alge/cs_gradient_boundary.cpp:197-640) with
   its simple limiter (clip_coeff; < 0: none, as with eqp->imligr < 0), hyd_p_flag 0, c_weight NULL, standard
   neighbourhood: value of pvar at I' of every boundary face from a least-squares gradient local
   to the face's cell -- neighbours in adjacency order, then the cell's boundary faces in
   ascending order with the BC-aware direction n + (1 - b) II'/b_dist; cocg accumulated in
   SINGLE precision (cs_cocg_t), its cofactors computed in float, the solve in double (:560-582).
   Note coefa enters WITHOUT inc (:505, :541). */
void
ora_b_iprime_lsq(int n_cells, const int *row_index, const int *col_id, int n_b_faces,
                 const int *b_face_cells, const double *cell_cen,
                 const double *b_face_u_normal, const double *b_dist, const double *diipb,
                 double clip_coeff, const double *pvar, const double *coefa,
                 const double *coefb, double *val_ip)
{
  /* cell -> boundary faces, ascending (cs_mesh_adjacencies.cpp:705-760) */
  int *idx = calloc((size_t)n_cells + 2, sizeof(int));
  int *ids = malloc(sizeof(int)*(size_t)(n_b_faces > 0 ? n_b_faces : 1));
  for (int f = 0; f < n_b_faces; f++) idx[b_face_cells[f] + 2]++;
  for (int c = 0; c < n_cells; c++) idx[c+2] += idx[c+1];
  for (int f = 0; f < n_b_faces; f++) ids[idx[b_face_cells[f] + 1]++] = f;

  const double eps_r_2 = 1e-3 * 1e-3;
  for (int f_id = 0; f_id < n_b_faces; f_id++) {
    const int c_id = b_face_cells[f_id];
    const double *ip = diipb + 3*(size_t)f_id;
    if (ip[0]*ip[0] + ip[1]*ip[1] + ip[2]*ip[2] < (b_dist[f_id]*b_dist[f_id]) * eps_r_2) {
      val_ip[f_id] = pvar[c_id];
      continue;
    }
    float cocg[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    double rhs[3] = {0., 0., 0.};
    const double var_i = pvar[c_id];
    double var_min = var_i, var_max = var_i;
    for (int k = row_index[c_id]; k < row_index[c_id+1]; k++) {
      const int c1 = col_id[k];
      var_min = fmin(var_min, pvar[c1]);
      var_max = fmax(var_max, pvar[c1]);
      double dc[3];
      for (int ll = 0; ll < 3; ll++)
        dc[ll] = cell_cen[3*(size_t)c1+ll] - cell_cen[3*(size_t)c_id+ll];
      const double ddc = 1. / (dc[0]*dc[0] + dc[1]*dc[1] + dc[2]*dc[2]);
      cocg[0] += dc[0]*dc[0]*ddc;
      cocg[1] += dc[1]*dc[1]*ddc;
      cocg[2] += dc[2]*dc[2]*ddc;
      cocg[3] += dc[0]*dc[1]*ddc;
      cocg[4] += dc[1]*dc[2]*ddc;
      cocg[5] += dc[0]*dc[2]*ddc;
      double pfac = pvar[c1] - var_i;
      pfac *= ddc;
      for (int ll = 0; ll < 3; ll++) rhs[ll] += dc[ll] * pfac;
    }
    for (int k = idx[c_id]; k < idx[c_id+1]; k++) {
      const int f = ids[k];
      const double a = coefa[f], b = coefb[f];
      const double ddif = 1. / b_dist[f];
      const double umcbdd = (1. - b) * ddif;
      double dif[3];
      for (int ll = 0; ll < 3; ll++)
        dif[ll] = b_face_u_normal[3*(size_t)f+ll] + umcbdd * diipb[3*(size_t)f+ll];
      cocg[0] += dif[0]*dif[0];
      cocg[1] += dif[1]*dif[1];
      cocg[2] += dif[2]*dif[2];
      cocg[3] += dif[0]*dif[1];
      cocg[4] += dif[1]*dif[2];
      cocg[5] += dif[0]*dif[2];
      double pfac = (a + (b-1.)*var_i);
      pfac *= ddif;
      for (int ll = 0; ll < 3; ll++) rhs[ll] += dif[ll] * pfac;
      const double var_f = a + b*var_i;        /* unreconstructed value, for the limiter */
      var_min = fmin(var_min, var_f);
      var_max = fmax(var_max, var_f);
    }
    /* cofactors: float operands, float arithmetic (the products are cs_cocg_t * cs_cocg_t) */
    const double a00 = cocg[1]*cocg[2] - cocg[4]*cocg[4];
    const double a01 = cocg[4]*cocg[5] - cocg[3]*cocg[2];
    const double a02 = cocg[3]*cocg[4] - cocg[1]*cocg[5];
    const double a11 = cocg[0]*cocg[2] - cocg[5]*cocg[5];
    const double a12 = cocg[3]*cocg[5] - cocg[0]*cocg[4];
    const double a22 = cocg[0]*cocg[1] - cocg[3]*cocg[3];
    const double det_inv = 1. / (cocg[0]*a00 + cocg[3]*a01 + cocg[5]*a02);
    double g[3];
    g[0] = (a00 * rhs[0] + a01 * rhs[1] + a02 * rhs[2]) * det_inv;
    g[1] = (a01 * rhs[0] + a11 * rhs[1] + a12 * rhs[2]) * det_inv;
    g[2] = (a02 * rhs[0] + a12 * rhs[1] + a22 * rhs[2]) * det_inv;
    const double grad_dot_diipb = g[0]*ip[0] + g[1]*ip[1] + g[2]*ip[2];
    val_ip[f_id] = _clip_ip(var_i + grad_dot_diipb, var_i, var_min, var_max, clip_coeff);
  }
  free(idx); free(ids);
}

/* (row_index, col_id): the cell -> cells adjacency in the order of cs_mesh_adjacencies
   (= the MSR structure); cell_b_faces: boundary faces by cell, ascending.  b_face_u_normal: 3
   per boundary face; grad: 3 per cell */
void
ora_gradient_lsq_b(int n_cells, const int *row_index, const int *col_id, int n_i_faces,
                   const int *i_face_cells, int n_b_faces, const int *b_face_cells,
                   const double *cell_cen, const double *b_face_u_normal, const double *b_dist,
                   const double *diipb, const double *val_ip, int inc, const double *pvar,
                   const double *coefa, const double *coefb, double *grad)
{
  float *cocg = calloc(6*(size_t)(n_cells > 0 ? n_cells : 1), sizeof(float));
  double *rhsv = calloc(3*(size_t)(n_cells > 0 ? n_cells : 1), sizeof(double));
  for (int c = 0; c < n_cells; c++) {
    float *g = cocg + 6*(size_t)c;
    for (int k = row_index[c]; k < row_index[c+1]; k++) {
      const int c1 = col_id[k];
      const double dc[3] = {cell_cen[3*c1] - cell_cen[3*c], cell_cen[3*c1+1] - cell_cen[3*c+1],
                            cell_cen[3*c1+2] - cell_cen[3*c+2]};
      const double ddc = 1. / (dc[0]*dc[0] + dc[1]*dc[1] + dc[2]*dc[2]);
      g[0] += dc[0] * dc[0] * ddc;
      g[1] += dc[1] * dc[1] * ddc;
      g[2] += dc[2] * dc[2] * ddc;
      g[3] += dc[0] * dc[1] * ddc;
      g[4] += dc[1] * dc[2] * ddc;
      g[5] += dc[0] * dc[2] * ddc;
    }
  }
  for (int f = 0; f < n_b_faces; f++) {       /* ascending faces = every cell's c2b order */
    float *g = cocg + 6*(size_t)b_face_cells[f];
    double d[3];
    _b_direction(f, b_face_u_normal, b_dist, diipb, d);
    g[0] += d[0] * d[0];
    g[1] += d[1] * d[1];
    g[2] += d[2] * d[2];
    g[3] += d[0] * d[1];
    g[4] += d[1] * d[2];
    g[5] += d[0] * d[2];
  }
  for (int c = 0; c < n_cells; c++) _gauss_inverse_6f(cocg + 6*(size_t)c);

  for (int f = 0; f < n_i_faces; f++) {
    const int ii = i_face_cells[2*f], jj = i_face_cells[2*f+1];
    const double dc[3] = {cell_cen[3*jj] - cell_cen[3*ii], cell_cen[3*jj+1] - cell_cen[3*ii+1],
                          cell_cen[3*jj+2] - cell_cen[3*ii+2]};
    const double pfac = (pvar[jj] - pvar[ii]) / (dc[0]*dc[0] + dc[1]*dc[1] + dc[2]*dc[2]);
    for (int ll = 0; ll < 3; ll++) {
      const double fctb = dc[ll] * pfac;
      if (ii < n_cells) rhsv[3*ii+ll] += fctb;
      if (jj < n_cells) rhsv[3*jj+ll] += fctb;
    }
  }
  for (int f = 0; f < n_b_faces; f++) {
    const int c = b_face_cells[f];
    const double val_f = inc*coefa[f] + coefb[f]*(val_ip != NULL ? val_ip[f] : pvar[c]);
    const double ddif = 1. / b_dist[f];
    double dif[3];
    _b_direction(f, b_face_u_normal, b_dist, diipb, dif);
    const double pfac = (val_f - pvar[c]) * ddif;
    for (int ll = 0; ll < 3; ll++) rhsv[3*c+ll] += dif[ll] * pfac;
  }
  for (int c = 0; c < n_cells; c++) {
    const float *g = cocg + 6*(size_t)c;
    const double *r = rhsv + 3*(size_t)c;
    grad[3*c]   = g[0]*r[0] + g[3]*r[1] + g[5]*r[2];
    grad[3*c+1] = g[3]*r[0] + g[1]*r[1] + g[4]*r[2];
    grad[3*c+2] = g[5]*r[0] + g[4]*r[1] + g[2]*r[2];
  }
  free(cocg); free(rhsv);
}

void
ora_gradient_lsq(int n_cells, const int *row_index, const int *col_id, int n_i_faces,
                 const int *i_face_cells, int n_b_faces, const int *b_face_cells,
                 const double *cell_cen, const double *b_face_u_normal, const double *b_dist,
                 int inc, const double *pvar, const double *coefa, const double *coefb,
                 double *grad)
{
  ora_gradient_lsq_b(n_cells, row_index, col_id, n_i_faces, i_face_cells, n_b_faces,
                     b_face_cells, cell_cen, b_face_u_normal, b_dist, NULL, NULL, inc, pvar,
                     coefa, coefb, grad);
}

/* cs_diffusion_potential with flux reconstruction (:8895-8982); i_geo: per interior face II'
   [3], JJ' [3], distance of the cell centres [1]; irecmf as
   cs_glob_velocity_pressure_param->irecmf */
void
ora_diffusion_potential_rc_b(int n_cells, int n_i_faces, const int *i_face_cells,
                             int n_b_faces, const int *b_face_cells, const double *i_geo,
                             const double *i_face_surf, const double *val_ip, int irecmf,
                             int init, int inc, const double *pvar, const double *grad,
                             const double *cofaf, const double *cofbf, const double *i_visc,
                             const double *b_visc, const double *visel, double *diverg)
{
  if (init >= 1)
    for (int i = 0; i < n_cells; i++) diverg[i] = 0.;
  for (int f = 0; f < n_i_faces; f++) {
    const int c0 = i_face_cells[2*f], c1 = i_face_cells[2*f+1];
    const double *ip = i_geo + 7*(size_t)f, *jp = ip + 3;
    const double *g0 = grad + 3*(size_t)c0, *g1 = grad + 3*(size_t)c1;
    double flux = i_visc[f]*(pvar[c0] - pvar[c1]);
    if (irecmf == 0) {
      const double dijx = ip[0] - jp[0], dijy = ip[1] - jp[1], dijz = ip[2] - jp[2];
      const double dpxf = 0.5*(visel[c0]*g0[0] + visel[c1]*g1[0]);
      const double dpyf = 0.5*(visel[c0]*g0[1] + visel[c1]*g1[1]);
      const double dpzf = 0.5*(visel[c0]*g0[2] + visel[c1]*g1[2]);
      flux += (dpxf*dijx + dpyf*dijy + dpzf*dijz) * i_face_surf[f]/ip[6];
    }
    else
      flux += i_visc[f] * (  (g0[0]*ip[0] + g0[1]*ip[1] + g0[2]*ip[2])
                           - (g1[0]*jp[0] + g1[1]*jp[1] + g1[2]*jp[2]));
    if (c0 < n_cells) diverg[c0] += flux;
    if (c1 < n_cells) diverg[c1] += -flux;
  }
  for (int f = 0; f < n_b_faces; f++) {
    const int c0 = b_face_cells[f];
    const double flux_d = inc*cofaf[f] + cofbf[f]*(val_ip != NULL ? val_ip[f] : pvar[c0]);
    diverg[c0] += b_visc[f]*flux_d;
  }
}

void
ora_diffusion_potential_rc(int n_cells, int n_i_faces, const int *i_face_cells, int n_b_faces,
                           const int *b_face_cells, const double *i_geo,
                           const double *i_face_surf, int irecmf, int init, int inc,
                           const double *pvar, const double *grad, const double *cofaf,
                           const double *cofbf, const double *i_visc, const double *b_visc,
                           const double *visel, double *diverg)
{
  ora_diffusion_potential_rc_b(n_cells, n_i_faces, i_face_cells, n_b_faces, b_face_cells, i_geo,
                               i_face_surf, NULL, irecmf, init, inc, pvar, grad, cofaf, cofbf,
                               i_visc, b_visc, visel, diverg);
}

/* Green-Gauss gradient with least-squares face values (imrgra = 4, CS_GRADIENT_GREEN_LSQ, the
   default of cs_glob_space_disc; cs_gradient.cpp:7697-7797): r_grad is the least-squares
   gradient above, then _reconstruct_scalar_gradient (:4167-4540, branch without hydrostatic
   pressure :4428-4506; no cell weighting, no disabled cell, no warped-face correction):
   interior faces in face order add (pfac + rfac) S n to both cells with the face weight
   (cs_mesh_quantities_t::weight) and the OF vector (dofij), boundary faces add
   (val_f - var) S n, then the sums are divided by the cell volume and, when corr_grad_lin
   (cs_mesh_quantities_t::corr_grad_lin, 3x3 per cell) is given, multiplied by it -- the
   warped-face correction CS_BAD_CELLS_WARPED_CORRECTION, on by default in the reference.
   val_ip: values at I' of the boundary faces, or NULL (II' = 0). */
void
ora_gradient_green_reconstruct(int n_cells, int n_i_faces, const int *i_face_cells,
                               int n_b_faces, const int *b_face_cells, const double *weight,
                               const double *dofij, const double *i_face_surf,
                               const double *i_face_u_normal, const double *b_face_surf,
                               const double *b_face_u_normal, const double *cell_vol,
                               const double *val_ip, int inc, const double *pvar,
                               const double *coefa, const double *coefb, const double *r_grad,
                               const double *corr_grad_lin, double *grad)
{
  for (int i = 0; i < 3*n_cells; i++) grad[i] = 0.;
  for (int f = 0; f < n_i_faces; f++) {
    const int c1 = i_face_cells[2*f], c2 = i_face_cells[2*f+1];
    const double ktpond = weight[f];
    const double d_var = pvar[c2] - pvar[c1];
    const double pfaci = (1.0-ktpond) * d_var;
    const double pfacj =     -ktpond  * d_var;
    const double *o = dofij + 3*(size_t)f, *g1 = r_grad + 3*(size_t)c1, *g2 = r_grad + 3*(size_t)c2;
    const double rfac = 0.5 * (  o[0]*(g1[0]+g2[0])
                               + o[1]*(g1[1]+g2[1])
                               + o[2]*(g1[2]+g2[2]));
    const double _pfaci = (pfaci + rfac) * i_face_surf[f];
    const double _pfacj = (pfacj + rfac) * i_face_surf[f];
    for (int j = 0; j < 3; j++) {
      const double rhsv1 =   _pfaci * i_face_u_normal[3*(size_t)f+j];
      const double rhsv2 = - _pfacj * i_face_u_normal[3*(size_t)f+j];
      if (c1 < n_cells) grad[3*(size_t)c1+j] += rhsv1;
      if (c2 < n_cells) grad[3*(size_t)c2+j] += rhsv2;
    }
  }
  for (int f = 0; f < n_b_faces; f++) {
    const int c = b_face_cells[f];
    const double val_f = inc*coefa[f] + coefb[f]*(val_ip != NULL ? val_ip[f] : pvar[c]);
    const double pfac = (val_f - pvar[c]) * b_face_surf[f];
    for (int j = 0; j < 3; j++)
      grad[3*(size_t)c+j] += pfac * b_face_u_normal[3*(size_t)f+j];
  }
  for (int c = 0; c < n_cells; c++) {
    const double dvol = 1. / cell_vol[c];
    grad[3*(size_t)c]   *= dvol;
    grad[3*(size_t)c+1] *= dvol;
    grad[3*(size_t)c+2] *= dvol;
    if (corr_grad_lin != NULL) {     /* CS_BAD_CELLS_WARPED_CORRECTION (:4519-4530) */
      double gradpa[3];
      for (int i = 0; i < 3; i++) {
        gradpa[i] = grad[3*(size_t)c+i];
        grad[3*(size_t)c+i] = 0.;
      }
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
          grad[3*(size_t)c+i] += corr_grad_lin[9*(size_t)c+3*i+j] * gradpa[j];
    }
  }
}

/* cs_divergence (src/alge/cs_divergence.cpp:1224-1290): signed sums of the face mass fluxes,
   interior faces in face order then boundary faces */
void
ora_divergence(int n_cells, int n_i_faces, const int *i_face_cells, int n_b_faces,
               const int *b_face_cells, int init, const double *i_massflux,
               const double *b_massflux, double *diverg)
{
  if (init >= 1)
    for (int i = 0; i < n_cells; i++) diverg[i] = 0.;
  for (int f = 0; f < n_i_faces; f++) {
    const int ii = i_face_cells[2*f], jj = i_face_cells[2*f+1];
    if (ii < n_cells) diverg[ii] += i_massflux[f];
    if (jj < n_cells) diverg[jj] += -i_massflux[f];
  }
  for (int f = 0; f < n_b_faces; f++) diverg[b_face_cells[f]] += b_massflux[f];
}

/* cs_face_diffusion_potential (src/alge/cs_convection_diffusion.cpp:8075-8310): the face
   fluxes cs_diffusion_potential sums, kept per face.  grad NULL: eqp->ircflu = 0 (:8236-8246,
   boundary flux from the cell value); else the reconstructed form (:8254-8296, always the
   II' - JJ' formula) with the boundary flux from val_ip (NULL: the cell value). */
void
ora_face_diffusion_potential(int n_cells, int n_i_faces, const int *i_face_cells,
                             int n_b_faces, const int *b_face_cells, const double *i_geo,
                             const double *i_face_surf, const double *val_ip, int init,
                             int inc, const double *pvar, const double *grad,
                             const double *cofaf, const double *cofbf, const double *i_visc,
                             const double *b_visc, const double *visel, double *i_massflux,
                             double *b_massflux)
{
  (void)n_cells;
  if (init >= 1) {
    for (int f = 0; f < n_i_faces; f++) i_massflux[f] = 0.;
    for (int f = 0; f < n_b_faces; f++) b_massflux[f] = 0.;
  }
  for (int f = 0; f < n_i_faces; f++) {
    const int c0 = i_face_cells[2*f], c1 = i_face_cells[2*f+1];
    if (grad == NULL) {
      i_massflux[f] += i_visc[f]*(pvar[c0] - pvar[c1]);
      continue;
    }
    const double *ip = i_geo + 7*(size_t)f, *jp = ip + 3;
    const double *g0 = grad + 3*(size_t)c0, *g1 = grad + 3*(size_t)c1;
    const double dpxf = 0.5*(visel[c0]*g0[0] + visel[c1]*g1[0]);
    const double dpyf = 0.5*(visel[c0]*g0[1] + visel[c1]*g1[1]);
    const double dpzf = 0.5*(visel[c0]*g0[2] + visel[c1]*g1[2]);
    const double dijx = ip[0] - jp[0], dijy = ip[1] - jp[1], dijz = ip[2] - jp[2];
    const double f_mass_flux =   i_visc[f]*(pvar[c0] - pvar[c1])
                               + (dpxf*dijx + dpyf*dijy + dpzf*dijz) * i_face_surf[f]/ip[6];
    i_massflux[f] += f_mass_flux;
  }
  for (int f = 0; f < n_b_faces; f++) {
    const int c0 = b_face_cells[f];
    const double flux_d
      = inc*cofaf[f] + cofbf[f]*((grad != NULL && val_ip != NULL) ? val_ip[f] : pvar[c0]);
    b_massflux[f] += b_visc[f]*flux_d;
  }
}

/*----------------------------------------------------------------------------
 * Hydrostatic pressure treatment (hyd_p_flag = 1: cs_glob_velocity_pressure_param->iphydr = 1,
 * the reference's default, hands iphydp = 1 and the exterior force f_ext [n_cells][3] to
 * cs_diffusion_potential, src/base/cs_pressure_correction.cpp:2113-2127).
 *----------------------------------------------------------------------------*/

/* (xb - xa) . xc, cs_math_3_distance_dot_product (cs_math.h:171-176) */
static inline double
_ddp(const double *xa, const double *xb, const double *xc)
{
  return ((xb[0] - xa[0])*xc[0]+(xb[1] - xa[1])*xc[1]+(xb[2] - xa[2])*xc[2]);
}

/* Values at I' (cs_gradient_boundary_iprime_lsq_s with hyd_p_flag = 1,
   cs_gradient_boundary.cpp:285-640: the interior differences gain (I - F).f_I - (J - F).f_J
   with F the centre of the shared face, the boundary terms (b - 1) (F_b - I).f_I, and f_I is
   added to the local gradient) and the face values of _compute_bc_flux_grad
   (cs_boundary_conditions_set_coeffs.cpp:168-192: val_f = inc a + b val_ip + b (F_b - I').f_I).
   c2f: interior face of every adjacency entry (cs_mesh_adjacencies cell_i_faces).  diipb: 3 per
   boundary face (zeros allowed). */
void
ora_b_face_values_hyd(int n_cells, const int *row_index, const int *col_id, const int *c2f,
                      int n_b_faces, const int *b_face_cells, const double *cell_cen,
                      const double *b_face_u_normal, const double *b_dist, const double *diipb,
                      const double *i_face_cog, const double *b_face_cog, const double *f_ext,
                      double clip_coeff, int inc, const double *pvar, const double *coefa,
                      const double *coefb, double *val_ip, double *val_f)
{
  int *idx = calloc((size_t)n_cells + 2, sizeof(int));
  int *ids = malloc(sizeof(int)*(size_t)(n_b_faces > 0 ? n_b_faces : 1));
  for (int f = 0; f < n_b_faces; f++) idx[b_face_cells[f] + 2]++;
  for (int c = 0; c < n_cells; c++) idx[c+2] += idx[c+1];
  for (int f = 0; f < n_b_faces; f++) ids[idx[b_face_cells[f] + 1]++] = f;

  const double eps_r_2 = 1e-3 * 1e-3;
  for (int f_id = 0; f_id < n_b_faces; f_id++) {
    const int c_id = b_face_cells[f_id];
    const double *ip = diipb + 3*(size_t)f_id;
    const double *cen = cell_cen + 3*(size_t)c_id, *fx = f_ext + 3*(size_t)c_id;
    const double var_i = pvar[c_id];
    double vip = var_i;
    if (!(ip[0]*ip[0] + ip[1]*ip[1] + ip[2]*ip[2] < (b_dist[f_id]*b_dist[f_id]) * eps_r_2)) {
      float cocg[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      double rhs[3] = {0., 0., 0.};
      double var_min = var_i, var_max = var_i;
      for (int k = row_index[c_id]; k < row_index[c_id+1]; k++) {
        const int c1 = col_id[k];
        var_min = fmin(var_min, pvar[c1]);
        var_max = fmax(var_max, pvar[c1]);
        const double *cen1 = cell_cen + 3*(size_t)c1;
        double dc[3];
        for (int ll = 0; ll < 3; ll++) dc[ll] = cen1[ll] - cen[ll];
        const double ddc = 1. / (dc[0]*dc[0] + dc[1]*dc[1] + dc[2]*dc[2]);
        cocg[0] += dc[0]*dc[0]*ddc;
        cocg[1] += dc[1]*dc[1]*ddc;
        cocg[2] += dc[2]*dc[2]*ddc;
        cocg[3] += dc[0]*dc[1]*ddc;
        cocg[4] += dc[1]*dc[2]*ddc;
        cocg[5] += dc[0]*dc[2]*ddc;
        double pfac = pvar[c1] - var_i;
        const double *cog = i_face_cog + 3*(size_t)c2f[k];
        const double dot_i = _ddp(cog, cen, fx);
        const double dot_j = _ddp(cog, cen1, f_ext + 3*(size_t)c1);
        pfac += dot_i - dot_j;
        pfac *= ddc;
        for (int ll = 0; ll < 3; ll++) rhs[ll] += dc[ll] * pfac;
      }
      for (int k = idx[c_id]; k < idx[c_id+1]; k++) {
        const int f = ids[k];
        const double a = coefa[f], b = coefb[f];
        const double ddif = 1. / b_dist[f];
        const double umcbdd = (1. - b) * ddif;
        double dif[3];
        for (int ll = 0; ll < 3; ll++)
          dif[ll] = b_face_u_normal[3*(size_t)f+ll] + umcbdd * diipb[3*(size_t)f+ll];
        cocg[0] += dif[0]*dif[0];
        cocg[1] += dif[1]*dif[1];
        cocg[2] += dif[2]*dif[2];
        cocg[3] += dif[0]*dif[1];
        cocg[4] += dif[1]*dif[2];
        cocg[5] += dif[0]*dif[2];
        double pfac = (a + (b-1.)*var_i);
        const double dot = _ddp(cen, b_face_cog + 3*(size_t)f, fx);
        pfac += (b-1.) * dot;
        pfac *= ddif;
        for (int ll = 0; ll < 3; ll++) rhs[ll] += dif[ll] * pfac;
        const double var_f = a + b*var_i;
        var_min = fmin(var_min, var_f);
        var_max = fmax(var_max, var_f);
      }
      const double a00 = cocg[1]*cocg[2] - cocg[4]*cocg[4];
      const double a01 = cocg[4]*cocg[5] - cocg[3]*cocg[2];
      const double a02 = cocg[3]*cocg[4] - cocg[1]*cocg[5];
      const double a11 = cocg[0]*cocg[2] - cocg[5]*cocg[5];
      const double a12 = cocg[3]*cocg[5] - cocg[0]*cocg[4];
      const double a22 = cocg[0]*cocg[1] - cocg[3]*cocg[3];
      const double det_inv = 1. / (cocg[0]*a00 + cocg[3]*a01 + cocg[5]*a02);
      double g[3];
      g[0] = (a00 * rhs[0] + a01 * rhs[1] + a02 * rhs[2]) * det_inv;
      g[1] = (a01 * rhs[0] + a11 * rhs[1] + a12 * rhs[2]) * det_inv;
      g[2] = (a02 * rhs[0] + a12 * rhs[1] + a22 * rhs[2]) * det_inv;
      g[0] += fx[0]; g[1] += fx[1]; g[2] += fx[2];
      vip = _clip_ip(var_i + (g[0]*ip[0] + g[1]*ip[1] + g[2]*ip[2]), var_i, var_min, var_max,
                     clip_coeff);
    }
    val_ip[f_id] = vip;
    double vf = inc * coefa[f_id] + coefb[f_id] * vip;
    const double vec_iprime[3] = {cen[0] + ip[0], cen[1] + ip[1], cen[2] + ip[2]};
    vf += coefb[f_id] * _ddp(vec_iprime, b_face_cog + 3*(size_t)f_id, fx);
    val_f[f_id] = vf;
  }
  free(idx); free(ids);
}

/* _lsq_scalar_gradient_hyd_p (cs_gradient.cpp:3230-3560, scatter form = the one-thread path);
   val_f from ora_b_face_values_hyd; diipb may be NULL */
void
ora_gradient_lsq_hyd(int n_cells, const int *row_index, const int *col_id, int n_i_faces,
                     const int *i_face_cells, int n_b_faces, const int *b_face_cells,
                     const double *cell_cen, const double *b_face_u_normal,
                     const double *b_dist, const double *diipb, const double *i_face_cog,
                     const double *b_face_cog, const double *f_ext, const double *val_f,
                     const double *pvar, double *grad)
{
  float *cocg = calloc(6*(size_t)(n_cells > 0 ? n_cells : 1), sizeof(float));
  double *rhsv = calloc(3*(size_t)(n_cells > 0 ? n_cells : 1), sizeof(double));
  for (int c = 0; c < n_cells; c++) {
    float *g = cocg + 6*(size_t)c;
    for (int k = row_index[c]; k < row_index[c+1]; k++) {
      const int c1 = col_id[k];
      const double dc[3] = {cell_cen[3*c1] - cell_cen[3*c], cell_cen[3*c1+1] - cell_cen[3*c+1],
                            cell_cen[3*c1+2] - cell_cen[3*c+2]};
      const double ddc = 1. / (dc[0]*dc[0] + dc[1]*dc[1] + dc[2]*dc[2]);
      g[0] += dc[0] * dc[0] * ddc;
      g[1] += dc[1] * dc[1] * ddc;
      g[2] += dc[2] * dc[2] * ddc;
      g[3] += dc[0] * dc[1] * ddc;
      g[4] += dc[1] * dc[2] * ddc;
      g[5] += dc[0] * dc[2] * ddc;
    }
  }
  for (int f = 0; f < n_b_faces; f++) {
    float *g = cocg + 6*(size_t)b_face_cells[f];
    double d[3];
    _b_direction(f, b_face_u_normal, b_dist, diipb, d);
    g[0] += d[0] * d[0];
    g[1] += d[1] * d[1];
    g[2] += d[2] * d[2];
    g[3] += d[0] * d[1];
    g[4] += d[1] * d[2];
    g[5] += d[0] * d[2];
  }
  for (int c = 0; c < n_cells; c++) _gauss_inverse_6f(cocg + 6*(size_t)c);

  for (int f = 0; f < n_i_faces; f++) {
    const int ii = i_face_cells[2*f], jj = i_face_cells[2*f+1];
    const double *ci = cell_cen + 3*(size_t)ii, *cj = cell_cen + 3*(size_t)jj;
    const double dc[3] = {cj[0] - ci[0], cj[1] - ci[1], cj[2] - ci[2]};
    const double dvarij = pvar[jj] - pvar[ii];
    const double *cog = i_face_cog + 3*(size_t)f;
    double pfac =   dvarij
                  + _ddp(cog, ci, f_ext + 3*(size_t)ii)
                  - _ddp(cog, cj, f_ext + 3*(size_t)jj);
    pfac /= (dc[0]*dc[0] + dc[1]*dc[1] + dc[2]*dc[2]);
    for (int ll = 0; ll < 3; ll++) {
      const double fctb = dc[ll] * pfac;
      if (ii < n_cells) rhsv[3*ii+ll] += fctb;
      if (jj < n_cells) rhsv[3*jj+ll] += fctb;
    }
  }
  for (int f = 0; f < n_b_faces; f++) {
    const int c = b_face_cells[f];
    const double poro = 0.;
    const double if_dot_fext = _ddp(cell_cen + 3*(size_t)c, b_face_cog + 3*(size_t)f,
                                    f_ext + 3*(size_t)c);
    const double ddif = 1. / b_dist[f];
    double dif[3];
    _b_direction(f, b_face_u_normal, b_dist, diipb, dif);
    const double pfac = ddif * (val_f[f] - pvar[c] - poro - if_dot_fext);
    for (int ll = 0; ll < 3; ll++) rhsv[3*c+ll] += dif[ll] * pfac;
  }
  for (int c = 0; c < n_cells; c++) {
    const float *g = cocg + 6*(size_t)c;
    const double *r = rhsv + 3*(size_t)c, *fx = f_ext + 3*(size_t)c;
    grad[3*c]   = g[0]*r[0] + g[3]*r[1] + g[5]*r[2] + fx[0];
    grad[3*c+1] = g[3]*r[0] + g[1]*r[1] + g[4]*r[2] + fx[1];
    grad[3*c+2] = g[5]*r[0] + g[4]*r[1] + g[2]*r[2] + fx[2];
  }
  free(cocg); free(rhsv);
}

/* _reconstruct_scalar_gradient, branch with hydrostatic pressure (cs_gradient.cpp:4314-4424) */
void
ora_gradient_green_reconstruct_hyd(int n_cells, int n_i_faces, const int *i_face_cells,
                                   int n_b_faces, const int *b_face_cells,
                                   const double *weight, const double *dofij,
                                   const double *i_face_surf, const double *i_face_u_normal,
                                   const double *b_face_surf, const double *b_face_u_normal,
                                   const double *cell_vol, const double *cell_cen,
                                   const double *i_face_cog, const double *f_ext,
                                   const double *val_f, const double *pvar,
                                   const double *r_grad, const double *corr_grad_lin,
                                   double *grad)
{
  for (int i = 0; i < 3*n_cells; i++) grad[i] = 0.;
  for (int f = 0; f < n_i_faces; f++) {
    const int c1 = i_face_cells[2*f], c2 = i_face_cells[2*f+1];
    const double ktpond = weight[f];
    const double poro[2] = {0., 0.};
    const double *f1 = f_ext + 3*(size_t)c1, *f2 = f_ext + 3*(size_t)c2;
    const double *cog = i_face_cog + 3*(size_t)f;
    double fexd[3], d_oi[3], d_oj[3];
    for (int k = 0; k < 3; k++) {
      fexd[k] = 0.5 * (f1[k] + f2[k]);
      d_oi[k] = cog[k] - cell_cen[3*(size_t)c1+k];
      d_oj[k] = cog[k] - cell_cen[3*(size_t)c2+k];
    }
    double pfaci
      =  ktpond
           * (  d_oi[0]*f1[0]
              + d_oi[1]*f1[1]
              + d_oi[2]*f1[2]
              + poro[0])
      +  (1.0 - ktpond)
           * (  d_oj[0]*f2[0]
              + d_oj[1]*f2[1]
              + d_oj[2]*f2[2]
              + poro[1]);
    double pfacj = pfaci;
    const double d_var = pvar[c2] - pvar[c1];
    pfaci += (1.0-ktpond) * d_var;
    pfacj -=      ktpond  * d_var;
    const double *o = dofij + 3*(size_t)f, *g1 = r_grad + 3*(size_t)c1, *g2 = r_grad + 3*(size_t)c2;
    const double rfac =
             - weight[f]      * (d_oi[0]*fexd[0] + d_oi[1]*fexd[1] + d_oi[2]*fexd[2])
         -  (1.0 - weight[f]) * (d_oj[0]*fexd[0] + d_oj[1]*fexd[1] + d_oj[2]*fexd[2])
         + (  o[0] * (g1[0]+g2[0])
            + o[1] * (g1[1]+g2[1])
            + o[2] * (g1[2]+g2[2])) * 0.5;
    const double _pfaci = (pfaci + rfac) * i_face_surf[f];
    const double _pfacj = (pfacj + rfac) * i_face_surf[f];
    for (int j = 0; j < 3; j++) {
      const double rhsv1 =   _pfaci * i_face_u_normal[3*(size_t)f+j];
      const double rhsv2 = - _pfacj * i_face_u_normal[3*(size_t)f+j];
      if (c1 < n_cells) grad[3*(size_t)c1+j] += rhsv1;
      if (c2 < n_cells) grad[3*(size_t)c2+j] += rhsv2;
    }
  }
  for (int f = 0; f < n_b_faces; f++) {
    const int c = b_face_cells[f];
    const double pfac = (val_f[f] - pvar[c]) * b_face_surf[f];
    for (int j = 0; j < 3; j++)
      grad[3*(size_t)c+j] += pfac * b_face_u_normal[3*(size_t)f+j];
  }
  for (int c = 0; c < n_cells; c++) {
    const double dvol = 1. / cell_vol[c];
    grad[3*(size_t)c]   *= dvol;
    grad[3*(size_t)c+1] *= dvol;
    grad[3*(size_t)c+2] *= dvol;
    if (corr_grad_lin != NULL) {
      double gradpa[3];
      for (int i = 0; i < 3; i++) {
        gradpa[i] = grad[3*(size_t)c+i];
        grad[3*(size_t)c+i] = 0.;
      }
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
          grad[3*(size_t)c+i] += corr_grad_lin[9*(size_t)c+3*i+j] * gradpa[j];
    }
  }
}

/* Iterative Green-Gauss gradient (imrgra = 0, CS_GRADIENT_GREEN_ITER; no hydrostatic term, no
   weighting): _initialize_scalar_gradient (cs_gradient.cpp:1329-1590, the non-reconstructed
   Green-Gauss gradient), then _iterative_scalar_gradient (:1891-2300): sweeps
   rhs = -grad V + sum over faces of the reconstructed face values, grad += cocg_it (rhs / V),
   until ||rhs / V|| < epsrgp ||grad_0|| or nswrgp - 1 sweeps; cocg_it from
   _compute_cell_cocg_it (:1776-1860) with the Cramer inverse (cs_math.h:1587-1610).
   diipb may be NULL (zeros).  Returns the number of sweeps done (the reference's n_sweeps). */
int
ora_gradient_green_iter(int n_cells, int n_i_faces, const int *i_face_cells, int n_b_faces,
                        const int *b_face_cells, const double *weight, const double *dofij,
                        const double *i_face_surf, const double *i_face_u_normal,
                        const double *b_face_surf, const double *b_face_u_normal,
                        const double *cell_vol, const double *diipb, int nswrgp, double epsrgp,
                        int inc_i, const double *pvar, const double *coefa,
                        const double *coefb, double *grad)
{
  const double inc = inc_i;
  const size_t n3 = 3*(size_t)(n_cells > 0 ? n_cells : 1);
  for (int i = 0; i < 3*n_cells; i++) grad[i] = 0.0;
  for (int f = 0; f < n_i_faces; f++) {
    const int ii = i_face_cells[2*f], jj = i_face_cells[2*f+1];
    const double ktpond = weight[f];
    const double pfaci = (1.0-ktpond) * (pvar[jj]-pvar[ii]) * i_face_surf[f];
    const double pfacj =     -ktpond  * (pvar[jj]-pvar[ii]) * i_face_surf[f];
    const double *n = i_face_u_normal + 3*(size_t)f;
    for (int j = 0; j < 3; j++) {
      const double fctb_ii =  pfaci * n[j];
      const double fctb_jj = -pfacj * n[j];
      if (ii < n_cells) grad[3*(size_t)ii+j] += fctb_ii;
      if (jj < n_cells) grad[3*(size_t)jj+j] += fctb_jj;
    }
  }
  for (int f = 0; f < n_b_faces; f++) {
    const int ii = b_face_cells[f];
    double pfac = inc*coefa[f] + (coefb[f]-1.0)*pvar[ii];
    pfac *= b_face_surf[f];
    for (int j = 0; j < 3; j++) grad[3*(size_t)ii+j] += pfac * b_face_u_normal[3*(size_t)f+j];
  }
  for (int c = 0; c < n_cells; c++) {
    const double dvol = 1. / cell_vol[c];
    for (int j = 0; j < 3; j++) grad[3*(size_t)c+j] *= dvol;
  }
  if (nswrgp < 1) return 0;

  /* cocg_it */
  double *cocg = malloc(sizeof(double)*3*n3);
  for (int c = 0; c < n_cells; c++) {
    double *a = cocg + 9*(size_t)c;
    a[0] = 1.0; a[1] = 0.0; a[2] = 0.0; a[3] = 0.0; a[4] = 1.0; a[5] = 0.0;
    a[6] = 0.0; a[7] = 0.0; a[8] = 1.0;
  }
  for (int f = 0; f < n_i_faces; f++) {
    const int c1 = i_face_cells[2*f], c2 = i_face_cells[2*f+1];
    const double dvol1 = (cell_vol[c1] > 0) ? 1. / cell_vol[c1] : 0.;
    const double dvol2 = (cell_vol[c2] > 0) ? 1. / cell_vol[c2] : 0.;
    const double pfac0 = -0.5 * i_face_surf[f];
    for (int i = 0; i < 3; i++) {
      const double pfac = pfac0 * dofij[3*(size_t)f+i];
      for (int j = 0; j < 3; j++) {
        const double vecfac = pfac*i_face_u_normal[3*(size_t)f+j];
        if (c1 < n_cells) cocg[9*(size_t)c1+3*i+j] +=  vecfac * dvol1;
        if (c2 < n_cells) cocg[9*(size_t)c2+3*i+j] += -vecfac * dvol2;
      }
    }
  }
  for (int c = 0; c < n_cells; c++) {
    double (*a)[3] = (double (*)[3])(cocg + 9*(size_t)c);
    const double a00 = a[1][1]*a[2][2] - a[2][1]*a[1][2];
    const double a01 = a[2][1]*a[0][2] - a[0][1]*a[2][2];
    const double a02 = a[0][1]*a[1][2] - a[1][1]*a[0][2];
    const double a10 = a[2][0]*a[1][2] - a[1][0]*a[2][2];
    const double a11 = a[0][0]*a[2][2] - a[2][0]*a[0][2];
    const double a12 = a[1][0]*a[0][2] - a[0][0]*a[1][2];
    const double a20 = a[1][0]*a[2][1] - a[2][0]*a[1][1];
    const double a21 = a[2][0]*a[0][1] - a[0][0]*a[2][1];
    const double a22 = a[0][0]*a[1][1] - a[1][0]*a[0][1];
    const double det_inv = 1. / (a[0][0]*a00 + a[1][0]*a01 + a[2][0]*a02);
    a[0][0] = a00 * det_inv; a[0][1] = a01 * det_inv; a[0][2] = a02 * det_inv;
    a[1][0] = a10 * det_inv; a[1][1] = a11 * det_inv; a[1][2] = a12 * det_inv;
    a[2][0] = a20 * det_inv; a[2][1] = a21 * det_inv; a[2][2] = a22 * det_inv;
  }

  double s = 0;
  for (int i = 0; i < 3*n_cells; i++) s += grad[i]*grad[i];
  const double rnorm = sqrt(s);
  if (rnorm <= 1.e-12) { free(cocg); return 0; }

  double *rhs = malloc(sizeof(double)*n3);
  int n_sweeps;
  for (n_sweeps = 1; n_sweeps < nswrgp; n_sweeps++) {
    for (int c = 0; c < n_cells; c++)
      for (int j = 0; j < 3; j++) rhs[3*(size_t)c+j] = -grad[3*(size_t)c+j] * cell_vol[c];
    for (int f = 0; f < n_i_faces; f++) {
      const int c1 = i_face_cells[2*f], c2 = i_face_cells[2*f+1];
      const double *o = dofij + 3*(size_t)f, *g1 = grad + 3*(size_t)c1, *g2 = grad + 3*(size_t)c2;
      double pfaci = 0.5 * (  o[0] * (g1[0]+g2[0])
                            + o[1] * (g1[1]+g2[1])
                            + o[2] * (g1[2]+g2[2]));
      double pfacj = pfaci;
      const double ktpond = weight[f];
      pfaci += (1.0-ktpond) * (pvar[c2] - pvar[c1]);
      pfacj -=      ktpond  * (pvar[c2] - pvar[c1]);
      pfaci *= i_face_surf[f];
      pfacj *= i_face_surf[f];
      for (int j = 0; j < 3; j++) {
        const double fctb_ii =  pfaci * i_face_u_normal[3*(size_t)f+j];
        const double fctb_jj = -pfacj * i_face_u_normal[3*(size_t)f+j];
        if (c1 < n_cells) rhs[3*(size_t)c1+j] += fctb_ii;
        if (c2 < n_cells) rhs[3*(size_t)c2+j] += fctb_jj;
      }
    }
    for (int f = 0; f < n_b_faces; f++) {
      const int c = b_face_cells[f];
      const double *g = grad + 3*(size_t)c;
      const double ip[3] = {diipb ? diipb[3*(size_t)f] : 0., diipb ? diipb[3*(size_t)f+1] : 0.,
                            diipb ? diipb[3*(size_t)f+2] : 0.};
      double pfac =   coefa[f] * inc
                    + coefb[f] * (ip[0] * g[0] + ip[1] * g[1] + ip[2] * g[2]);
      pfac += (coefb[f] -1.0) * pvar[c];
      pfac *= b_face_surf[f];
      for (int ll = 0; ll < 3; ll++) rhs[3*(size_t)c+ll] += pfac * b_face_u_normal[3*(size_t)f+ll];
    }
    double l2 = 0.;
    for (int c = 0; c < n_cells; c++) {
      const double dvol = 1. / cell_vol[c];
      double *r = rhs + 3*(size_t)c, *g = grad + 3*(size_t)c;
      const double *a = cocg + 9*(size_t)c;
      r[0] *= dvol; r[1] *= dvol; r[2] *= dvol;
      l2 += r[0]*r[0] + r[1]*r[1] + r[2]*r[2];
      g[0] += r[0] * a[0] + r[1] * a[3] + r[2] * a[6];
      g[1] += r[0] * a[1] + r[1] * a[4] + r[2] * a[7];
      g[2] += r[0] * a[2] + r[1] * a[5] + r[2] * a[8];
    }
    if (sqrt(l2) < epsrgp*rnorm) break;
  }
  free(cocg); free(rhs);
  return n_sweeps;
}

/* _scalar_gradient_clipping (cs_gradient.cpp:808-1215), standard neighbourhood: the limiter of
   the cell gradient, after any gradient type.  clip_mode 0 (CS_GRADIENT_LIMIT_CELL): the
   variation |IJ.grad_I| is kept under climgp times the largest |var_I - var_J| of the cell;
   1 (_FACE): the same with the mean gradient of the two cells, then the smallest factor of
   the cell and its neighbours (_gradient_update_face_clip_factor :691-758); 2 (_CELL_RC): the
   values reconstructed at I' / J' of the cell's interior faces stay within climgp times the
   range of the neighbouring values and boundary face values.  c2f / c2f_sgn: face and
   orientation of every adjacency entry; i_geo: II' [3], JJ' [3], distance per interior face. */
void
ora_gradient_clip(int n_cells, const int *row_index, const int *col_id, const int *c2f,
                  const short *c2f_sgn, int n_b_faces, const int *b_face_cells,
                  const double *cell_cen, const double *i_geo, int clip_mode, double climgp,
                  const double *pvar, const double *val_f, double *grad)
{
  if (clip_mode < 0 || climgp < 0) return;
  double *clip_factor = malloc(sizeof(double)*(size_t)(n_cells > 0 ? n_cells : 1));
  double *factor = malloc(sizeof(double)*(size_t)(n_cells > 0 ? n_cells : 1));
  if (clip_mode == 0 || clip_mode == 1) {
    for (int c1 = 0; c1 < n_cells; c1++) {
      double denum = 0, denom = 0;
      const double *g1 = grad + 3*(size_t)c1;
      for (int k = row_index[c1]; k < row_index[c1+1]; k++) {
        const int c2 = col_id[k];
        const double *g2 = grad + 3*(size_t)c2;
        double dist[3];
        for (int j = 0; j < 3; j++) dist[j] = cell_cen[3*(size_t)c1+j] - cell_cen[3*(size_t)c2+j];
        double dist1;
        if (clip_mode == 0)
          dist1 = fabs(dist[0]*g1[0] + dist[1]*g1[1] + dist[2]*g1[2]);
        else {
          const double gradm[3] = {g1[0] + g2[0], g1[1] + g2[1], g1[2] + g2[2]};
          dist1 = 0.5 * fabs(dist[0]*gradm[0] + dist[1]*gradm[1] + dist[2]*gradm[2]);
        }
        const double dvar = fabs(pvar[c1] - pvar[c2]);
        denum = fmax(denum, dist1);
        denom = fmax(denom, dvar);
      }
      double factor1 = 1.;
      if (denum > climgp * denom) factor1 = climgp * denom/denum;
      factor[c1] = factor1;
      clip_factor[c1] = (clip_mode == 0) ? factor1 : 1;
    }
    if (clip_mode == 1)
      for (int c1 = 0; c1 < n_cells; c1++) {
        double l_min_factor = factor[c1];
        for (int k = row_index[c1]; k < row_index[c1+1]; k++)
          l_min_factor = fmin(l_min_factor, factor[col_id[k]]);
        clip_factor[c1] = fmin(clip_factor[c1], l_min_factor);
      }
  }
  else {
    double *v_min = malloc(sizeof(double)*(size_t)(n_cells > 0 ? n_cells : 1));
    double *v_max = malloc(sizeof(double)*(size_t)(n_cells > 0 ? n_cells : 1));
    for (int c1 = 0; c1 < n_cells; c1++) {
      v_min[c1] = pvar[c1]; v_max[c1] = pvar[c1];
      for (int k = row_index[c1]; k < row_index[c1+1]; k++) {
        v_min[c1] = fmin(v_min[c1], pvar[col_id[k]]);
        v_max[c1] = fmax(v_max[c1], pvar[col_id[k]]);
      }
    }
    for (int f = 0; f < n_b_faces; f++) {
      const int c1 = b_face_cells[f];
      v_min[c1] = fmin(v_min[c1], val_f[f]);
      v_max[c1] = fmax(v_max[c1], val_f[f]);
    }
    for (int c1 = 0; c1 < n_cells; c1++) {
      const double vmin = (v_min[c1] - pvar[c1]) * climgp, vmax = (v_max[c1] - pvar[c1]) * climgp;
      const double *g = grad + 3*(size_t)c1;
      double ccf = 1.;
      for (int k = row_index[c1]; k < row_index[c1+1]; k++) {
        const double *dkkpf = i_geo + 7*(size_t)c2f[k] + ((c2f_sgn[k] > 0) ? 0 : 3);
        double sub_factor = 1.;
        const double v_r = dkkpf[0]*g[0] + dkkpf[1]*g[1] + dkkpf[2]*g[2];
        if (v_r < vmin)
          sub_factor = fabs(vmin) / fabs(v_r);
        else if (v_r > vmax)
          sub_factor = fabs(vmax) / fabs(v_r);
        ccf = fmin(ccf, sub_factor);
      }
      clip_factor[c1] = ccf;
    }
    free(v_min); free(v_max);
  }
  for (int c = 0; c < n_cells; c++) {
    const double factor1 = clip_factor[c];
    if (factor1 < 1)
      for (int j = 0; j < 3; j++) grad[3*(size_t)c+j] *= factor1;
  }
  free(clip_factor); free(factor);
}
