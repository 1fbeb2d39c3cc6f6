/*============================================================================
 * oracle/cs_oracle_native.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the reference's hierarchy construction for a matrix
 * "mapped from native" (face-based) arrays, i.e. the entry of a caller that built
 * its cs_matrix_t with cs_matrix_set_coefficients(edges, da, xa)
 * (src/alge/cs_matrix.cpp:5315; cs_grid_create_from_shared takes the use_faces /
 * use_xa branch, src/alge/cs_grid.cpp:7485-7491):
 *
 *   aggregation     _automatic_aggregation_fc            cs_grid.cpp:3997-4263
 *                   (one thread; with several threads the reference switches to
 *                    _automatic_aggregation_dx_msr, cs_grid.cpp:8106-8122)
 *   coarse faces    _coarsen_face_cell                   cs_grid.cpp:637-788
 *   cell quantities _compute_coarse_cell_quantities      cs_grid.cpp:4778-4820
 *   values          _compute_coarse_quantities_native    cs_grid.cpp:5144-5558
 *   matrix          _matrix_from_native                  cs_grid.cpp:7105-7190
 *
 * Symmetric scalar matrices (isym = 1, db_size = eb_size = 1), one rank.
 * Pinned bit for bit against oracle/_ref/libcsref.so (tests/test_oracle.py,
 * tests/golden/make_golden_native.py).  Only tests/, smoke() and bench.py's CPU
 * legs may use this file.
 *============================================================================*/

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "cs_oracle.h"

static void *_ndup(const void *p, size_t s)
{ if (p == NULL) return NULL; void *q = malloc(s > 0 ? s : 1); memcpy(q, p, s); return q; }

/*----------------------------------------------------------------------------
 * _automatic_aggregation_fc (cs_grid.cpp:3997-4263), CS_GRID_COARSENING_SPD_DX,
 * symmetric.  f_c_cell has f_n_cols_ext entries and is initialised here to -2
 * (_coarse_init, cs_grid.cpp:520-522).  Returns the number of coarse cells.
 *----------------------------------------------------------------------------*/

int
ora_aggregation_fc(int level, int f_n_cells, int f_n_cells_ext, int f_n_faces,
                   const int *face_cell, const double *da, const double *xa,
                   int max_aggregation, int *f_c_cell)
{
  const double epsilon = 1.e-6;
  const double penalization_threshold = 1e4;      /* cs_grid.cpp:136 */
  const int ncoarse = 8;
  int npass_max = 10, inc_nei = 0, _max_aggregation = 1, npass = 0;
  int aggr_count = f_n_cells, r_n_faces = f_n_faces, c_n_cells = 0;

  int *c_cardinality = malloc(sizeof(int)*(f_n_cells_ext > 0 ? f_n_cells_ext : 1));
  int *c_aggr_count = malloc(sizeof(int)*(f_n_cells_ext > 0 ? f_n_cells_ext : 1));
  int *f_c_face = malloc(sizeof(int)*(f_n_faces > 0 ? f_n_faces : 1));
  int *merge_flag = malloc(sizeof(int)*(f_n_faces > 0 ? f_n_faces : 1));

  for (int ii = 0; ii < f_n_cells_ext; ii++) {
    f_c_cell[ii] = -2;
    c_cardinality[ii] = -1;
    c_aggr_count[ii] = 1;
  }
  for (int f = 0; f < f_n_faces; f++) {
    merge_flag[f] = f + 1;
    f_c_face[f] = 0;
  }

  /* cardinality = number of neighbours - 1; penalised cells at level 0 (:4088-4108) */
  if (level == 0) {
    double *xv_sum = calloc(f_n_cells_ext > 0 ? f_n_cells_ext : 1, sizeof(double));
    for (int f = 0; f < f_n_faces; f++) {
      const int ii = face_cell[2*f], jj = face_cell[2*f+1];
      c_cardinality[ii] += 1;
      c_cardinality[jj] += 1;
      xv_sum[ii] += fabs(xa[f]);
      xv_sum[jj] += fabs(xa[f]);
    }
    for (int ii = 0; ii < f_n_cells; ii++)
      if (da[ii] > penalization_threshold * xv_sum[ii]) f_c_cell[ii] = -1;
    free(xv_sum);
  }
  else {
    for (int f = 0; f < f_n_faces; f++) {
      c_cardinality[face_cell[2*f]] += 1;
      c_cardinality[face_cell[2*f+1]] += 1;
    }
  }

  do {
    inc_nei++;
    npass++;
    const int n_faces = r_n_faces;
    _max_aggregation++;
    if (_max_aggregation > max_aggregation) _max_aggregation = max_aggregation;

    for (int f = 0; f < n_faces; f++) {
      f_c_face[f] = merge_flag[f];
      merge_flag[f] = 0;
    }
    r_n_faces = 0;
    const double ag_threshold = 1. - epsilon;

    for (int f = 0; f < n_faces; f++) {
      const int c_face = f_c_face[f] - 1;
      const int ii = face_cell[2*c_face], jj = face_cell[2*c_face+1];
      /* never across a rank boundary; at least one of the two not aggregated yet */
      if ((ii < f_n_cells && jj < f_n_cells) && (f_c_cell[ii] < -1 || f_c_cell[jj] < -1)) {
        int count = 0;
        const double f_da0_da1_inv
          = ((c_cardinality[ii]+inc_nei)*(c_cardinality[jj]+inc_nei)) / (da[ii] * da[jj]);
        const double f_xa0_xa1 = xa[c_face] * xa[c_face];
        const double aggr_crit = f_da0_da1_inv * f_xa0_xa1;
        if (aggr_crit > ag_threshold) {
          if (f_c_cell[ii] > -1 && f_c_cell[jj] < -1) {
            if (c_aggr_count[f_c_cell[ii]] < _max_aggregation + 1) {
              f_c_cell[jj] = f_c_cell[ii];
              c_aggr_count[f_c_cell[ii]] += 1;
              count++;
            }
          }
          else if (f_c_cell[ii] < -1 && f_c_cell[jj] > -1) {
            if (c_aggr_count[f_c_cell[jj]] < _max_aggregation + 1) {
              f_c_cell[ii] = f_c_cell[jj];
              c_aggr_count[f_c_cell[jj]] += 1;
              count++;
            }
          }
          else if (f_c_cell[ii] < -1 && f_c_cell[jj] < -1) {
            f_c_cell[ii] = c_n_cells;
            f_c_cell[jj] = c_n_cells;
            c_aggr_count[c_n_cells] += 1;
            c_n_cells++;
            count++;
          }
        }
        if (count == 0 && (f_c_cell[ii] < -1 || f_c_cell[jj] < -1)) {
          merge_flag[r_n_faces] = c_face + 1;
          r_n_faces++;
        }
      }
    }

    aggr_count = 0;
    for (int i = 0; i < f_n_cells; i++)
      if (f_c_cell[i] < -1) aggr_count++;
    if (aggr_count == 0 || (c_n_cells + aggr_count)*ncoarse < f_n_cells || r_n_faces == 0)
      npass_max = npass;
  } while (npass < npass_max);

  for (int i = 0; i < f_n_cells; i++) {
    if (f_c_cell[i] < -1) {
      f_c_cell[i] = c_n_cells;
      c_n_cells++;
    }
  }

  free(c_cardinality); free(c_aggr_count); free(f_c_face); free(merge_flag);
  return c_n_cells;
}

/*----------------------------------------------------------------------------
 * _coarsen_face_cell (cs_grid.cpp:637-788): coarse faces numbered in the order in
 * which the fine faces first reach them; coarse_face[f] = +-(coarse face + 1) or 0.
 * c_face_cell must hold 2*f_n_faces entries.  Returns the number of coarse faces.
 *----------------------------------------------------------------------------*/

int
ora_coarsen_face_cell(int f_n_faces, const int *f_face_cell, const int *coarse_row,
                      int c_n_cells, int *coarse_face, int *c_face_cell)
{
  int *idx = calloc(c_n_cells + 1, sizeof(int));
  for (int f = 0; f < f_n_faces; f++) {
    coarse_face[f] = 0;
    const int ii = coarse_row[f_face_cell[2*f]], jj = coarse_row[f_face_cell[2*f+1]];
    if (ii < jj && ii > -1) idx[ii+1] += 1;
    else if (ii > jj && jj > -1) idx[jj+1] += 1;
  }
  for (int ii = 0; ii < c_n_cells; ii++) idx[ii+1] += idx[ii];
  const int size = idx[c_n_cells];
  int *id = malloc(sizeof(int)*(size > 0 ? size : 1));
  int *face = calloc(size > 0 ? size : 1, sizeof(int));
  int *cnt = calloc(c_n_cells > 0 ? c_n_cells : 1, sizeof(int));
  int c_n_faces = 0;

  for (int f = 0; f < f_n_faces; f++) {
    int ii = coarse_row[f_face_cell[2*f]], jj = coarse_row[f_face_cell[2*f+1]];
    int sign = 1, kk;
    if (ii != jj && ii > -1 && jj > -1) {
      if (ii > jj) { sign = -1; kk = ii; ii = jj; jj = kk; }
      const int start_id = idx[ii], end_id = idx[ii] + cnt[ii];
      for (kk = start_id; kk < end_id; kk++)
        if (id[kk] == jj) break;
      if (kk == end_id) {
        id[kk] = jj;
        face[kk] = c_n_faces + 1;
        c_face_cell[2*c_n_faces] = ii;
        c_face_cell[2*c_n_faces+1] = jj;
        c_n_faces++;
        cnt[ii] += 1;
      }
      coarse_face[f] = sign*face[kk];
    }
  }
  free(idx); free(id); free(face); free(cnt);
  return c_n_faces;
}

/*----------------------------------------------------------------------------
 * MSR matrix of a level from its face arrays: cs_matrix_structure_create(MSR)
 * (columns sorted in every row) + cs_matrix_set_coefficients, and the entry ->
 * face map of _matrix_from_native (cs_grid.cpp:7146-7186).  Unique faces.
 *----------------------------------------------------------------------------*/

static int _cmp_int(const void *a, const void *b)
{ return (*(const int *)a > *(const int *)b) - (*(const int *)a < *(const int *)b); }

static void
_msr_from_native(ora_grid_t *g, const double *da)
{
  const int n = g->n_rows, nf = g->n_faces;
  const int *fc = g->face_cell;
  int *ri = calloc(n + 1, sizeof(int));
  for (int f = 0; f < nf; f++) {
    if (fc[2*f] < n) ri[fc[2*f]+1]++;
    if (fc[2*f+1] < n) ri[fc[2*f+1]+1]++;
  }
  for (int i = 0; i < n; i++) ri[i+1] += ri[i];
  const int nnz = ri[n];
  int *ci = malloc(sizeof(int)*(nnz > 0 ? nnz : 1));
  int *fill = malloc(sizeof(int)*(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) fill[i] = ri[i];
  for (int f = 0; f < nf; f++) {
    const int i = fc[2*f], j = fc[2*f+1];
    if (i < n) ci[fill[i]++] = j;
    if (j < n) ci[fill[j]++] = i;
  }
  for (int i = 0; i < n; i++)
    qsort(ci + ri[i], ri[i+1] - ri[i], sizeof(int), _cmp_int);
  free(fill);

  g->nnz = nnz;
  g->row_index = ri; g->col_id = ci;
  g->d_val = _ndup(da, sizeof(double)*n);
  g->x_val = calloc(nnz > 0 ? nnz : 1, sizeof(double));
  g->cell_face = malloc(sizeof(int)*(nnz > 0 ? nnz : 1));
  g->cell_face_sgn = malloc(sizeof(short)*(nnz > 0 ? nnz : 1));
  for (int f = 0; f < nf; f++) {
    const int i = fc[2*f], j = fc[2*f+1];
    if (i < n)
      for (int k = ri[i]; k < ri[i+1]; k++)
        if (ci[k] == j) { g->cell_face[k] = f; g->cell_face_sgn[k] = 1; g->x_val[k] = g->xa[f]; }
    if (j < n)
      for (int k = ri[j]; k < ri[j+1]; k++)
        if (ci[k] == i) { g->cell_face[k] = f; g->cell_face_sgn[k] = -1; g->x_val[k] = g->xa[f]; }
  }
}

/* Level 0 (cs_grid_create_from_shared, cs_grid.cpp:7425-7562, mapped-from-native branch) */
ora_grid_t *
ora_grid_create_fine_native(int n_rows, int n_faces, const int *face_cell, const double *da,
                            const double *xa, const double *cell_cen, const double *cell_vol,
                            const double *face_u_normal, const double *face_surf)
{
  ora_grid_t *g = calloc(1, sizeof(ora_grid_t));
  g->level = 0; g->n_rows = n_rows; g->n_cols_ext = n_rows; g->n_faces = n_faces;
  g->face_cell = _ndup(face_cell, sizeof(int)*2*n_faces);
  g->xa = _ndup(xa, sizeof(double)*n_faces);
  g->cell_cen = _ndup(cell_cen, sizeof(double)*3*n_rows);
  g->cell_vol = _ndup(cell_vol, sizeof(double)*n_rows);
  g->face_u_normal = _ndup(face_u_normal, sizeof(double)*3*n_faces);
  g->face_surf = _ndup(face_surf, sizeof(double)*n_faces);
  _msr_from_native(g, da);
  return g;
}

/*----------------------------------------------------------------------------
 * One coarsening step (cs_grid_coarsen, cs_grid.cpp:8099-8240, relaxation > 0).
 * n_threads == 1: _automatic_aggregation_fc; otherwise the MSR aggregation.
 *----------------------------------------------------------------------------*/

ora_grid_t *
ora_grid_coarsen_native(const ora_grid_t *f, int max_aggregation, double relax_param,
                        int n_threads)
{
  const int f_n_cells = f->n_rows, f_n_faces = f->n_faces;
  const int *f_face_cell = f->face_cell;
  const double *f_xa = f->xa, *f_da = f->d_val;

  ora_grid_t *c = calloc(1, sizeof(ora_grid_t));
  c->level = f->level + 1;
  c->coarse_row = malloc(sizeof(int)*(f_n_cells > 0 ? f_n_cells : 1));
  if (n_threads == 1)
    c->n_rows = ora_aggregation_fc(f->level, f_n_cells, f_n_cells, f_n_faces, f_face_cell,
                                   f_da, f_xa, max_aggregation, c->coarse_row);
  else
    c->n_rows = ora_aggregation_dx_msr(f->level, f_n_cells, f->row_index, f->col_id,
                                       f->d_val, f->x_val, max_aggregation, n_threads,
                                       c->coarse_row);
  const int c_n_cells = c->n_rows;
  c->n_cols_ext = c_n_cells;
  const int *coarse_row = c->coarse_row;

  int *coarse_face = malloc(sizeof(int)*(f_n_faces > 0 ? f_n_faces : 1));
  int *c_face_cell = malloc(sizeof(int)*2*(f_n_faces > 0 ? f_n_faces : 1));
  const int c_n_faces = ora_coarsen_face_cell(f_n_faces, f_face_cell, coarse_row, c_n_cells,
                                              coarse_face, c_face_cell);
  c->n_faces = c_n_faces;
  c->face_cell = realloc(c_face_cell, sizeof(int)*2*(c_n_faces > 0 ? c_n_faces : 1));
  c_face_cell = c->face_cell;

  /* _compute_coarse_cell_quantities (:4778-4820) */
  double *c_cell_vol = calloc(c_n_cells > 0 ? c_n_cells : 1, sizeof(double));
  double *c_cell_cen = calloc(3*(c_n_cells > 0 ? c_n_cells : 1), sizeof(double));
  for (int ii = 0; ii < f_n_cells; ii++) {
    const int ic = coarse_row[ii];
    if (ic > -1) {
      c_cell_vol[ic] += f->cell_vol[ii];
      c_cell_cen[3*ic]   += f->cell_vol[ii]*f->cell_cen[3*ii];
      c_cell_cen[3*ic+1] += f->cell_vol[ii]*f->cell_cen[3*ii+1];
      c_cell_cen[3*ic+2] += f->cell_vol[ii]*f->cell_cen[3*ii+2];
    }
  }
  for (int ic = 0; ic < c_n_cells; ic++) {
    c_cell_cen[3*ic]   /= c_cell_vol[ic];
    c_cell_cen[3*ic+1] /= c_cell_vol[ic];
    c_cell_cen[3*ic+2] /= c_cell_vol[ic];
  }
  c->cell_vol = c_cell_vol; c->cell_cen = c_cell_cen;

  /* _compute_coarse_quantities_native (:5144-5558), relax_param > 0, isym = 1 */
  const int ncf = (c_n_faces > 0 ? c_n_faces : 1);
  double *c_xa = calloc(ncf, sizeof(double));
  double *c_xa0 = calloc(ncf, sizeof(double));
  double *c_xa0ij = calloc(3*ncf, sizeof(double));
  double *c_face_normal = calloc(3*ncf, sizeof(double));

  if (relax_param <= 0) {
    for (int face_id = 0; face_id < f_n_faces; face_id++) {
      int c_face = coarse_face[face_id];
      if (c_face > 0) c_xa[c_face - 1] += f_xa[face_id];
      else if (c_face < 0) c_xa[-c_face - 1] += f_xa[face_id];
    }
  }
  else {
    for (int face_id = 0; face_id < f_n_faces; face_id++) {
      int c_face = coarse_face[face_id];
      if (c_face == 0) continue;
      double sgn;
      if (c_face > 0) { c_face -= 1; sgn = 1.; }
      else { c_face = -c_face - 1; sgn = -1.; }
      if (f->level == 0) {
        /* f_xa0 = xa; xa0ij and the face normal come from the mesh (:5303-5335) */
        const int ii = f_face_cell[2*face_id], jj = f_face_cell[2*face_id+1];
        double sdij[3];
        for (int kk = 0; kk < 3; kk++)
          sdij[kk] = sgn * (f->cell_cen[3*jj+kk] - f->cell_cen[3*ii+kk]);
        c_xa0[c_face] += f_xa[face_id];
        const double sgn_surf = sgn*f->face_surf[face_id];
        for (int k = 0; k < 3; k++)
          c_face_normal[3*c_face+k] += f->face_u_normal[3*face_id+k] * sgn_surf;
        c_xa0ij[3*c_face]   += f_xa[face_id] * sdij[0];
        c_xa0ij[3*c_face+1] += f_xa[face_id] * sdij[1];
        c_xa0ij[3*c_face+2] += f_xa[face_id] * sdij[2];
      }
      else {
        c_xa0[c_face] += f->xa0[face_id];
        for (int k = 0; k < 3; k++)
          c_face_normal[3*c_face+k] += sgn*f->face_normal[3*face_id+k];
        c_xa0ij[3*c_face]   += sgn*f->xa0ij[3*face_id];
        c_xa0ij[3*c_face+1] += sgn*f->xa0ij[3*face_id+1];
        c_xa0ij[3*c_face+2] += sgn*f->xa0ij[3*face_id+2];
      }
    }

    for (int c_face = 0; c_face < c_n_faces; c_face++)
      c_xa[c_face] = c_xa0[c_face];

    for (int c_face = 0; c_face < c_n_faces; c_face++) {
      const int ic = c_face_cell[2*c_face], jc = c_face_cell[2*c_face+1];
      const double dsigjg
        =   (c_cell_cen[3*jc]   - c_cell_cen[3*ic])   * c_face_normal[3*c_face]
          + (c_cell_cen[3*jc+1] - c_cell_cen[3*ic+1]) * c_face_normal[3*c_face+1]
          + (c_cell_cen[3*jc+2] - c_cell_cen[3*ic+2]) * c_face_normal[3*c_face+2];
      const double dsxaij
        =   c_xa0ij[3*c_face]   * c_face_normal[3*c_face]
          + c_xa0ij[3*c_face+1] * c_face_normal[3*c_face+1]
          + c_xa0ij[3*c_face+2] * c_face_normal[3*c_face+2];
      if (fabs(dsigjg) > 1.e-12) {               /* cs_math_epzero */
        const double agij = dsxaij/dsigjg;
        c_xa[c_face] = agij;
        if (agij < c_xa0[c_face] || agij > 0.)
          c_xa[c_face] = c_xa0[c_face];
      }
    }
    for (int c_face = 0; c_face < c_n_faces; c_face++)
      c_xa[c_face] = relax_param*c_xa[c_face] + (1. - relax_param)*c_xa0[c_face];
  }

  /* diagonal: non-differential part of the fine rows, then the coarse faces (:5455-5534) */
  double *w1 = malloc(sizeof(double)*(f_n_cells > 0 ? f_n_cells : 1));
  for (int ii = 0; ii < f_n_cells; ii++) w1[ii] = f_da[ii];
  for (int face_id = 0; face_id < f_n_faces; face_id++) {
    w1[f_face_cell[2*face_id]]   += f_xa[face_id];
    w1[f_face_cell[2*face_id+1]] += f_xa[face_id];
  }
  double *c_da = calloc(c_n_cells > 0 ? c_n_cells : 1, sizeof(double));
  for (int ii = 0; ii < f_n_cells; ii++) {
    const int ic = coarse_row[ii];
    if (ic > -1 && w1[ii] > 0.) c_da[ic] += w1[ii];
  }
  for (int c_face = 0; c_face < c_n_faces; c_face++) {
    c_da[c_face_cell[2*c_face]]   -= c_xa[c_face];
    c_da[c_face_cell[2*c_face+1]] -= c_xa[c_face];
  }
  free(w1);
  free(coarse_face);

  c->xa = c_xa; c->xa0 = c_xa0; c->xa0ij = c_xa0ij; c->face_normal = c_face_normal;
  _msr_from_native(c, c_da);
  free(c_da);
  return c;
}

/* _setup_hierarchy stop criteria (cs_multigrid.cpp:2454-2584) over the native steps */
ora_mg_t *
ora_mg_setup_native(ora_grid_t *fine, int aggregation_limit, int n_levels_max,
                    int n_g_rows_min, double p0p1_relax, int n_threads)
{
  ora_mg_t *mg = calloc(1, sizeof(ora_mg_t));
  mg->grids = calloc(n_levels_max > 0 ? n_levels_max : 1, sizeof(ora_grid_t *));
  mg->grids[0] = fine;
  mg->n_levels = 1;
  ora_grid_t *g = fine;
  long n_g_rows = g->n_rows, n_g_rows_prev;
  while (1) {
    n_g_rows_prev = n_g_rows;
    if (mg->n_levels >= n_levels_max) break;
    ora_grid_t *c;
    if (p0p1_relax > 0)
      c = ora_grid_coarsen_native(g, aggregation_limit, p0p1_relax, n_threads);
    else if (mg->n_levels == 1 && n_threads == 1) {
      /* relaxation <= 0: Galerkin values (cs_grid.cpp:8242-8246) and no faces below level 0
         (c->use_faces = false, :8055), but level 0 itself is still aggregated over its
         face list (:8106-8114) */
      int *f_c = malloc(sizeof(int)*(g->n_rows > 0 ? g->n_rows : 1));
      const int nc = ora_aggregation_fc(0, g->n_rows, g->n_rows, g->n_faces, g->face_cell,
                                        g->d_val, g->xa, aggregation_limit, f_c);
      c = ora_grid_coarsen_galerkin_map(g, f_c, nc);
      free(f_c);
    }
    else
      c = ora_grid_coarsen(g, aggregation_limit, p0p1_relax, n_threads);
    n_g_rows = c->n_rows;
    int add = 1;
    if (n_g_rows == 0) add = 0;
    else if (mg->n_levels > 1) {
      if (n_g_rows < n_g_rows_min || n_g_rows > (0.8 * n_g_rows_prev)) add = 0;
    }
    if (!add) { ora_grid_free(c); break; }
    mg->grids[mg->n_levels++] = c;
    g = c;
  }
  return mg;
}
