/*============================================================================
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * Thin extern "C" driver around the UNMODIFIED reference (code_saturne)
 * CPU implementation of the src/alge linear-solve path, linked from the
 * objects oracle/Makefile.ref compiles in place from /root/reference.
 *
 * It only calls the reference's public (and cs_grid_priv.h) API:
 *   cs_matrix_structure_create_msr_shared / cs_matrix_create /
 *   cs_matrix_transfer_coefficients_msr / cs_matrix_set_mesh_association
 *     (src/alge/cs_matrix.cpp:4518, :4682, :5451, :6328)
 *   cs_matrix_vector_multiply            (src/alge/cs_matrix.cpp:6405)
 *   cs_sles_it_create/_setup/_solve      (src/alge/cs_sles_it.cpp:4659,:4964,:5194)
 *   cs_multigrid_create/_setup/_solve    (src/alge/cs_multigrid.cpp:4644,:5275,:5362)
 *   cs_grid_create_from_shared / cs_grid_coarsen
 *                                        (src/alge/cs_grid.cpp:7425, :7980)
 *   cs_dot_xx, cs_dot_xx_xy_yz ...       (src/alge/cs_blas.cpp:1629-1905)
 *   cs_multigrid_smoother_create/_setup/_solve
 *                                        (src/alge/cs_multigrid_smoother.cpp:1726,:1825,:2008)
 *   cs_multigrid_pc_create, cs_sles_it_transfer_pc
 *                                        (src/alge/cs_multigrid.cpp:5682, cs_sles_it.cpp)
 *
 * Used (a) to pin oracle/*.c restatements and generate tests/golden fixtures,
 * (b) as the CPU baseline of bench.py ("kind": "reference").
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the resulting oracle/_ref/libcsref.so.
 *============================================================================*/

#include "base/cs_defs.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>
#include <omp.h>

#include "bft/bft_error.h"
#include "bft/bft_printf.h"
#include "base/cs_mem.h"
#include "base/cs_halo.h"
#include "base/cs_timer.h"
#include "mesh/cs_mesh.h"
#include "mesh/cs_mesh_adjacencies.h"
#include "mesh/cs_mesh_quantities.h"
#include "base/cs_numbering.h"
#include "base/cs_field.h"
#include "alge/cs_matrix_building.h"
#include "alge/cs_matrix_assembler.h"
#include "alge/cs_convection_diffusion.h"
#include "alge/cs_divergence.h"
#include "base/cs_parameters.h"
#include "base/cs_velocity_pressure.h"
#include "alge/cs_gradient.h"
#include "base/cs_boundary_conditions_set_coeffs.h"
#include "alge/cs_blas.h"
#include "alge/cs_matrix.h"
#include "alge/cs_matrix_default.h"
#include "alge/cs_sles.h"
#include "alge/cs_sles_it.h"
#include "alge/cs_sles_pc.h"
#include "alge/cs_multigrid.h"
#include "alge/cs_multigrid_smoother.h"
#include "alge/cs_grid.h"
#include "alge/cs_grid_priv.h"

namespace {

struct ref_system {
  cs_lnum_t  n_rows = 0;
  cs_lnum_t  n_faces = 0;
  bool       symmetric = true;
  cs_lnum_t *row_index = nullptr;
  cs_lnum_t *col_id = nullptr;
  cs_matrix_structure_t *ms = nullptr;
  cs_matrix_t           *m = nullptr;
  /* mesh association (owned copies) */
  cs_lnum_t    *c2f = nullptr;
  short int    *c2f_sgn = nullptr;
  cs_real_3_t  *cell_cen = nullptr;
  cs_real_t    *cell_vol = nullptr;
  cs_nreal_3_t *face_u_normal = nullptr;
  cs_real_t    *face_surf = nullptr;
};

struct ref_hierarchy {
  std::vector<cs_grid_t *> grids;
};

bool g_initialized = false;
bool g_quiet = true;

int _printf_quiet(const char *const format, va_list arg_ptr)
{
  if (g_quiet) return 0;
  return vprintf(format, arg_ptr);
}

double _now()
{
  return std::chrono::duration<double>(
    std::chrono::steady_clock::now().time_since_epoch()).count();
}

void _set_glob_mesh_faces(const ref_system *s, const cs_lnum_t *face_cells)
{
  /* cs_multigrid_setup reads cs_glob_mesh->n_i_faces / i_face_cells
     (src/alge/cs_multigrid.cpp:5283,:5312). */
  if (cs_glob_mesh == nullptr)
    cs_glob_mesh = cs_mesh_create();
  cs_mesh_t *mesh = cs_glob_mesh;
  CS_FREE(mesh->i_face_cells);
  mesh->n_cells = s->n_rows;
  mesh->n_cells_with_ghosts = s->n_rows;
  mesh->n_i_faces = s->n_faces;
  CS_MALLOC(mesh->i_face_cells, (s->n_faces > 0) ? s->n_faces : 1, cs_lnum_2_t);
  if (face_cells != nullptr)
    memcpy(mesh->i_face_cells, face_cells, sizeof(cs_lnum_2_t)*s->n_faces);
}

} // namespace

extern "C" {

/* Initialise reference globals; n_threads <= 0 keeps the OpenMP default. */
int
csref_init(int n_threads, int quiet)
{
  if (n_threads > 0)
    omp_set_num_threads(n_threads);
  cs_glob_n_threads = omp_get_max_threads();
  g_quiet = (quiet != 0);
  if (!g_initialized) {
    bft_printf_proxy_set(_printf_quiet);
    g_initialized = true;
  }
  return cs_glob_n_threads;
}

/* The reference's cs_matrix_t of a system (ref_harness_b200.cpp hands it to cs_sles_solve). */
void *
csref_system_matrix(void *sys)
{
  return static_cast<ref_system *>(sys)->m;
}

/* Build a reference MSR cs_matrix_t from plain arrays (values are copied). */
void *
csref_system_create(int            n_rows,
                    const int     *row_index,
                    const int     *col_id,
                    const double  *d_val,
                    const double  *x_val,
                    int            symmetric)
{
  ref_system *s = new ref_system;
  s->n_rows = n_rows;
  s->symmetric = (symmetric != 0);
  const cs_lnum_t nnz = row_index[n_rows];

  CS_MALLOC(s->row_index, n_rows+1, cs_lnum_t);
  CS_MALLOC(s->col_id, (nnz > 0) ? nnz : 1, cs_lnum_t);
  memcpy(s->row_index, row_index, sizeof(cs_lnum_t)*(n_rows+1));
  memcpy(s->col_id, col_id, sizeof(cs_lnum_t)*nnz);

  s->ms = cs_matrix_structure_create_msr_shared(true, n_rows, n_rows,
                                                s->row_index, s->col_id,
                                                nullptr, nullptr);
  s->m = cs_matrix_create(s->ms);

  cs_real_t *d, *x;
  CS_MALLOC(d, n_rows, cs_real_t);
  CS_MALLOC(x, (nnz > 0) ? nnz : 1, cs_real_t);
  memcpy(d, d_val, sizeof(cs_real_t)*n_rows);
  memcpy(x, x_val, sizeof(cs_real_t)*nnz);
  cs_matrix_transfer_coefficients_msr(s->m, s->symmetric, 1, 1,
                                      s->row_index, s->col_id, &d, &x);
  return s;
}

/* MSR matrix with db x db diagonal blocks and scalar extra-diagonal entries
   (cs_matrix_transfer_coefficients_msr with db_size = db, eb_size = 1; fill type
   CS_MATRIX_BLOCK_D).  d_val: n_rows*db*db, row-major blocks. */
void *
csref_system_create_block(int            n_rows,
                          int            db,
                          const int     *row_index,
                          const int     *col_id,
                          const double  *d_val,
                          const double  *x_val,
                          int            symmetric)
{
  ref_system *s = new ref_system;
  s->n_rows = n_rows;
  s->symmetric = (symmetric != 0);
  const cs_lnum_t nnz = row_index[n_rows];

  CS_MALLOC(s->row_index, n_rows+1, cs_lnum_t);
  CS_MALLOC(s->col_id, (nnz > 0) ? nnz : 1, cs_lnum_t);
  memcpy(s->row_index, row_index, sizeof(cs_lnum_t)*(n_rows+1));
  memcpy(s->col_id, col_id, sizeof(cs_lnum_t)*nnz);

  s->ms = cs_matrix_structure_create_msr_shared(true, n_rows, n_rows,
                                                s->row_index, s->col_id,
                                                nullptr, nullptr);
  s->m = cs_matrix_create(s->ms);

  cs_real_t *d, *x;
  CS_MALLOC(d, (size_t)n_rows*db*db, cs_real_t);
  CS_MALLOC(x, (nnz > 0) ? nnz : 1, cs_real_t);
  memcpy(d, d_val, sizeof(cs_real_t)*(size_t)n_rows*db*db);
  memcpy(x, x_val, sizeof(cs_real_t)*nnz);
  cs_matrix_transfer_coefficients_msr(s->m, s->symmetric, db, 1,
                                      s->row_index, s->col_id, &d, &x);
  return s;
}

/* CS_MATRIX_CSR from an edge list (the reference's native way to a CSR matrix:
   cs_matrix_structure_create + cs_matrix_set_coefficients, cs_matrix.cpp:4328, :5315).
   xa: one value per edge (symmetric) or two (lower-to-upper, upper-to-lower). */
void *
csref_system_create_csr(int            n_rows,
                        int            n_edges,
                        const int     *edges,
                        const double  *da,
                        const double  *xa,
                        int            symmetric)
{
  ref_system *s = new ref_system;
  s->n_rows = n_rows;
  s->symmetric = (symmetric != 0);
  s->row_index = nullptr;
  s->col_id = nullptr;
  s->ms = cs_matrix_structure_create(CS_MATRIX_CSR, n_rows, n_rows, n_edges,
                                     (const cs_lnum_2_t *)edges, nullptr, nullptr);
  s->m = cs_matrix_create(s->ms);
  cs_matrix_set_coefficients(s->m, s->symmetric, 1, 1, n_edges,
                             (const cs_lnum_2_t *)edges, da, xa);
  return s;
}

/* MSR matrix "mapped from native" (cs_matrix_structure_create(CS_MATRIX_MSR, faces) +
   cs_matrix_set_coefficients(edges, da, xa), cs_matrix.cpp:4328, :5315): the entry of a
   caller that holds face-based arrays.  cs_grid_create_from_shared then takes its
   use_faces / use_xa branch (cs_grid.cpp:7485-7491), so cs_grid_coarsen runs
   _automatic_aggregation_fc (one thread) or _automatic_aggregation_dx_msr, then
   _coarsen_face_cell, _compute_coarse_cell_quantities, _compute_coarse_quantities_native
   and _matrix_from_native.  da, xa and the geometry are kept (the matrix maps them). */
struct ref_native_arrays {
  cs_lnum_2_t *face_cells; cs_real_t *da, *xa;
};
static std::vector<std::pair<void *, ref_native_arrays>> g_native;

void *
csref_system_create_native(int            n_rows,
                           int            n_faces,
                           const int     *face_cells,
                           const double  *da,
                           const double  *xa,
                           const double  *cell_cen,
                           const double  *cell_vol,
                           const double  *face_u_normal,
                           const double  *face_surf)
{
  ref_system *s = new ref_system;
  s->n_rows = n_rows;
  s->n_faces = n_faces;
  s->symmetric = true;
  ref_native_arrays na;
  CS_MALLOC(na.face_cells, n_faces > 0 ? n_faces : 1, cs_lnum_2_t);
  CS_MALLOC(na.da, n_rows, cs_real_t);
  CS_MALLOC(na.xa, n_faces > 0 ? n_faces : 1, cs_real_t);
  memcpy(na.face_cells, face_cells, sizeof(cs_lnum_2_t)*n_faces);
  memcpy(na.da, da, sizeof(cs_real_t)*n_rows);
  memcpy(na.xa, xa, sizeof(cs_real_t)*n_faces);
  s->ms = cs_matrix_structure_create(CS_MATRIX_MSR, n_rows, n_rows, n_faces,
                                     na.face_cells, nullptr, nullptr);
  s->m = cs_matrix_create(s->ms);
  cs_matrix_set_coefficients(s->m, true, 1, 1, n_faces, na.face_cells, na.da, na.xa);

  CS_MALLOC(s->cell_cen, n_rows, cs_real_3_t);
  CS_MALLOC(s->cell_vol, n_rows, cs_real_t);
  CS_MALLOC(s->face_u_normal, n_faces > 0 ? n_faces : 1, cs_nreal_3_t);
  CS_MALLOC(s->face_surf, n_faces > 0 ? n_faces : 1, cs_real_t);
  memcpy(s->cell_cen, cell_cen, sizeof(cs_real_3_t)*n_rows);
  memcpy(s->cell_vol, cell_vol, sizeof(cs_real_t)*n_rows);
  for (cs_lnum_t i = 0; i < n_faces; i++) {
    for (int k = 0; k < 3; k++)
      s->face_u_normal[i][k] = face_u_normal[i*3+k];
    s->face_surf[i] = face_surf[i];
  }
  cs_matrix_set_mesh_association(s->m, nullptr, nullptr, nullptr,
                                 s->cell_cen, s->cell_vol,
                                 s->face_u_normal, s->face_surf);
  _set_glob_mesh_faces(s, face_cells);
  g_native.push_back({s, na});
  return s;
}

/* Matrix built through the reference's assembler on one rank
   (cs_matrix_assembler_create / _add_g_ids / _compute, cs_matrix_assembler.cpp:2366, :2673,
   :2734; cs_matrix_create_from_assembler; cs_matrix_assembler_values_init / _add_g /
   _finalize, cs_matrix.cpp:5635, cs_matrix_assembler.cpp:3276, :3119). */
struct ref_assembler {
  cs_matrix_assembler_t *ma; cs_matrix_assembler_values_t *mav;
};
static std::vector<std::pair<void *, ref_assembler>> g_assemblers;

void *
csref_assembler_system_create(unsigned long long        l0,
                              int                       n_rows,
                              int                       n,
                              const unsigned long long *row_g_id,
                              const unsigned long long *col_g_id)
{
  ref_system *s = new ref_system;
  s->n_rows = n_rows;
  s->symmetric = false;
  cs_gnum_t l_range[2] = {(cs_gnum_t)l0, (cs_gnum_t)(l0 + n_rows)};
  ref_assembler ra;
  ra.ma = cs_matrix_assembler_create(l_range, true);
  cs_matrix_assembler_add_g_ids(ra.ma, n, (const cs_gnum_t *)row_g_id,
                                (const cs_gnum_t *)col_g_id);
  cs_matrix_assembler_compute(ra.ma);
  s->m = cs_matrix_create_from_assembler(CS_MATRIX_MSR, ra.ma);
  ra.mav = cs_matrix_assembler_values_init(s->m, 1, 1);
  g_assemblers.push_back({s, ra});
  return s;
}

static ref_assembler *_find_assembler(void *sys)
{
  for (size_t i = 0; i < g_assemblers.size(); i++)
    if (g_assemblers[i].first == sys) return &(g_assemblers[i].second);
  return nullptr;
}

void
csref_assembler_values_add(void                     *sys,
                           int                       n,
                           const unsigned long long *row_g_id,
                           const unsigned long long *col_g_id,
                           const double             *val)
{
  ref_assembler *ra = _find_assembler(sys);
  cs_matrix_assembler_values_add_g(ra->mav, n, (const cs_gnum_t *)row_g_id,
                                   (const cs_gnum_t *)col_g_id, val);
}

/* returns the number of extra-diagonal entries */
int
csref_assembler_values_finalize(void *sys)
{
  ref_system *s = static_cast<ref_system *>(sys);
  ref_assembler *ra = _find_assembler(sys);
  cs_matrix_assembler_values_finalize(&(ra->mav));
  const cs_lnum_t *ri, *ci; const cs_real_t *d, *x;
  cs_matrix_get_msr_arrays(s->m, &ri, &ci, &d, &x);
  return ri[s->n_rows];
}

/* MSR arrays of any system (copies) */
void
csref_get_msr_arrays(void *sys, int *row_index, int *col_id, double *d_val, double *x_val)
{
  ref_system *s = static_cast<ref_system *>(sys);
  const cs_lnum_t *ri, *ci; const cs_real_t *d, *x;
  cs_matrix_get_msr_arrays(s->m, &ri, &ci, &d, &x);
  const cs_lnum_t n = s->n_rows, nnz = ri[n];
  if (row_index) memcpy(row_index, ri, sizeof(cs_lnum_t)*(n+1));
  if (col_id) memcpy(col_id, ci, sizeof(cs_lnum_t)*nnz);
  if (d_val) memcpy(d_val, d, sizeof(cs_real_t)*n);
  if (x_val) memcpy(x_val, x, sizeof(cs_real_t)*nnz);
}

/* Copy the CSR arrays out; returns the number of entries (call with null pointers
   first to size the buffers). */
int
csref_get_csr_arrays(void *sys, int *row_index, int *col_id, double *val)
{
  ref_system *s = static_cast<ref_system *>(sys);
  const cs_lnum_t *ri = nullptr, *ci = nullptr;
  const cs_real_t *v = nullptr;
  cs_matrix_get_csr_arrays(s->m, &ri, &ci, &v);
  if (ri == nullptr) return -1;
  const int nnz = ri[s->n_rows];
  if (row_index != nullptr) memcpy(row_index, ri, sizeof(int)*(s->n_rows + 1));
  if (col_id != nullptr) memcpy(col_id, ci, sizeof(int)*nnz);
  if (val != nullptr) memcpy(val, v, sizeof(double)*nnz);
  return nnz;
}

/* Associate mesh data (same role as src/alge/cs_sles_default.cpp:826-838). */
void
csref_system_set_mesh(void          *sys,
                      int            n_faces,
                      const int     *face_cells,    /* 2*n_faces */
                      const int     *c2f,           /* nnz: face id per entry */
                      const short   *c2f_sgn,       /* nnz */
                      const double  *cell_cen,      /* 3*n_rows */
                      const double  *cell_vol,      /* n_rows */
                      const double  *face_u_normal, /* 3*n_faces */
                      const double  *face_surf)     /* n_faces */
{
  ref_system *s = static_cast<ref_system *>(sys);
  const cs_lnum_t nnz = s->row_index[s->n_rows];
  s->n_faces = n_faces;

  CS_MALLOC(s->c2f, nnz, cs_lnum_t);
  CS_MALLOC(s->c2f_sgn, nnz, short int);
  CS_MALLOC(s->cell_cen, s->n_rows, cs_real_3_t);
  CS_MALLOC(s->cell_vol, s->n_rows, cs_real_t);
  CS_MALLOC(s->face_u_normal, n_faces, cs_nreal_3_t);
  CS_MALLOC(s->face_surf, n_faces, cs_real_t);
  memcpy(s->c2f, c2f, sizeof(cs_lnum_t)*nnz);
  memcpy(s->c2f_sgn, c2f_sgn, sizeof(short int)*nnz);
  memcpy(s->cell_cen, cell_cen, sizeof(cs_real_3_t)*s->n_rows);
  memcpy(s->cell_vol, cell_vol, sizeof(cs_real_t)*s->n_rows);
  for (cs_lnum_t i = 0; i < n_faces; i++) {
    for (int k = 0; k < 3; k++)
      s->face_u_normal[i][k] = face_u_normal[i*3+k];
    s->face_surf[i] = face_surf[i];
  }

  cs_matrix_set_mesh_association(s->m, s->row_index, s->c2f, s->c2f_sgn,
                                 s->cell_cen, s->cell_vol,
                                 s->face_u_normal, s->face_surf);
  _set_glob_mesh_faces(s, face_cells);
}

void
csref_system_destroy(void *sys)
{
  ref_system *s = static_cast<ref_system *>(sys);
  if (s == nullptr) return;
  cs_matrix_destroy(&(s->m));
  if (s->ms != nullptr)
    cs_matrix_structure_destroy(&(s->ms));
  for (size_t i = 0; i < g_assemblers.size(); i++) {
    if (g_assemblers[i].first == sys) {
      if (g_assemblers[i].second.mav != nullptr)
        cs_matrix_assembler_values_finalize(&(g_assemblers[i].second.mav));
      cs_matrix_assembler_destroy(&(g_assemblers[i].second.ma));
      g_assemblers.erase(g_assemblers.begin() + i);
      break;
    }
  }
  for (size_t i = 0; i < g_native.size(); i++) {
    if (g_native[i].first == sys) {
      CS_FREE(g_native[i].second.face_cells);
      CS_FREE(g_native[i].second.da);
      CS_FREE(g_native[i].second.xa);
      g_native.erase(g_native.begin() + i);
      break;
    }
  }
  CS_FREE(s->row_index); CS_FREE(s->col_id);
  CS_FREE(s->c2f); CS_FREE(s->c2f_sgn);
  CS_FREE(s->cell_cen); CS_FREE(s->cell_vol);
  CS_FREE(s->face_u_normal); CS_FREE(s->face_surf);
  delete s;
}

/* y = A.x (or (A-D).x): reference _mat_vec_p_l_msr through the public entry. */
void
csref_spmv(void *sys, int exclude_diag, double *x, double *y)
{
  ref_system *s = static_cast<ref_system *>(sys);
  if (exclude_diag)
    cs_matrix_vector_multiply_partial(s->m, CS_MATRIX_SPMV_E, x, y);
  else
    cs_matrix_vector_multiply(s->m, x, y);
}

double csref_dot(int n, const double *x, const double *y)
{ return cs_dot(n, x, y); }

double csref_dot_xx(int n, const double *x)
{ return cs_dot_xx(n, x); }

void csref_dot_xx_xy_yz(int n, const double *x, const double *y,
                        const double *z, double *r)
{ cs_dot_xx_xy_yz(n, x, y, z, r, r+1, r+2); }

void csref_dot_xx_xy(int n, const double *x, const double *y, double *r)
{ cs_dot_xx_xy(n, x, y, r, r+1); }

void csref_dot_xy_yz(int n, const double *x, const double *y,
                     const double *z, double *r)
{ cs_dot_xy_yz(n, x, y, z, r, r+1); }

/* cs_sles_it solve. type: cs_sles_it_type_t value (0 PCG, 3 Jacobi, 4 BiCGStab
   ...). vx holds the initial guess when use_vx_ini != 0 and the result.
   out[0] = n_iter, out[1] = residual, out[2] = initial residual,
   out[3] = setup seconds, out[4] = solve seconds. Returns convergence state. */
int
csref_sles_it_solve(void          *sys,
                    int            type,
                    int            poly_degree,
                    int            n_max_iter,
                    double         precision,
                    double         r_norm,
                    const double  *rhs,
                    double        *vx,
                    int            use_vx_ini,
                    double        *out)
{
  ref_system *s = static_cast<ref_system *>(sys);
  void *c = cs_sles_it_create((cs_sles_it_type_t)type, poly_degree,
                              n_max_iter, false);
  double t0 = _now();
  cs_sles_it_setup(c, "csref", s->m, 0);
  double t1 = _now();
  int n_iter = 0;
  double residual = -1;
  cs_sles_convergence_state_t cvg
    = cs_sles_it_solve(c, "csref", s->m, -1, precision, r_norm,
                       &n_iter, &residual, rhs,
                       use_vx_ini ? vx : nullptr, vx, 0, nullptr);
  double t2 = _now();
  out[0] = n_iter;
  out[1] = residual;
  out[2] = cs_sles_it_get_last_initial_residual((const cs_sles_it_t *)c);
  out[3] = t1 - t0;
  out[4] = t2 - t1;
  cs_sles_it_free(c);
  cs_sles_it_destroy(&c);
  return (int)cvg;
}

/* cs_multigrid V-cycle solve (solver mode).
   iopt: [0..2] descent/ascent/coarse solver type, [3] n_max_cycles,
         [4..6] n_max_iter, [7..9] poly_degree, [10] aggregation_limit,
         [11] coarsening_type, [12] n_max_levels, [13] min_g_rows,
         [14] cs_multigrid_type_t (0 V-cycle, 1 K-cycle)
   dopt: [0..2] precision_mult, [3] p0p1_relax
   out : [0] n_equiv_iter, [1] residual, [2] setup s, [3] solve s */
int
csref_multigrid_solve(void          *sys,
                      const int     *iopt,
                      const double  *dopt,
                      double         precision,
                      double         r_norm,
                      const double  *rhs,
                      double        *vx,
                      int            use_vx_ini,
                      int            verbosity,
                      double        *out)
{
  ref_system *s = static_cast<ref_system *>(sys);
  cs_multigrid_t *mg = cs_multigrid_create((cs_multigrid_type_t)iopt[14], nullptr);
  cs_multigrid_set_solver_options(mg,
                                  (cs_sles_it_type_t)iopt[0],
                                  (cs_sles_it_type_t)iopt[1],
                                  (cs_sles_it_type_t)iopt[2],
                                  iopt[3], iopt[4], iopt[5], iopt[6],
                                  iopt[7], iopt[8], iopt[9],
                                  dopt[0], dopt[1], dopt[2]);
  cs_multigrid_set_coarsening_options(mg, iopt[10],
                                      (cs_grid_coarsening_t)iopt[11],
                                      iopt[12], (cs_gnum_t)iopt[13],
                                      dopt[3], 0);
  bool q = g_quiet;
  if (verbosity > 0) g_quiet = false;
  double t0 = _now();
  cs_multigrid_setup(mg, "csref", s->m, verbosity);
  double t1 = _now();
  int n_iter = 0;
  double residual = -1;
  cs_sles_convergence_state_t cvg
    = cs_multigrid_solve(mg, "csref", s->m, verbosity, precision, r_norm,
                         &n_iter, &residual, rhs,
                         use_vx_ini ? vx : nullptr, vx, 0, nullptr);
  double t2 = _now();
  g_quiet = q;
  out[0] = n_iter;
  out[1] = residual;
  out[2] = t1 - t0;
  out[3] = t2 - t1;
  cs_multigrid_free(mg);
  void *ctx = mg;
  cs_multigrid_destroy(&ctx);
  return (int)cvg;
}

/* cs_multigrid_smoother solve (fixed sweep count, no convergence test).
   type: cs_sles_it_type_t value (3 Jacobi, 13 L1 Jacobi, 14 relaxed Jacobi,
   15 RJ2, 16 RJ3).  out[0] = n_iter.  Returns the convergence state. */
int
csref_smoother_solve(void          *sys,
                     int            type,
                     int            n_max_iter,
                     const double  *rhs,
                     double        *vx,
                     int            use_vx_ini,
                     double        *out)
{
  ref_system *s = static_cast<ref_system *>(sys);
  cs_sles_it_t *c = cs_multigrid_smoother_create((cs_sles_it_type_t)type, 0,
                                                 n_max_iter);
  if (c == nullptr)
    return -100;
  cs_multigrid_smoother_setup(c, "csref", s->m, 0);
  int n_iter = 0;
  double residual = -1;
  cs_sles_convergence_state_t cvg
    = cs_multigrid_smoother_solve(c, "csref", s->m, -1, 0., 1.,
                                  &n_iter, &residual, rhs,
                                  use_vx_ini ? vx : nullptr, vx, 0, nullptr);
  out[0] = n_iter;
  void *ctx = c;
  cs_sles_it_free(ctx);
  cs_sles_it_destroy(&ctx);
  return (int)cvg;
}

/* FCG (solver_type 1; the reference's production pressure solver layout) or PCG (0,
   _conjugate_gradient_sr) preconditioned by one V-cycle (cs_multigrid_pc_create +
   cs_sles_it_transfer_pc).
   iopt/dopt as csref_multigrid_solve ([3] n_max_cycles is forced to 1).
   out: [0] n_iter, [1] residual, [2] initial residual, [3] setup s, [4] solve s */
int
csref_it_mgpc_solve(void          *sys,
                    int            solver_type,
                    const int     *iopt,
                     const double  *dopt,
                     int            n_max_iter,
                     double         precision,
                     double         r_norm,
                     const double  *rhs,
                     double        *vx,
                     int            use_vx_ini,
                     double        *out)
{
  ref_system *s = static_cast<ref_system *>(sys);
  cs_sles_it_t *c = cs_sles_it_create((cs_sles_it_type_t)solver_type, -1,
                                      n_max_iter, false);
  cs_sles_pc_t *pc = cs_multigrid_pc_create((cs_multigrid_type_t)iopt[14]);
  cs_multigrid_t *mg = (cs_multigrid_t *)cs_sles_pc_get_context(pc);
  cs_multigrid_set_solver_options(mg,
                                  (cs_sles_it_type_t)iopt[0],
                                  (cs_sles_it_type_t)iopt[1],
                                  (cs_sles_it_type_t)iopt[2],
                                  1, iopt[4], iopt[5], iopt[6],
                                  iopt[7], iopt[8], iopt[9],
                                  dopt[0], dopt[1], dopt[2]);
  cs_multigrid_set_coarsening_options(mg, iopt[10],
                                      (cs_grid_coarsening_t)iopt[11],
                                      iopt[12], (cs_gnum_t)iopt[13],
                                      dopt[3], 0);
  cs_sles_it_transfer_pc(c, &pc);
  double t0 = _now();
  cs_sles_it_setup(c, "csref", s->m, 0);
  double t1 = _now();
  int n_iter = 0;
  double residual = -1;
  cs_sles_convergence_state_t cvg
    = cs_sles_it_solve(c, "csref", s->m, -1, precision, r_norm,
                       &n_iter, &residual, rhs,
                       use_vx_ini ? vx : nullptr, vx, 0, nullptr);
  double t2 = _now();
  out[0] = n_iter;
  out[1] = residual;
  out[2] = cs_sles_it_get_last_initial_residual(c);
  out[3] = t1 - t0;
  out[4] = t2 - t1;
  void *ctx = c;
  cs_sles_it_free(ctx);
  cs_sles_it_destroy(&ctx);
  return (int)cvg;
}

int
csref_fcg_mgpc_solve(void *sys, const int *iopt, const double *dopt, int n_max_iter,
                     double precision, double r_norm, const double *rhs, double *vx,
                     int use_vx_ini, double *out)
{
  return csref_it_mgpc_solve(sys, (int)CS_SLES_FCG, iopt, dopt, n_max_iter, precision,
                             r_norm, rhs, vx, use_vx_ini, out);
}

/* cs_matrix_compute_coeffs, scalar cases -- symmetric (iconvp = 0, _sym_coeffs_msr) and
   convection-diffusion (iconvp = 1, _coeffs_msr<is_thermal>, :407-520; the system must have
   been created non-symmetric): the reference fills the MSR
   coefficients of the system's matrix from face viscosities, the implicit boundary
   coefficient and the unsteady term (_sym_coeffs_msr, src/alge/cs_matrix_building.cpp:136-
   206; relaxation, the (1 + 1e-7) shift when there is no Dirichlet condition :2243-2262).
   The mesh globals it reads (cs_glob_mesh faces, cs_glob_mesh_adjacencies,
   cs_glob_mesh_quantities) are set up from the system's face -> cell table.
   d_out[n_rows], x_out[nnz] receive the MSR coefficients. */
int
csref_matrix_compute_coeffs(void          *sys,
                            int            n_b_faces,
                            const int     *b_face_cells,
                            int            iconvp,
                            int            idiffp,
                            int            ndircp,
                            double         thetap,
                            double         relaxp,
                            const double  *coefbp,
                            const double  *cofbfp,
                            const double  *rovsdt,
                            const double  *i_massflux,
                            const double  *b_massflux,
                            const double  *i_visc,
                            const double  *b_visc,
                            const double  *xcpp,     /* NULL: imucpp = 0 */
                            double        *d_out,
                            double        *x_out)
{
  ref_system *s = static_cast<ref_system *>(sys);
  cs_mesh_t *mesh = cs_glob_mesh;
  if (mesh == nullptr || mesh->n_i_faces != s->n_faces)
    return -1;                      /* csref_system_set_mesh first */

  CS_FREE(mesh->b_face_cells);
  mesh->n_b_faces = n_b_faces;
  mesh->n_b_faces_all = n_b_faces;
  CS_MALLOC(mesh->b_face_cells, (n_b_faces > 0) ? n_b_faces : 1, cs_lnum_t);
  memcpy(mesh->b_face_cells, b_face_cells, sizeof(cs_lnum_t)*n_b_faces);
  if (mesh->b_face_numbering != nullptr)
    cs_numbering_destroy(&(mesh->b_face_numbering));
  mesh->b_face_numbering = cs_numbering_create_default(n_b_faces);
  if (mesh->i_face_numbering != nullptr)
    cs_numbering_destroy(&(mesh->i_face_numbering));
  mesh->i_face_numbering = cs_numbering_create_default(mesh->n_i_faces);

  if (cs_glob_mesh_quantities == nullptr)
    cs_glob_mesh_quantities = cs_mesh_quantities_create();
  cs_mesh_adjacencies_finalize();
  cs_mesh_adjacencies_initialize();
  cs_mesh_adjacencies_update_mesh();

  cs_field_bc_coeffs_t bc;
  memset(&bc, 0, sizeof(bc));
  bc.bf = const_cast<cs_real_t *>(cofbfp);
  bc.b = const_cast<cs_real_t *>(coefbp);

  cs_matrix_compute_coeffs(s->m, nullptr, iconvp, idiffp, ndircp, thetap, relaxp,
                           (xcpp != nullptr) ? 1 : 0, &bc, rovsdt, i_massflux, b_massflux,
                           i_visc, b_visc, xcpp);

  const cs_lnum_t *ri, *ci;
  const cs_real_t *d, *x;
  cs_matrix_get_msr_arrays(s->m, &ri, &ci, &d, &x);
  memcpy(d_out, d, sizeof(double)*s->n_rows);
  memcpy(x_out, x, sizeof(double)*ri[s->n_rows]);
  return 0;
}

/* cs_matrix_compute_coeffs<3> (vector variable, scalar diffusion, eb_size 1) on a system
   created with 3x3 diagonal blocks (csref_system_create_block + csref_system_set_mesh);
   mesh globals as set by a previous csref_matrix_compute_coeffs call are rebuilt here too.
   coefbp / cofbfp: n_b_faces x 9; fimp: n_rows x 9 or NULL; d_out: n_rows x 9. */
int
csref_matrix_compute_coeffs_v3(void          *sys,
                               int            n_b_faces,
                               const int     *b_face_cells,
                               int            iconvp,
                               int            idiffp,
                               int            ndircp,
                               double         thetap,
                               double         relaxp,
                               const double  *coefbp,
                               const double  *cofbfp,
                               const double  *fimp,
                               const double  *i_massflux,
                               const double  *b_massflux,
                               const double  *i_visc,
                               const double  *b_visc,
                               double        *d_out,
                               double        *x_out)
{
  ref_system *s = static_cast<ref_system *>(sys);
  cs_mesh_t *mesh = cs_glob_mesh;
  if (mesh == nullptr || mesh->n_i_faces != s->n_faces)
    return -1;

  CS_FREE(mesh->b_face_cells);
  mesh->n_b_faces = n_b_faces;
  mesh->n_b_faces_all = n_b_faces;
  CS_MALLOC(mesh->b_face_cells, (n_b_faces > 0) ? n_b_faces : 1, cs_lnum_t);
  memcpy(mesh->b_face_cells, b_face_cells, sizeof(cs_lnum_t)*n_b_faces);
  if (mesh->b_face_numbering != nullptr)
    cs_numbering_destroy(&(mesh->b_face_numbering));
  mesh->b_face_numbering = cs_numbering_create_default(n_b_faces);
  if (mesh->i_face_numbering != nullptr)
    cs_numbering_destroy(&(mesh->i_face_numbering));
  mesh->i_face_numbering = cs_numbering_create_default(mesh->n_i_faces);
  if (cs_glob_mesh_quantities == nullptr)
    cs_glob_mesh_quantities = cs_mesh_quantities_create();
  cs_mesh_adjacencies_finalize();
  cs_mesh_adjacencies_initialize();
  cs_mesh_adjacencies_update_mesh();

  cs_field_bc_coeffs_t bc;
  memset(&bc, 0, sizeof(bc));
  bc.bf = const_cast<cs_real_t *>(cofbfp);
  bc.b = const_cast<cs_real_t *>(coefbp);

  cs_matrix_compute_coeffs<3>(s->m, nullptr, iconvp, idiffp, 1, ndircp, 1, thetap, relaxp,
                              &bc, reinterpret_cast<const cs_real_t (*)[3][3]>(fimp),
                              i_massflux, b_massflux, i_visc, b_visc);

  const cs_lnum_t *ri, *ci;
  const cs_real_t *d, *x;
  cs_matrix_get_msr_arrays(s->m, &ri, &ci, &d, &x);
  memcpy(d_out, d, sizeof(double)*9*s->n_rows);
  memcpy(x_out, x, sizeof(double)*ri[s->n_rows]);
  return 0;
}

int
csref_matrix_compute_coeffs_sym(void *sys, int n_b_faces, const int *b_face_cells, int idiffp,
                                int ndircp, double thetap, double relaxp, const double *cofbfp,
                                const double *rovsdt, const double *i_visc,
                                const double *b_visc, double *d_out, double *x_out)
{
  return csref_matrix_compute_coeffs(sys, n_b_faces, b_face_cells, 0, idiffp, ndircp, thetap,
                                     relaxp, nullptr, cofbfp, rovsdt, nullptr, nullptr, i_visc,
                                     b_visc, nullptr, d_out, x_out);
}

/* cs_diffusion_potential (src/alge/cs_convection_diffusion.cpp:8743-8990) without flux
   reconstruction (eqp->ircflu = 0) on the mesh of a system with mesh association: the
   divergence of the diffusive flux of pvar, interior faces then boundary faces
   (flux_diff = inc*af + bf*pvar, cs_boundary_conditions_set_coeffs.cpp:172-185).
   Boundary geometry (b_geo: per boundary face its unit normal [3], centre [3], distance to
   the cell centre [1]) and i_face_cog feed the reference's boundary reconstruction
   (cs_gradient_boundary_iprime_lsq_s), whose result is multiplied by diipb = 0 here.
   diverg: n_rows values, read (init = 0) and written. */
int
csref_diffusion_potential(void          *sys,
                          const int     *face_cells,
                          int            n_b_faces,
                          const int     *b_face_cells,
                          const double  *b_geo,        /* 7 per boundary face */
                          const double  *i_face_cog,   /* 3 per interior face */
                          int            init,
                          int            inc,
                          const double  *pvar,
                          const double  *cofaf,
                          const double  *cofbf,
                          const double  *i_visc,
                          const double  *b_visc,
                          double        *diverg)
{
  ref_system *s = static_cast<ref_system *>(sys);
  _set_glob_mesh_faces(s, face_cells);
  cs_mesh_t *mesh = cs_glob_mesh;
  const cs_lnum_t n_cells = s->n_rows, n_i_faces = s->n_faces;

  CS_FREE(mesh->b_face_cells);
  mesh->n_b_faces = n_b_faces;
  mesh->n_b_faces_all = n_b_faces;
  CS_MALLOC(mesh->b_face_cells, (n_b_faces > 0) ? n_b_faces : 1, cs_lnum_t);
  memcpy(mesh->b_face_cells, b_face_cells, sizeof(cs_lnum_t)*n_b_faces);
  if (mesh->b_face_numbering != nullptr)
    cs_numbering_destroy(&(mesh->b_face_numbering));
  mesh->b_face_numbering = cs_numbering_create_default(n_b_faces);
  if (mesh->i_face_numbering != nullptr)
    cs_numbering_destroy(&(mesh->i_face_numbering));
  mesh->i_face_numbering = cs_numbering_create_default(n_i_faces);

  if (cs_glob_mesh_quantities == nullptr)
    cs_glob_mesh_quantities = cs_mesh_quantities_create();
  cs_mesh_quantities_t *mq = cs_glob_mesh_quantities;
  cs_mesh_adjacencies_finalize();
  cs_mesh_adjacencies_initialize();
  cs_mesh_adjacencies_update_mesh();

  /* geometry of an orthogonal mesh: II' = JJ' = 0 */
  std::vector<cs_real_t> b_dist(n_b_faces > 0 ? n_b_faces : 1), i_dist(n_i_faces > 0 ? n_i_faces : 1, 1.);
  std::vector<cs_nreal_t> b_un(3*(n_b_faces > 0 ? n_b_faces : 1));
  std::vector<cs_real_t> b_cog(3*(n_b_faces > 0 ? n_b_faces : 1));
  std::vector<cs_rreal_t> zero_b(3*(n_b_faces > 0 ? n_b_faces : 1), 0.),
                          zero_i(3*(n_i_faces > 0 ? n_i_faces : 1), 0.);
  for (cs_lnum_t f = 0; f < n_b_faces; f++) {
    for (int k = 0; k < 3; k++) {
      b_un[3*f+k] = b_geo[7*f+k];
      b_cog[3*f+k] = b_geo[7*f+3+k];
    }
    b_dist[f] = b_geo[7*f+6];
  }
  mq->cell_cen = (cs_real_3_t *)s->cell_cen;
  mq->b_face_u_normal = (cs_nreal_3_t *)b_un.data();
  mq->b_dist = b_dist.data();
  mq->i_dist = i_dist.data();
  mq->i_face_cog = (cs_real_3_t *)const_cast<double *>(i_face_cog);
  mq->b_face_cog = (cs_real_3_t *)b_cog.data();
  mq->i_face_surf = s->face_surf;
  mq->diipb = (cs_rreal_3_t *)zero_b.data();
  mq->diipf = (cs_rreal_3_t *)zero_i.data();
  mq->djjpf = (cs_rreal_3_t *)zero_i.data();

  cs_equation_param_t eqp = cs_parameters_equation_param_default();
  eqp.ircflu = 0;
  eqp.verbosity = 0;

  cs_field_bc_coeffs_t bc;
  memset(&bc, 0, sizeof(bc));
  std::vector<cs_real_t> a(n_b_faces > 0 ? n_b_faces : 1, 0.), b(n_b_faces > 0 ? n_b_faces : 1, 1.);
  bc.a = a.data(); bc.b = b.data();
  bc.af = const_cast<cs_real_t *>(cofaf);
  bc.bf = const_cast<cs_real_t *>(cofbf);

  std::vector<cs_real_t> _pvar(pvar, pvar + n_cells), visel(n_cells > 0 ? n_cells : 1, 1.);
  std::vector<cs_real_t> _div(n_cells > 0 ? n_cells : 1, 0.);
  if (!init) memcpy(_div.data(), diverg, sizeof(cs_real_t)*n_cells);

  cs_diffusion_potential(nullptr, &eqp, mesh, mq, init, inc, 0, nullptr, _pvar.data(), &bc,
                         i_visc, b_visc, visel.data(), _div.data());
  memcpy(diverg, _div.data(), sizeof(cs_real_t)*n_cells);

  /* the quantities point to this call's arrays */
  mq->cell_cen = nullptr; mq->b_face_u_normal = nullptr; mq->b_dist = nullptr;
  mq->i_dist = nullptr; mq->i_face_cog = nullptr; mq->b_face_cog = nullptr;
  mq->i_face_surf = nullptr; mq->diipb = nullptr; mq->diipf = nullptr; mq->djjpf = nullptr;
  CS_FREE(bc.val_f); CS_FREE(bc.flux_diff);
  return 0;
}

/* The same operator WITH flux reconstruction (eqp->ircflu = 1, cs_convection_diffusion.cpp:
   8895-8972) and the least-squares cell gradient it calls (imrgra = 1: CS_GRADIENT_LSQ,
   standard neighbourhood; _lsq_scalar_gradient cs_gradient.cpp:2695-2900, cocg
   _compute_cell_cocg_lsq :2501-2680).  The geometry is whatever the caller passes: i_geo holds
   per interior face II' [3], JJ' [3] and the distance of the cell centres [1]; cell centres are
   those of the system; II' of the boundary faces stays zero (boundary values are then the cell
   values through the BC coefficients a = coefa, b = coefb).  irecmf selects the reconstruction
   formula (cs_glob_velocity_pressure_param->irecmf).  grad_out (3 per cell, may be NULL)
   receives the gradient of pvar computed by cs_gradient_scalar_synced_input with the same
   arguments cs_diffusion_potential uses. */
int
csref_diffusion_potential_rc(void          *sys,
                             const int     *face_cells,
                             int            n_b_faces,
                             const int     *b_face_cells,
                             const double  *b_geo,        /* 7 per boundary face */
                             const double  *i_face_cog,   /* 3 per interior face */
                             const double  *i_geo,        /* 7 per interior face */
                             const double  *diipb,        /* 3 per boundary face, or NULL */
                             int            imrgra,       /* 1: LSQ, 4: Green-Gauss with LSQ
                                                             face values (global default),
                                                             0: iterative Green-Gauss (needs
                                                             the quantities of 4 as well) */
                             const double  *g_geo,        /* imrgra 4: weight [1], OF [3] per
                                                             interior face */
                             const double  *b_face_surf,  /* imrgra 4: per boundary face */
                             const double  *corr_grad_lin, /* imrgra 4: 3x3 per cell (warped
                                                             face correction), or NULL: the
                                                             correction is switched off */
                             int            irecmf,
                             int            init,
                             int            inc,
                             const double  *pvar,
                             const double  *coefa,
                             const double  *coefb,
                             const double  *cofaf,
                             const double  *cofbf,
                             const double  *i_visc,
                             const double  *b_visc,
                             const double  *visel,
                             double        *diverg,
                             double        *grad_out,
                             int            ircflu,       /* eqp->ircflu: 1 = reconstruct */
                             double        *i_massflux,   /* both non-NULL: call */
                             double        *b_massflux,   /* cs_face_diffusion_potential */
                             const double  *frcxt,        /* non-NULL: iphydp = 1 with this
                                                             exterior force, 3 per cell */
                             int            imligr,       /* gradient limiter: -1 none, 0 cell,
                                                             1 face, 2 cell at I'/J' */
                             double         climgr)       /* its factor (< 0: none) */
{
  ref_system *s = static_cast<ref_system *>(sys);
  _set_glob_mesh_faces(s, face_cells);
  cs_mesh_t *mesh = cs_glob_mesh;
  const cs_lnum_t n_cells = s->n_rows, n_i_faces = s->n_faces;

  CS_FREE(mesh->b_face_cells);
  mesh->n_b_faces = n_b_faces;
  mesh->n_b_faces_all = n_b_faces;
  CS_MALLOC(mesh->b_face_cells, (n_b_faces > 0) ? n_b_faces : 1, cs_lnum_t);
  memcpy(mesh->b_face_cells, b_face_cells, sizeof(cs_lnum_t)*n_b_faces);
  if (mesh->b_face_numbering != nullptr)
    cs_numbering_destroy(&(mesh->b_face_numbering));
  mesh->b_face_numbering = cs_numbering_create_default(n_b_faces);
  if (mesh->i_face_numbering != nullptr)
    cs_numbering_destroy(&(mesh->i_face_numbering));
  mesh->i_face_numbering = cs_numbering_create_default(n_i_faces);
  cs_mesh_update_b_cells(mesh);

  if (cs_glob_mesh_quantities == nullptr)
    cs_glob_mesh_quantities = cs_mesh_quantities_create();
  cs_mesh_quantities_t *mq = cs_glob_mesh_quantities;
  cs_mesh_adjacencies_finalize();
  cs_mesh_adjacencies_initialize();
  cs_mesh_adjacencies_update_mesh();
  cs_gradient_free_quantities();

  const cs_lnum_t nb1 = (n_b_faces > 0) ? n_b_faces : 1, ni1 = (n_i_faces > 0) ? n_i_faces : 1;
  std::vector<cs_real_t> b_dist(nb1), i_dist(ni1), b_cog(3*nb1);
  std::vector<cs_nreal_t> b_un(3*nb1);
  std::vector<cs_rreal_t> zero_b(3*nb1, 0.), diipf(3*ni1), djjpf(3*ni1);
  if (diipb != nullptr)
    for (cs_lnum_t k = 0; k < 3*n_b_faces; k++) zero_b[k] = diipb[k];
  for (cs_lnum_t f = 0; f < n_b_faces; f++) {
    for (int k = 0; k < 3; k++) {
      b_un[3*f+k] = b_geo[7*f+k];
      b_cog[3*f+k] = b_geo[7*f+3+k];
    }
    b_dist[f] = b_geo[7*f+6];
  }
  for (cs_lnum_t f = 0; f < n_i_faces; f++) {
    for (int k = 0; k < 3; k++) {
      diipf[3*f+k] = i_geo[7*f+k];
      djjpf[3*f+k] = i_geo[7*f+3+k];
    }
    i_dist[f] = i_geo[7*f+6];
  }
  mq->cell_cen = (cs_real_3_t *)s->cell_cen;
  mq->b_face_u_normal = (cs_nreal_3_t *)b_un.data();
  mq->b_dist = b_dist.data();
  mq->i_dist = i_dist.data();
  mq->i_face_cog = (cs_real_3_t *)const_cast<double *>(i_face_cog);
  mq->b_face_cog = (cs_real_3_t *)b_cog.data();
  mq->i_face_surf = s->face_surf;
  mq->diipb = (cs_rreal_3_t *)zero_b.data();
  mq->diipf = (cs_rreal_3_t *)diipf.data();
  mq->djjpf = (cs_rreal_3_t *)djjpf.data();

  /* quantities of the Green-Gauss reconstruction (_reconstruct_scalar_gradient,
     cs_gradient.cpp:4167): face weights, OF vectors, boundary surfaces; interior normals,
     surfaces and cell volumes are those of the system's mesh association */
  std::vector<cs_real_t> weight(ni1, 0.5), dofij(3*ni1, 0.), b_surf(nb1, 0.);
  int no_disable = 0;
  if (imrgra == 4 || imrgra == 0) {
    for (cs_lnum_t f = 0; f < n_i_faces; f++) {
      weight[f] = g_geo[4*f];
      for (int k = 0; k < 3; k++) dofij[3*f+k] = g_geo[4*f+1+k];
    }
    for (cs_lnum_t f = 0; f < n_b_faces; f++) b_surf[f] = b_face_surf[f];
    mq->weight = weight.data();
    mq->dofij = (cs_real_3_t *)dofij.data();
    mq->b_face_surf = b_surf.data();
    mq->i_face_u_normal = s->face_u_normal;
    mq->cell_vol = s->cell_vol;
    mq->has_disable_flag = 0;
    mq->c_disable_flag = &no_disable;
    mq->corr_grad_lin = (cs_real_33_t *)const_cast<double *>(corr_grad_lin);
  }
  const unsigned mq_flag_prev = cs_glob_mesh_quantities_flag;
  if (corr_grad_lin == nullptr)
    cs_glob_mesh_quantities_flag &= ~((unsigned)CS_BAD_CELLS_WARPED_CORRECTION);
  else
    cs_glob_mesh_quantities_flag |= CS_BAD_CELLS_WARPED_CORRECTION;

  cs_equation_param_t eqp = cs_parameters_equation_param_default();
  eqp.ircflu = ircflu;
  eqp.imrgra = imrgra;     /* 1: least squares, standard neighbourhood; 4: Green-Gauss with
                              least-squares face values */
  eqp.imligr = imligr;     /* gradient limiter (-1: none); with one the values at I' of the
                              boundary faces are clipped too (b_rc_clip_factor = 1) */
  eqp.climgr = climgr;
  if (imligr == CS_GRADIENT_LIMIT_CELL_RC)
    cs_mesh_adjacencies_update_cell_i_faces();
  eqp.iwgrec = 0;
  eqp.verbosity = 0;
  cs_velocity_pressure_param_t *vp = cs_get_glob_velocity_pressure_param();
  const int irecmf_prev = vp->irecmf;
  vp->irecmf = irecmf;

  cs_field_bc_coeffs_t bc;
  memset(&bc, 0, sizeof(bc));
  bc.a = const_cast<cs_real_t *>(coefa); bc.b = const_cast<cs_real_t *>(coefb);
  bc.af = const_cast<cs_real_t *>(cofaf); bc.bf = const_cast<cs_real_t *>(cofbf);

  std::vector<cs_real_t> _pvar(pvar, pvar + n_cells), _visel(visel, visel + n_cells);
  std::vector<cs_real_t> _div(n_cells > 0 ? n_cells : 1, 0.);
  if (!init) memcpy(_div.data(), diverg, sizeof(cs_real_t)*n_cells);

  /* hydrostatic pressure treatment (iphydr = 1, the default of
     cs_glob_velocity_pressure_param): the exterior force enters the gradient and the face
     values */
  std::vector<cs_real_t> frcxt_v;
  if (frcxt != nullptr) frcxt_v.assign(frcxt, frcxt + 3*(size_t)n_cells);
  const int iphydp = (frcxt != nullptr) ? 1 : 0;
  cs_real_3_t *_frcxt = (frcxt != nullptr) ? (cs_real_3_t *)frcxt_v.data() : nullptr;

  if (i_massflux != nullptr && b_massflux != nullptr)
    /* the face-flux sibling (cs_convection_diffusion.cpp:8075-8310): same BC update and
       gradient, fluxes kept per face; always the irecmf = 0 formula */
    cs_face_diffusion_potential(nullptr, &eqp, mesh, mq, init, inc, iphydp, _frcxt, _pvar.data(),
                                &bc, i_visc, b_visc, _visel.data(), i_massflux, b_massflux);
  else {
    cs_diffusion_potential(nullptr, &eqp, mesh, mq, init, inc, iphydp, _frcxt, _pvar.data(), &bc,
                           i_visc, b_visc, _visel.data(), _div.data());
    memcpy(diverg, _div.data(), sizeof(cs_real_t)*n_cells);
  }

  if (grad_out != nullptr && ircflu) {
    /* bc.val_f was filled by cs_diffusion_potential's update of the face values */
    cs_gradient_scalar_synced_input("csref",
                                    (imrgra == 4) ? CS_GRADIENT_GREEN_LSQ
                                    : ((imrgra == 0) ? CS_GRADIENT_GREEN_ITER : CS_GRADIENT_LSQ),
                                    CS_HALO_STANDARD, inc, eqp.nswrgr,
                                    iphydp, 1, 0, (cs_gradient_limit_t)(eqp.imligr), eqp.epsrgr,
                                    eqp.climgr, _frcxt, &bc, _pvar.data(), nullptr,
                                    (cs_real_3_t *)grad_out, nullptr);
  }

  vp->irecmf = irecmf_prev;
  cs_gradient_free_quantities();
  mq->cell_cen = nullptr; mq->b_face_u_normal = nullptr; mq->b_dist = nullptr;
  mq->i_dist = nullptr; mq->i_face_cog = nullptr; mq->b_face_cog = nullptr;
  mq->i_face_surf = nullptr; mq->diipb = nullptr; mq->diipf = nullptr; mq->djjpf = nullptr;
  mq->weight = nullptr; mq->dofij = nullptr; mq->b_face_surf = nullptr;
  mq->i_face_u_normal = nullptr; mq->cell_vol = nullptr; mq->c_disable_flag = nullptr;
  mq->corr_grad_lin = nullptr;
  cs_glob_mesh_quantities_flag = mq_flag_prev;
  CS_FREE(bc.val_f); CS_FREE(bc.flux_diff);
  return 0;
}

/* cs_divergence (src/alge/cs_divergence.cpp:1224-1290): diverg[c] (+)= the signed sum of the
   face mass fluxes of the cell, the reference's real function on the system's faces */
int
csref_divergence(void          *sys,
                 const int     *face_cells,
                 int            n_b_faces,
                 const int     *b_face_cells,
                 int            init,
                 const double  *i_massflux,
                 const double  *b_massflux,
                 double        *diverg)
{
  ref_system *s = static_cast<ref_system *>(sys);
  _set_glob_mesh_faces(s, face_cells);
  cs_mesh_t *mesh = cs_glob_mesh;
  CS_FREE(mesh->b_face_cells);
  mesh->n_b_faces = n_b_faces;
  mesh->n_b_faces_all = n_b_faces;
  CS_MALLOC(mesh->b_face_cells, (n_b_faces > 0) ? n_b_faces : 1, cs_lnum_t);
  memcpy(mesh->b_face_cells, b_face_cells, sizeof(cs_lnum_t)*n_b_faces);
  if (mesh->b_face_numbering != nullptr)
    cs_numbering_destroy(&(mesh->b_face_numbering));
  mesh->b_face_numbering = cs_numbering_create_default(n_b_faces);
  if (mesh->i_face_numbering != nullptr)
    cs_numbering_destroy(&(mesh->i_face_numbering));
  mesh->i_face_numbering = cs_numbering_create_default(s->n_faces);
  cs_divergence(mesh, init, i_massflux, b_massflux, diverg);
  return 0;
}

/* Build the grid hierarchy with the reference's cs_grid_coarsen, following
   the stop criteria of _setup_hierarchy (src/alge/cs_multigrid.cpp:2454-2584). */
void *
csref_hierarchy_build(void   *sys,
                      int     aggregation_limit,
                      int     coarsening_type,
                      int     n_levels_max,
                      int     n_g_rows_min,
                      double  p0p1_relax,
                      int     verbosity)
{
  ref_system *s = static_cast<ref_system *>(sys);
  ref_hierarchy *h = new ref_hierarchy;
  bool q = g_quiet;
  if (verbosity > 0) g_quiet = false;

  cs_grid_t *f
    = cs_grid_create_from_shared(cs_glob_mesh ? cs_glob_mesh->n_i_faces : 0,
                                 1, 1,
                                 cs_glob_mesh ? cs_glob_mesh->i_face_cells
                                              : nullptr,
                                 s->m, false);
  h->grids.push_back(f);
  cs_grid_t *g = f;
  cs_gnum_t n_g_rows = g->n_g_rows, n_g_rows_prev = 0;

  while (true) {
    n_g_rows_prev = n_g_rows;
    if ((int)h->grids.size() >= n_levels_max)
      break;
    cs_grid_t *c = cs_grid_coarsen(g, CS_ALLOC_HOST,
                                   (cs_grid_coarsening_t)coarsening_type,
                                   aggregation_limit, verbosity,
                                   1, 300, 500, p0p1_relax);
    n_g_rows = c->n_g_rows;
    bool add = true;
    if (n_g_rows == 0)
      add = false;
    else if (h->grids.size() > 1) {
      if (   n_g_rows < (cs_gnum_t)n_g_rows_min
          || n_g_rows > (0.8 * n_g_rows_prev))
        add = false;
    }
    if (!add) {
      cs_grid_destroy(&c);
      break;
    }
    h->grids.push_back(c);
    g = c;
  }
  g_quiet = q;
  return h;
}

int
csref_hierarchy_n_levels(void *hier)
{ return (int)static_cast<ref_hierarchy *>(hier)->grids.size(); }

/* info: [0] n_rows, [1] nnz (off-diagonal), [2] n_faces */
void
csref_hierarchy_level_info(void *hier, int level, int *info)
{
  cs_grid_t *g = static_cast<ref_hierarchy *>(hier)->grids[level];
  const cs_lnum_t *ri, *ci; const cs_real_t *d, *x;
  cs_matrix_get_msr_arrays(g->matrix, &ri, &ci, &d, &x);
  info[0] = g->n_rows;
  info[1] = ri[g->n_rows];
  info[2] = g->n_faces;
}

/* Copy a level out. coarse_row (size: parent's n_rows) is the fine->coarse
   map FROM the parent level TO this level (level >= 1), NULL to skip. */
void
csref_hierarchy_level_copy(void    *hier,
                           int      level,
                           int     *coarse_row,
                           int     *row_index,
                           int     *col_id,
                           double  *d_val,
                           double  *x_val,
                           double  *cell_cen,
                           double  *cell_vol)
{
  ref_hierarchy *h = static_cast<ref_hierarchy *>(hier);
  cs_grid_t *g = h->grids[level];
  const cs_lnum_t *ri, *ci; const cs_real_t *d, *x;
  cs_matrix_get_msr_arrays(g->matrix, &ri, &ci, &d, &x);
  const cs_lnum_t n = g->n_rows, nnz = ri[n];
  if (coarse_row != nullptr && level > 0)
    memcpy(coarse_row, g->coarse_row,
           sizeof(cs_lnum_t)*h->grids[level-1]->n_rows);
  if (row_index) memcpy(row_index, ri, sizeof(cs_lnum_t)*(n+1));
  if (col_id) memcpy(col_id, ci, sizeof(cs_lnum_t)*nnz);
  if (d_val) memcpy(d_val, d, sizeof(cs_real_t)*n);
  if (x_val) memcpy(x_val, x, sizeof(cs_real_t)*nnz);
  if (cell_cen && g->cell_cen) memcpy(cell_cen, g->cell_cen, sizeof(cs_real_3_t)*n);
  if (cell_vol && g->cell_vol) memcpy(cell_vol, g->cell_vol, sizeof(cs_real_t)*n);
}

void
csref_hierarchy_destroy(void *hier)
{
  ref_hierarchy *h = static_cast<ref_hierarchy *>(hier);
  for (size_t i = h->grids.size(); i > 0; i--)
    cs_grid_destroy(&(h->grids[i-1]));
  delete h;
}

} /* extern "C" */
