/*============================================================================
 * cs_sles_b200.cpp -- reference-side binding of libcsb200.so (the file a
 * code_saturne maintainer adds as src/alge/cs_sles_b200.cpp; C++, no CUDA).
 *
 * The six cs_sles callbacks (typedefs src/alge/cs_sles.h:91-215) forward to
 * the csb_* entry points of include/csb200.h, which have the same argument
 * lists with a csb_matrix_t in place of the cs_matrix_t.  The only real work
 * here is cs_matrix_t -> csb_matrix_t, done with the reference's public
 * accessors (cs_matrix_get_msr_arrays src/alge/cs_matrix.cpp:6199,
 * cs_matrix_get_mesh_association :6365, cs_matrix_get_halo): the structure is
 * uploaded once per matrix structure, later calls with new coefficients only
 * move the values (cs_matrix_set_coefficients semantics).
 *
 * Compiled and exercised by tests/test_gpu_shim.py through
 * oracle/_ref/libcsref_b200.so (oracle/Makefile.ref), i.e. against the
 * reference's real cs_sles registry and a cs_matrix_t the reference built.
 *============================================================================*/

#include "base/cs_defs.h"

#include <vector>

#include "bft/bft_error.h"
#include "base/cs_halo.h"
#include "mesh/cs_mesh.h"
#include "alge/cs_matrix.h"
#include "alge/cs_sles.h"

#include "csb200.h"
#include "cs_sles_b200.h"

namespace {

struct b200_ctx_t {                       /* cs_sles_t "context" */
  cs_sles_b200_type_t  type;
  int                  n_max_iter;
  csb_multigrid_t     *mg = nullptr;      /* CS_SLES_B200_MULTIGRID */
  csb_sles_it_t       *it = nullptr;      /* the cs_sles_it based types */
  /* device image of the cs_matrix_t */
  csb_matrix_t        *a_dev = nullptr;
  csb_halo_t          *halo_dev = nullptr;
  const cs_matrix_t   *a_host = nullptr;
  const cs_lnum_t     *row_index = nullptr;   /* structure the image was built from */
  const cs_lnum_t     *col_id = nullptr;
  const cs_real_t     *d_val = nullptr, *x_val = nullptr;
  const cs_real_t     *da = nullptr, *xa = nullptr;   /* matrix mapped from native arrays */
  cs_lnum_t            n_rows = 0, n_cols = 0;
  int                  n_uploads = 0, n_value_updates = 0;
};

void
_create_solver(b200_ctx_t *x)
{
  switch (x->type) {
  case CS_SLES_B200_MULTIGRID:
    x->mg = csb_multigrid_create(CSB_MULTIGRID_V_CYCLE);
    if (x->mg != nullptr && x->n_max_iter > 0)       /* cs_multigrid_define defaults, cycles */
      csb_multigrid_set_solver_options(x->mg, CSB_SLES_PCG, CSB_SLES_PCG, CSB_SLES_PCG,
                                       x->n_max_iter, 2, 10, 500, 0, 0, 0, 1., 1., 1.);
    break;
  case CS_SLES_B200_FCG_MG:
    {
      /* cs_sles_it_create(CS_SLES_FCG, -1, ...) + cs_multigrid_pc_create +
         cs_sles_it_transfer_pc, as cs_sles_default.cpp does for the pressure */
      x->it = csb_sles_it_create(CSB_SLES_FCG, -1, x->n_max_iter, 1);
      csb_multigrid_t *pc = csb_multigrid_pc_create(CSB_MULTIGRID_V_CYCLE);
      if (x->it != nullptr && pc != nullptr) csb_sles_it_transfer_pc(x->it, &pc);
    }
    break;
  case CS_SLES_B200_PCG:
    x->it = csb_sles_it_create(CSB_SLES_PCG, 0, x->n_max_iter, 1);
    break;
  case CS_SLES_B200_FCG:
    x->it = csb_sles_it_create(CSB_SLES_FCG, 0, x->n_max_iter, 1);
    break;
  }
  if (x->mg == nullptr && x->it == nullptr)
    bft_error(__FILE__, __LINE__, 0, "cs_sles_b200: %s", csb_last_error());
}

void
_drop_device_matrix(b200_ctx_t *x)
{
  if (x->a_dev != nullptr) csb_matrix_destroy(&x->a_dev);
  if (x->halo_dev != nullptr) csb_halo_destroy(&x->halo_dev);
  x->a_host = nullptr;
  x->row_index = nullptr; x->col_id = nullptr; x->d_val = nullptr; x->x_val = nullptr;
  x->da = nullptr; x->xa = nullptr;
}

/* cs_matrix_t -> csb_matrix_t.  Arrays are borrowed from the matrix and copied to HBM;
   when only the coefficients changed since the last call (same structure arrays), only
   the values move. */
void
_sync_device_matrix(b200_ctx_t *x, const cs_matrix_t *a)
{
  if (cs_matrix_get_type(a) != CS_MATRIX_MSR)
    bft_error(__FILE__, __LINE__, 0,
              "cs_sles_b200: MSR matrices only (got %s).", cs_matrix_get_type_name(a));
  if (cs_matrix_get_diag_block_size(a) != 1 || cs_matrix_get_extra_diag_block_size(a) != 1)
    bft_error(__FILE__, __LINE__, 0, "cs_sles_b200: scalar systems only in this binding.");

  const cs_lnum_t *row_index, *col_id;
  const cs_real_t *d_val, *x_val;
  cs_matrix_get_msr_arrays(a, &row_index, &col_id, &d_val, &x_val);
  const cs_lnum_t n_rows = cs_matrix_get_n_rows(a), n_cols = cs_matrix_get_n_columns(a);
  const int symmetric = cs_matrix_is_symmetric(a) ? 1 : 0;

  /* A matrix mapped from native (face-based) arrays -- cs_matrix_set_coefficients(edges, da,
     xa), cs_matrix_is_mapped_from_native (src/alge/cs_matrix.cpp:5315, :5722): the reference
     builds its multigrid hierarchy from the faces (cs_grid_create_from_shared,
     src/alge/cs_grid.cpp:7485-7491), so does the device matrix made by
     csb_matrix_create_native.  The edges are the mesh's interior faces, as cs_multigrid_setup
     itself takes them (src/alge/cs_multigrid.cpp:5283, :5312). */
  if (cs_matrix_is_mapped_from_native(a) && symmetric && cs_glob_mesh != nullptr
      && cs_matrix_get_halo(a) == nullptr) {
    const cs_real_t *da = cs_matrix_get_diagonal(a), *xa = cs_matrix_get_extra_diagonal(a);
    if (x->a_dev != nullptr && x->row_index == row_index && x->col_id == col_id
        && x->n_rows == n_rows && x->n_cols == n_cols && x->xa != nullptr) {
      if (x->a_host == a && x->da == da && x->xa == xa)
        return;
      if (csb_matrix_set_coefficients_native(x->a_dev, da, xa) != 0)
        bft_error(__FILE__, __LINE__, 0, "cs_sles_b200: %s", csb_last_error());
      x->a_host = a; x->da = da; x->xa = xa;
      x->n_value_updates += 1;
      return;
    }
    _drop_device_matrix(x);
    x->a_dev = csb_matrix_create_native(n_rows, n_cols, cs_glob_mesh->n_i_faces,
                                        (const csb_lnum_t *)cs_glob_mesh->i_face_cells, da, xa, 1);
    if (x->a_dev == nullptr)
      bft_error(__FILE__, __LINE__, 0, "cs_sles_b200: %s", csb_last_error());
    x->n_uploads += 1;
    const cs_real_3_t *cen;
    const cs_nreal_3_t *f_u_normal;
    const cs_real_t *vol, *f_surf;
    cs_matrix_get_mesh_association(a, nullptr, nullptr, nullptr, &cen, &vol, &f_u_normal, &f_surf);
    if (cen != nullptr
        && csb_matrix_set_mesh_association(x->a_dev, cs_glob_mesh->n_i_faces, nullptr, nullptr,
                                           (const cs_real_t *)cen, vol,
                                           (const cs_real_t *)f_u_normal, f_surf) != 0)
      bft_error(__FILE__, __LINE__, 0, "cs_sles_b200: %s", csb_last_error());
    x->a_host = a;
    x->row_index = row_index; x->col_id = col_id; x->da = da; x->xa = xa;
    x->n_rows = n_rows; x->n_cols = n_cols;
    return;
  }

  if (x->a_dev != nullptr && x->row_index == row_index && x->col_id == col_id
      && x->n_rows == n_rows && x->n_cols == n_cols && x->xa == nullptr) {
    if (x->a_host == a && x->d_val == d_val && x->x_val == x_val)
      return;                               /* same object, same arrays: nothing to move */
    if (csb_matrix_set_coefficients(x->a_dev, symmetric, d_val, x_val) != 0)
      bft_error(__FILE__, __LINE__, 0, "cs_sles_b200: %s", csb_last_error());
    x->a_host = a; x->d_val = d_val; x->x_val = x_val;
    x->n_value_updates += 1;
    /* the matrix's own arrays serve the host side of the hierarchy setup (level 0 is
       aggregated from them: no device-to-host copy); they live as long as the coefficients */
    csb_matrix_share_host_arrays(x->a_dev, col_id, d_val, x_val);
    return;
  }

  _drop_device_matrix(x);
  x->a_dev = csb_matrix_create_msr(n_rows, n_cols, row_index, col_id, d_val, x_val, symmetric);
  if (x->a_dev == nullptr)
    bft_error(__FILE__, __LINE__, 0, "cs_sles_b200: %s", csb_last_error());
  x->n_uploads += 1;
  csb_matrix_share_host_arrays(x->a_dev, col_id, d_val, x_val);

  /* geometry for the P0/P1 relaxation of the coarse matrices (cs_grid.cpp:6606-6990) */
  const cs_lnum_t *c2f_idx, *c2f;
  const short int *c2f_sgn;
  const cs_real_3_t *cell_cen;
  const cs_nreal_3_t *face_u_normal;
  const cs_real_t *cell_vol, *face_surf;
  cs_matrix_get_mesh_association(a, &c2f_idx, &c2f, &c2f_sgn, &cell_cen, &cell_vol,
                                 &face_u_normal, &face_surf);
  if (c2f != nullptr && cs_glob_mesh != nullptr) {
    if (csb_matrix_set_mesh_association(x->a_dev, cs_glob_mesh->n_i_faces, c2f, c2f_sgn,
                                        (const cs_real_t *)cell_cen, cell_vol,
                                        (const cs_real_t *)face_u_normal, face_surf) != 0)
      bft_error(__FILE__, __LINE__, 0, "cs_sles_b200: %s", csb_last_error());
  }

  /* cs_halo_t (src/base/cs_halo.h:74-171): standard (face-adjacent) ghosts only */
  const cs_halo_t *h = cs_matrix_get_halo(a);
  if (h != nullptr && h->n_c_domains > 0 && cs_glob_n_ranks > 1) {
    std::vector<cs_lnum_t> si(h->n_c_domains + 1), gi(h->n_c_domains + 1);
    for (int i = 0; i <= h->n_c_domains; i++) {          /* drop the extended-halo half */
      si[i] = h->send_index[2*i];
      gi[i] = h->index[2*i];
    }
    x->halo_dev = csb_halo_create(nullptr, cs_glob_rank_id, h->n_c_domains, h->c_domain_rank,
                                  h->n_local_elts, si.data(), h->send_list, gi.data());
    if (x->halo_dev == nullptr)
      bft_error(__FILE__, __LINE__, 0, "cs_sles_b200: %s", csb_last_error());
    csb_matrix_set_halo(x->a_dev, x->halo_dev);
  }

  x->a_host = a;
  x->row_index = row_index; x->col_id = col_id; x->d_val = d_val; x->x_val = x_val;
  x->n_rows = n_rows; x->n_cols = n_cols;
}

/* cs_sles_setup_t */
void
_setup(void *context, const char *name, const cs_matrix_t *a, int verbosity)
{
  b200_ctx_t *x = static_cast<b200_ctx_t *>(context);
  _sync_device_matrix(x, a);
  if (x->mg != nullptr) {
    csb_multigrid_setup(x->mg, name, x->a_dev, verbosity);
    if (!csb_multigrid_is_setup(x->mg))
      bft_error(__FILE__, __LINE__, 0, "cs_sles_b200 setup [%s]: %s", name, csb_last_error());
  }
  else
    csb_sles_it_setup(x->it, name, x->a_dev, verbosity);
}

/* cs_sles_solve_t: same enum values as cs_sles_convergence_state_t (cs_sles.h:57-68) */
cs_sles_convergence_state_t
_solve(void *context, const char *name, const cs_matrix_t *a, int verbosity, double precision,
       double r_norm, int *n_iter, double *residual, const cs_real_t *rhs, cs_real_t *vx_ini,
       cs_real_t *vx, size_t aux_size, void *aux_vectors)
{
  b200_ctx_t *x = static_cast<b200_ctx_t *>(context);
  const csb_matrix_t *before = x->a_dev;
  _sync_device_matrix(x, a);
  if (before != x->a_dev || (x->mg != nullptr && !csb_multigrid_is_setup(x->mg)))
    _setup(context, name, a, verbosity);
  csb_sles_convergence_state_t cvg = (x->mg != nullptr)
    ? csb_multigrid_solve(x->mg, name, x->a_dev, verbosity, precision, r_norm, n_iter, residual,
                          rhs, vx_ini, vx, aux_size, aux_vectors)
    : csb_sles_it_solve(x->it, name, x->a_dev, verbosity, precision, r_norm, n_iter, residual,
                        rhs, vx_ini, vx, aux_size, aux_vectors);
  if (cvg == CSB_SLES_BREAKDOWN && csb_last_error()[0] != '\0')
    bft_error(__FILE__, __LINE__, 0, "cs_sles_b200 solve [%s]: %s", name, csb_last_error());
  return (cs_sles_convergence_state_t)cvg;
}

/* cs_sles_free_t: drop the hierarchy / preconditioner and the device matrix, keep settings */
void
_free(void *context)
{
  b200_ctx_t *x = static_cast<b200_ctx_t *>(context);
  if (x->mg != nullptr) csb_multigrid_free(x->mg);
  if (x->it != nullptr) csb_sles_it_free(x->it);
  _drop_device_matrix(x);
}

void
_log(const void *context, cs_log_t log_type)
{
  const b200_ctx_t *x = static_cast<const b200_ctx_t *>(context);
  if (x->mg != nullptr) csb_multigrid_log(x->mg, (int)log_type);
  else csb_sles_it_log(x->it, (int)log_type);
}

void *
_copy(const void *context)
{
  const b200_ctx_t *x = static_cast<const b200_ctx_t *>(context);
  b200_ctx_t *n = new b200_ctx_t();
  n->type = x->type; n->n_max_iter = x->n_max_iter;
  if (x->mg != nullptr) n->mg = static_cast<csb_multigrid_t *>(csb_multigrid_copy(x->mg));
  if (x->it != nullptr) n->it = static_cast<csb_sles_it_t *>(csb_sles_it_copy(x->it));
  return n;
}

void
_destroy(void **context)
{
  b200_ctx_t *x = static_cast<b200_ctx_t *>(*context);
  if (x == nullptr) return;
  _free(x);
  if (x->mg != nullptr) { void *p = x->mg; csb_multigrid_destroy(&p); }
  if (x->it != nullptr) { void *p = x->it; csb_sles_it_destroy(&p); }
  delete x;
  *context = nullptr;
}

} // namespace

cs_sles_t *
cs_sles_b200_define(int f_id, const char *name, cs_sles_b200_type_t type, int n_max_iter)
{
  /* one rank <-> one device, as the reference's device path selects it
     (cs_base_cuda_select_default_device, src/base/cs_base_cuda.cu:752) */
  if (csb_get_device() < 0 && csb_init(0) != 0)
    bft_error(__FILE__, __LINE__, 0, "cs_sles_b200: %s", csb_last_error());
  b200_ctx_t *x = new b200_ctx_t();
  x->type = type;
  x->n_max_iter = n_max_iter;
  _create_solver(x);
  return cs_sles_define(f_id, name, x, "cs_sles_b200_t", _setup, _solve, _free, _log, _copy,
                        _destroy);
}

void
cs_sles_b200_get_transfer_counts(const void *context, int counts[2])
{
  const b200_ctx_t *x = static_cast<const b200_ctx_t *>(context);
  counts[0] = x->n_uploads;
  counts[1] = x->n_value_updates;
}
