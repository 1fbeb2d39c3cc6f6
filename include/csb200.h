/*============================================================================
 * csb200.h -- C-ABI of the B200-native linear-solve path for code_saturne.
 *
 * Drop-in boundary for the reference's src/alge hot path (SpMV -> Krylov ->
 * multigrid V-cycle -> halo exchange).  Every entry point below replaces the
 * reference interface cited next to it; argument meaning, return values and
 * error behaviour follow that interface.  Plain pointers and sizes only.
 *
 * Conventions (src/base/cs_defs.h:317-335): cs_lnum_t = int32, cs_real_t =
 * double.  Pointers named *_h must be host pointers; rhs/vx/x/y arguments may
 * be host OR device pointers (detected with cudaPointerGetAttributes, like
 * cs_check_device_ptr in src/alge/cs_sles_it.cpp:5300-5332): host arrays are
 * staged to the device and copied back inside the call.
 *
 * Error convention: functions returning int give 0 on success, < 0 on error
 * (csb_last_error() returns the message); solve functions return the
 * reference's cs_sles_convergence_state_t values.  There is NO CPU fallback:
 * without a usable CUDA device csb_init() fails and nothing else works.
 *============================================================================*/

#ifndef CSB200_H
#define CSB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int    csb_lnum_t;   /* cs_lnum_t */
typedef double csb_real_t;   /* cs_real_t */
typedef unsigned long long csb_gnum_t;   /* cs_gnum_t (64-bit global ids) */

/* cs_sles_convergence_state_t, src/alge/cs_sles.h:57-68 */
typedef enum {
  CSB_SLES_DIVERGED      = -3,
  CSB_SLES_BREAKDOWN     = -2,
  CSB_SLES_MAX_ITERATION = -1,
  CSB_SLES_ITERATING     =  0,
  CSB_SLES_CONVERGED     =  1
} csb_sles_convergence_state_t;

/* cs_sles_it_type_t, src/alge/cs_sles_it.h:55-77 (same numeric values) */
typedef enum {
  CSB_SLES_PCG      = 0,
  CSB_SLES_FCG      = 1,   /* flexible CG, cs_sles_it.cpp:832-985 */
  CSB_SLES_IPCG     = 2,   /* inexact-preconditioning CG, cs_sles_it.cpp:1016-1203 */
  CSB_SLES_JACOBI   = 3,
  CSB_SLES_BICGSTAB = 4,
  CSB_SLES_GCR      = 6,   /* preconditioned GCR with restart, cs_sles_it.cpp:3272-3480; also the
                              fallback solver on a device (:5383-5385) */
  /* multigrid smoothers without convergence test (cs_sles_it.h:79-83,
     src/alge/cs_multigrid_smoother.cpp); created with csb_multigrid_smoother_create or
     selected through csb_multigrid_set_solver_options with precision_mult < 0 */
  /* process-local Gauss-Seidel family (cs_multigrid_smoother.cpp:1053-1630), sequential
     (one-thread) semantics reproduced exactly with level-scheduled sweeps: rows of one
     dependency level are updated together, levels in order */
  CSB_SLES_P_GAUSS_SEIDEL     = 8,   /* _p_gauss_seidel_msr      :1053 forward sweeps */
  CSB_SLES_P_SYM_GAUSS_SEIDEL = 9,   /* _p_sym_gauss_seidel_msr  :1179 forward + backward */
  CSB_SLES_TS_F_GAUSS_SEIDEL  = 17,  /* _ts_f_gauss_seidel_msr   :1389 one lower solve */
  CSB_SLES_TS_B_GAUSS_SEIDEL  = 18,  /* _ts_b_gauss_seidel_msr   :1519 one upper solve */
  CSB_SLES_L1_JACOBI = 13,  /* _l1_jacobi       :804-905 */
  CSB_SLES_R_JACOBI  = 14,  /* _relaxed_jacobi  :244-373, w = 2/3 */
  CSB_SLES_RJ2       = 15,  /* scheduled relaxation {1.7319, 0.5695} */
  CSB_SLES_RJ3       = 16   /* scheduled relaxation {2.2473, 0.8573, 0.5296} */
} csb_sles_it_type_t;

/* cs_multigrid_type_t, src/alge/cs_multigrid.h:58-65 */
typedef enum {
  CSB_MULTIGRID_V_CYCLE = 0,
  CSB_MULTIGRID_K_CYCLE = 1    /* as a preconditioner (one cycle): csb_multigrid_pc_create */
} csb_multigrid_type_t;

typedef struct _csb_matrix_t     csb_matrix_t;      /* cs_matrix_t (MSR) */
typedef struct _csb_halo_t       csb_halo_t;        /* cs_halo_t */
typedef struct _csb_sles_it_t    csb_sles_it_t;     /* cs_sles_it_t */
typedef struct _csb_multigrid_t  csb_multigrid_t;   /* cs_multigrid_t */
typedef struct _csb_pcomm_t      csb_pcomm_t;       /* peer-memory communicator */
typedef struct _csb_pcomm_group_t csb_pcomm_group_t; /* ranks as threads of one process */

/*----------------------------------------------------------------------------
 * Runtime (replaces cs_base_cuda_select_default_device / stream table,
 * src/base/cs_base_cuda.cu:732-815)
 *----------------------------------------------------------------------------*/

/* Select CUDA device (>= 0) and create the library stream. < 0 on failure. */
int          csb_init(int device_id);
/* Device selected by csb_init in this thread, -1 before it (cs_base_cuda_get_device,
   src/base/cs_base_cuda.cu:700). */
int          csb_get_device(void);
void         csb_finalize(void);
const char  *csb_last_error(void);
const char  *csb_version(void);
/* name of the kernel csb_matrix_vector_multiply launches for scalar MSR matrices (k_spmv_rows
   by default; CSB_SPMV_KERNEL=pipeline|stream selects the others) */
const char  *csb_spmv_kernel_name(void);

/* Use an externally owned cudaStream_t (passed as void *) for all work;
   NULL restores the library's own stream
   (cs_matrix_spmv_cuda_set_stream, src/alge/cs_matrix_spmv_cuda.cu). */
int          csb_set_stream(void *cuda_stream);
int          csb_stream_synchronize(void);

/* Device memory helpers (CS_MALLOC_HD / cs_copy_h2d / cs_copy_d2h,
   src/base/cs_mem.h:48-55) */
void        *csb_malloc_device(size_t n_bytes);
void         csb_free_device(void *p);
int          csb_copy_h2d(void *dst_d, const void *src_h, size_t n_bytes);
int          csb_copy_d2h(void *dst_h, const void *src_d, size_t n_bytes);
void        *csb_malloc_pinned(size_t n_bytes);
void         csb_free_pinned(void *p);

/* Number of kernel launches issued by this library since the last reset
   (bench.py "gpu_launches"). */
long long    csb_launch_count(int reset);

/*----------------------------------------------------------------------------
 * Matrix (cs_matrix_t, MSR)
 *----------------------------------------------------------------------------*/

/* cs_matrix_structure_create_msr_shared + cs_matrix_create +
   cs_matrix_transfer_coefficients_msr  (src/alge/cs_matrix.cpp:4518, :4682,
   :5451).  Arrays may be host or device pointers; they are copied to HBM.
   n_cols_ext = n_rows + number of ghost columns (halo). */
csb_matrix_t *csb_matrix_create_msr(csb_lnum_t         n_rows,
                                    csb_lnum_t         n_cols_ext,
                                    const csb_lnum_t  *row_index,
                                    const csb_lnum_t  *col_id,
                                    const csb_real_t  *d_val,
                                    const csb_real_t  *x_val,
                                    int                symmetric);

/* Lend the host copies of the matrix arrays (same content as given to csb_matrix_create_msr).
   They must stay valid and unchanged until the matrix is destroyed, its values are replaced
   (csb_matrix_set_coefficients / _compute_coeffs drop the lent values) or this is called
   again with NULL.  The reference's cs_matrix_t shares the caller's arrays the same way
   (src/alge/cs_matrix.cpp:5451); here they spare csb_multigrid_setup the device-to-host copy
   of level 0 for the host-side aggregation.  Returns 0. */
int csb_matrix_share_host_arrays(csb_matrix_t *m, const csb_lnum_t *col_id,
                                 const csb_real_t *d_val, const csb_real_t *x_val);
/* Give the cached page-locked blocks of the host pool (hierarchy setup) back to the system */
void csb_host_pool_trim(void);

/* MSR with db_size x db_size diagonal blocks and scalar extra-diagonal entries (fill type
   CS_MATRIX_BLOCK_D: cs_matrix_transfer_coefficients_msr with db_size = 3, eb_size = 1; the
   velocity systems).  d_val: n_rows*db_size*db_size, row-major blocks; vectors hold db_size
   values per row.  SpMV = _b_mat_vec_p_l_msr_3 (src/alge/cs_matrix_spmv.cpp:2394-2455);
   solver on such matrices: CSB_SLES_BICGSTAB with the point-Jacobi preconditioner of
   cs_sles_pc.cpp:383-388.  db_size 1 and 3. */
csb_matrix_t *csb_matrix_create_msr_block(csb_lnum_t         n_rows,
                                          csb_lnum_t         n_cols_ext,
                                          int                db_size,
                                          const csb_lnum_t  *row_index,
                                          const csb_lnum_t  *col_id,
                                          const csb_real_t  *d_val,
                                          const csb_real_t  *x_val,
                                          int                symmetric);
int  csb_matrix_get_diag_block_size(const csb_matrix_t *m);

/* CS_MATRIX_CSR (cs_matrix_struct_csr_t, src/alge/cs_matrix_priv.h:143-160): the
   diagonal is an ordinary entry of (row_index, col_id, val).  csb_matrix_vector_multiply
   on such a matrix sums every row in CSR order like _mat_vec_p_l_csr
   (src/alge/cs_matrix_spmv.cpp:1890-1955, exclude_diag skips col_id == row); the solvers
   work on the MSR split made at creation (host arrays only). */
csb_matrix_t *csb_matrix_create_csr(csb_lnum_t         n_rows,
                                    csb_lnum_t         n_cols_ext,
                                    const csb_lnum_t  *row_index,
                                    const csb_lnum_t  *col_id,
                                    const csb_real_t  *val,
                                    int                symmetric);

/* Matrix "mapped from native arrays": what a caller that holds face-based coefficients does
   in the reference -- cs_matrix_structure_create(CS_MATRIX_MSR, n_rows, n_cols_ext, n_edges,
   edges, ...) + cs_matrix_set_coefficients(m, symmetric, 1, 1, n_edges, edges, da, xa)
   (src/alge/cs_matrix.cpp:4328, :5315).  edges[n_edges][2] (host), da[n_rows], xa[n_edges]
   (symmetric: one value per edge; host or device).  Two edges may not join the same pair of
   cells.  The result is an MSR matrix like the others (sorted columns) that also keeps the
   edges and their values: cs_matrix_is_mapped_from_native (cs_matrix.cpp:5722) is true for
   it in the reference, and csb_multigrid_setup then builds the hierarchy the way the
   reference does for such a matrix (cs_grid_create_from_shared, src/alge/cs_grid.cpp:7485-7491):
   face-list aggregation _automatic_aggregation_fc (:3997-4263) when one aggregation block is
   used, coarse faces numbered by _coarsen_face_cell (:637-788), coarse values summed in face
   order by _compute_coarse_cell_quantities / _compute_coarse_quantities_native (:4778,
   :5144-5558) and turned into MSR by _matrix_from_native (:7105).  For the geometry call
   csb_matrix_set_mesh_association with n_faces = n_edges and NULL c2f / c2f_sgn. */
csb_matrix_t *csb_matrix_create_native(csb_lnum_t         n_rows,
                                       csb_lnum_t         n_cols_ext,
                                       csb_lnum_t         n_edges,
                                       const csb_lnum_t  *edges,
                                       const csb_real_t  *da,
                                       const csb_real_t  *xa,
                                       int                symmetric);
/* new values on the same edges */
int  csb_matrix_set_coefficients_native(csb_matrix_t *m, const csb_real_t *da,
                                        const csb_real_t *xa);

/*----------------------------------------------------------------------------
 * Matrix assembler (cs_matrix_assembler_t, src/alge/cs_matrix_assembler.cpp): matrices whose
 * graph is a list of (row, column) global-id pairs rather than the mesh faces -- CDO systems,
 * internal coupling.  One rank: every id inside l_range.
 *
 *   ma = csb_matrix_assembler_create(l_range, 1);            cs_matrix_assembler_create :2366
 *   csb_matrix_assembler_add_g_ids(ma, n, rows, cols); ...   _add_g_ids :2673
 *   csb_matrix_assembler_compute(ma);                        _compute :2734
 *   m = csb_matrix_create_from_assembler(ma);                cs_matrix_structure_create_from_assembler
 *   mav = csb_matrix_assembler_values_init(m, ma, 1, 1);     cs_matrix.cpp:5635
 *   csb_matrix_assembler_values_add_g(mav, n, rows, cols, vals); ...   :3276
 *   csb_matrix_assembler_values_finalize(&mav);              :3119
 *
 * Every entry is the sum of its coefficients in arrival order, starting from zero: the
 * reference's one-thread result bit for bit (_msr_assembler_values_add, cs_matrix.cpp:2422).
 * The matrix is non-symmetric as far as its flag goes (_set_fill_info(matrix, false, ...)).
 *----------------------------------------------------------------------------*/
typedef struct _csb_matrix_assembler_t        csb_matrix_assembler_t;
typedef struct _csb_matrix_assembler_values_t csb_matrix_assembler_values_t;

csb_matrix_assembler_t *csb_matrix_assembler_create(const csb_gnum_t l_range[2],
                                                    int separate_diag /* must be 1: MSR */);
void        csb_matrix_assembler_destroy(csb_matrix_assembler_t **ma);
int         csb_matrix_assembler_add_g_ids(csb_matrix_assembler_t *ma, csb_lnum_t n,
                                           const csb_gnum_t *row_g_id,
                                           const csb_gnum_t *col_g_id);
int         csb_matrix_assembler_compute(csb_matrix_assembler_t *ma);
csb_lnum_t  csb_matrix_assembler_get_n_rows(const csb_matrix_assembler_t *ma);
/* cs_matrix_assembler_get_row_index / _get_col_ids (:2900, :2926) as copies; returns the
   number of extra-diagonal entries (-1 before _compute); either pointer may be NULL */
csb_lnum_t  csb_matrix_assembler_get_structure(const csb_matrix_assembler_t *ma,
                                               csb_lnum_t *row_index, csb_lnum_t *col_id);
csb_matrix_t *csb_matrix_create_from_assembler(const csb_matrix_assembler_t *ma);
csb_matrix_assembler_values_t *
csb_matrix_assembler_values_init(csb_matrix_t *m, const csb_matrix_assembler_t *ma,
                                 int diag_block_size, int extra_diag_block_size);
int  csb_matrix_assembler_values_add_g(csb_matrix_assembler_values_t *mav, csb_lnum_t n,
                                       const csb_gnum_t *row_g_id, const csb_gnum_t *col_g_id,
                                       const csb_real_t *val);
/* cs_matrix_assembler_values_done (:3510): fold what was added so far into the matrix */
int  csb_matrix_assembler_values_done(csb_matrix_assembler_values_t *mav);
/* Coefficients of ALL the pairs declared with csb_matrix_assembler_add_g_ids, in declaration
   order (val[k] for the k-th pair; host or device array): the re-assembly of a time loop that
   adds the same ids at every step, as ONE kernel over a plan kept in HBM. */
int  csb_matrix_assembler_values_add_declared(csb_matrix_assembler_values_t *mav,
                                              const csb_real_t *val);
int  csb_matrix_assembler_values_finalize(csb_matrix_assembler_values_t **mav);

/* Replace coefficient values only (cs_matrix_set_coefficients on an existing
   structure, src/alge/cs_matrix.cpp:5315). */
int  csb_matrix_set_coefficients(csb_matrix_t      *m,
                                 int                symmetric,
                                 const csb_real_t  *d_val,
                                 const csb_real_t  *x_val);

/* cs_matrix_set_mesh_association (src/alge/cs_matrix.cpp:6328): cell -> face
   adjacency aligned with the MSR entries (c2f_idx == row_index), geometry
   used by the P0/P1-relaxed coarse matrices.  Host or device pointers. */
int  csb_matrix_set_mesh_association(csb_matrix_t      *m,
                                     csb_lnum_t         n_faces,
                                     const csb_lnum_t  *c2f,
                                     const short       *c2f_sgn,
                                     const csb_real_t  *cell_cen,      /* [n_cols_ext][3] */
                                     const csb_real_t  *cell_vol,      /* [n_cols_ext] */
                                     const csb_real_t  *face_u_normal, /* [n_faces][3] */
                                     const csb_real_t  *face_surf);    /* [n_faces] */

/* Attach a halo (cs_matrix_structure_create_msr_shared "halo" argument). */
/* Boundary faces by adjacent cell (cs_mesh_t::b_face_cells, n_b_faces entries): builds the
   cell -> boundary faces adjacency of cs_mesh_adjacencies (src/mesh/cs_mesh_adjacencies.cpp:
   705-760) used by csb_matrix_compute_coeffs. */
int  csb_matrix_set_boundary_faces(csb_matrix_t *m, csb_lnum_t n_b_faces,
                                   const csb_lnum_t *b_face_cells);

/* cs_matrix_compute_coeffs (src/alge/cs_matrix_building.cpp:1948-2283, same argument order
   without the field pointer; bc_coeffs->b = coefbp, ->bf = cofbfp) for scalar variables: the
   MSR coefficients are computed in HBM from per-face / per-cell arrays, which may be host or
   device pointers:
     iconvp = 0  symmetric diffusion matrix, _sym_coeffs_msr (:136-206); needs rovsdt[n_rows],
                 i_visc[n_faces], b_visc / cofbfp[n_b_faces]
     iconvp = 1  convection-diffusion, _coeffs_msr (:407-520): also i_massflux[n_faces],
                 b_massflux / coefbp[n_b_faces], and xcpp[n_cols] when imucpp != 0
   relaxp < 1 scales the diagonal by 1/relaxp, ndircp <= 0 applies the (1 + 1e-7) shift
   (:2243-2262).  Needs the mesh association (face and face sign of every MSR entry) and, for
   boundary terms, csb_matrix_set_boundary_faces.  Bit-identical to the reference's CPU
   result: same per-cell summation order. */
int  csb_matrix_compute_coeffs(csb_matrix_t *m, int iconvp, int idiffp, int ndircp,
                               double thetap, double relaxp, int imucpp,
                               const csb_real_t *coefbp, const csb_real_t *cofbfp,
                               const csb_real_t *rovsdt, const csb_real_t *i_massflux,
                               const csb_real_t *b_massflux, const csb_real_t *i_visc,
                               const csb_real_t *b_visc, const csb_real_t *xcpp);

/* cs_matrix_compute_coeffs<3> (src/alge/cs_matrix_building.cpp:2284-2600) for a vector
   variable with scalar diffusion (tensorial_diffusion 1, eb_size 1), on a matrix created with
   csb_matrix_create_msr_block (db 3): _sym_coeffs_msr<3> (:250-353) when iconvp = 0,
   _coeffs_msr<3> (:560-692) when iconvp = 1.  fimp[n_rows x 3 x 3] may be NULL; coefbp and
   cofbfp hold a 3x3 block per boundary face.  Same conventions as csb_matrix_compute_coeffs. */
int  csb_matrix_compute_coeffs_vector(csb_matrix_t *m, int iconvp, int idiffp, int ndircp,
                                      double thetap, double relaxp, const csb_real_t *coefbp,
                                      const csb_real_t *cofbfp, const csb_real_t *fimp,
                                      const csb_real_t *i_massflux,
                                      const csb_real_t *b_massflux, const csb_real_t *i_visc,
                                      const csb_real_t *b_visc);

/* cs_diffusion_potential (src/alge/cs_convection_diffusion.cpp:8743-8990; the explicit
   right-hand-side operator of the pressure correction) without
   flux reconstruction (eqp->ircflu = 0; on an orthogonal mesh the reconstructed branch adds
   exact zeros): init >= 1 zeroes diverg first, then
     diverg[c0] += i_visc[f] (pvar[c0] - pvar[c1]),  diverg[c1] -= the same   (interior faces)
     diverg[c]  += b_visc[f] (inc cofaf[f] + cofbf[f] pvar[c])                 (boundary faces)
   summed per cell in the reference's order (ascending interior face ids, then ascending
   boundary face ids): bit-identical to its one-thread result.  Uses the mesh association of m
   (face of every entry) and csb_matrix_set_boundary_faces; host or device pointers. */
int  csb_diffusion_potential(const csb_matrix_t *m, int init, int inc, const csb_real_t *pvar,
                             const csb_real_t *cofaf, const csb_real_t *cofbf,
                             const csb_real_t *i_visc, const csb_real_t *b_visc,
                             csb_real_t *diverg);

/* Reconstruction on non-orthogonal meshes.
   csb_matrix_set_face_geometry: the face quantities cs_mesh_quantities_t holds
   (src/mesh/cs_mesh_quantities.h:101-125): i_dist[n_faces], diipf / djjpf[n_faces][3] (II',
   JJ'), b_face_u_normal[n_b_faces][3], b_dist[n_b_faces].  After
   csb_matrix_set_mesh_association and csb_matrix_set_boundary_faces.
   csb_matrix_set_boundary_iprime: II' of the boundary faces (cs_mesh_quantities_t::diipb,
   [n_b_faces][3]; zero until given, NULL resets it; to be called after
   csb_matrix_set_face_geometry).  With it the value of the variable at I' of every boundary
   face is reconstructed like cs_gradient_boundary_iprime_lsq_s
   (src/alge/cs_gradient_boundary.cpp:197-640; its limiter: csb_matrix_set_gradient_limiter), face values are
   inc*coefa + coefb*val_ip and boundary fluxes inc*cofaf + cofbf*val_ip
   (cs_boundary_conditions_update_bc_coeff_face_values,
   src/base/cs_boundary_conditions_set_coeffs.cpp:125-195), and the boundary directions of the
   least-squares system become n + II'/b_dist (cs_gradient.cpp:2632-2636, :2880-2885).
   csb_gradient_scalar_lsq: cs_gradient_scalar_synced_input with CS_GRADIENT_LSQ (imrgra = 1),
   standard neighbourhood, no weighting (src/alge/cs_gradient.cpp:9651; limiters: see
   csb_matrix_set_gradient_limiter;
   _lsq_scalar_gradient :2695-2900, cocg _compute_cell_cocg_lsq :2501-2680 -- kept in single
   precision and inverted by the float Gauss-Jordan elimination like the reference's
   cs_cocg_t); grad[n_rows][3]; boundary face value inc*coefa + coefb*(pvar at I').
   csb_diffusion_potential_rc: cs_diffusion_potential with eqp->ircflu = 1 (:8895-8982), irecmf
   as cs_glob_velocity_pressure_param->irecmf (0 or 1).
   All bit-identical to the reference's one-thread results; host or device pointers. */
int  csb_matrix_set_face_geometry(csb_matrix_t *m, const csb_real_t *i_dist,
                                  const csb_real_t *diipf, const csb_real_t *djjpf,
                                  const csb_real_t *b_face_u_normal, const csb_real_t *b_dist);
int  csb_matrix_set_boundary_iprime(csb_matrix_t *m, const csb_real_t *diipb);
/* Green-Gauss gradient with least-squares face values (imrgra = 4, CS_GRADIENT_GREEN_LSQ: the
   default of cs_glob_space_disc, src/base/cs_parameters.cpp:470): the least-squares gradient,
   then _reconstruct_scalar_gradient (src/alge/cs_gradient.cpp:7697-7797, :4167-4540; no
   hydrostatic term, no cell weighting, no disabled cell).
   csb_matrix_set_green_geometry: weight[n_faces], dofij[n_faces][3] (OF), b_face_surf[n_b_faces]
   of cs_mesh_quantities_t and, optionally, corr_grad_lin[n_rows][3][3] (warped-face correction,
   CS_BAD_CELLS_WARPED_CORRECTION: on by default in the reference; NULL = off); after
   csb_matrix_set_face_geometry.
   csb_matrix_set_gradient_type: eqp->imrgra for csb_diffusion_potential_rc (1, the initial
   value, or 4).  csb_gradient_scalar: the gradient alone for imrgra 1 or 4. */
int  csb_matrix_set_green_geometry(csb_matrix_t *m, const csb_real_t *weight,
                                   const csb_real_t *dofij, const csb_real_t *b_face_surf,
                                   const csb_real_t *corr_grad_lin);
int  csb_matrix_set_gradient_type(csb_matrix_t *m, int imrgra);
/* imrgra = 0, the iterative Green-Gauss gradient (CS_GRADIENT_GREEN_ITER:
   _initialize_scalar_gradient src/alge/cs_gradient.cpp:1329-1590, _iterative_scalar_gradient
   :1891-2300, cocg_it :1776-1860; without the hydrostatic treatment), is accepted by
   csb_matrix_set_gradient_type and csb_gradient_scalar too; it needs the quantities of
   csb_matrix_set_green_geometry.  csb_matrix_set_gradient_sweeps: eqp->nswrgr, eqp->epsrgr
   (defaults 100, 1e-4); csb_matrix_get_gradient_sweeps: sweeps of the last call (the
   reference's n_sweeps; one reduction per sweep is read by the host for the test). */
/* csb_matrix_set_gradient_limiter: eqp->imligr (-1 none, the default; 0 / 1 / 2 =
   CS_GRADIENT_LIMIT_CELL / _FACE / _CELL_RC: _scalar_gradient_clipping, :808-1215, standard
   neighbourhood), eqp->climgr and eqp->b_rc_clip_factor (limiter of the values at I' of the
   boundary faces, src/alge/cs_gradient_boundary.cpp:612-628, active when imligr >= 0;
   reference default 1).  Applies to every gradient type and to the operators built on them. */
int  csb_matrix_set_gradient_limiter(csb_matrix_t *m, int imligr, double climgr,
                                     double b_rc_clip_factor);
int  csb_matrix_set_gradient_sweeps(csb_matrix_t *m, int nswrgr, double epsrgr);
int  csb_matrix_get_gradient_sweeps(const csb_matrix_t *m);
int  csb_gradient_scalar(const csb_matrix_t *m, int imrgra, int inc, const csb_real_t *pvar,
                         const csb_real_t *coefa, const csb_real_t *coefb, csb_real_t *grad);
int  csb_gradient_scalar_lsq(const csb_matrix_t *m, int inc, const csb_real_t *pvar,
                             const csb_real_t *coefa, const csb_real_t *coefb, csb_real_t *grad);
int  csb_diffusion_potential_rc(const csb_matrix_t *m, int irecmf, int init, int inc,
                                const csb_real_t *pvar, const csb_real_t *coefa,
                                const csb_real_t *coefb, const csb_real_t *cofaf,
                                const csb_real_t *cofbf, const csb_real_t *i_visc,
                                const csb_real_t *b_visc, const csb_real_t *visel,
                                csb_real_t *diverg);

/* The other two face operators of the pressure correction.
   csb_divergence: cs_divergence (src/alge/cs_divergence.cpp:1224-1290), diverg[n_rows] (+)= the
   signed sum of the face mass fluxes i_massflux[n_faces], b_massflux[n_b_faces] of every cell,
   in the reference's order (ascending interior faces, then boundary faces).
   csb_face_diffusion_potential: cs_face_diffusion_potential
   (src/alge/cs_convection_diffusion.cpp:8075-8310), the face fluxes cs_diffusion_potential
   sums, kept per face: ircflu = 0 without reconstruction (coefa, coefb, visel may be NULL),
   ircflu = 1 with it (gradient of csb_matrix_set_gradient_type, values at I' of the boundary
   faces); init >= 1 starts from zero, init = 0 adds to the arrays given.
   Both bit-identical to the reference; host or device pointers; csb_divergence of the fluxes
   of csb_face_diffusion_potential equals csb_diffusion_potential[_rc] (irecmf 0). */
int  csb_divergence(const csb_matrix_t *m, int init, const csb_real_t *i_massflux,
                    const csb_real_t *b_massflux, csb_real_t *diverg);
int  csb_face_diffusion_potential(const csb_matrix_t *m, int ircflu, int init, int inc,
                                  const csb_real_t *pvar, const csb_real_t *coefa,
                                  const csb_real_t *coefb, const csb_real_t *cofaf,
                                  const csb_real_t *cofbf, const csb_real_t *i_visc,
                                  const csb_real_t *b_visc, const csb_real_t *visel,
                                  csb_real_t *i_massflux, csb_real_t *b_massflux);

/* Hydrostatic pressure treatment (hyd_p_flag / iphydp = 1 with the exterior force f_ext / frcxt
   [n_rows][3]; cs_glob_velocity_pressure_param->iphydr = 1 is the reference's default and
   cs_pressure_correction hands it to these operators, src/base/cs_pressure_correction.cpp:
   2113-2127).  The *_hyd entry points carry the two arguments in the reference's positions;
   the plain ones are the same calls with the flag off.  What changes: the least-squares
   gradient (_lsq_scalar_gradient_hyd_p, src/alge/cs_gradient.cpp:3230-3560), the Green-Gauss
   pass (:4314-4424), the values at I' and the face values of the boundary faces
   (src/alge/cs_gradient_boundary.cpp:285-640, src/base/cs_boundary_conditions_set_coeffs.cpp:
   176-192).  Needs the face centres: csb_matrix_set_face_cog (i_face_cog[n_faces][3],
   b_face_cog[n_b_faces][3] of cs_mesh_quantities_t).  Flag values 0 and 2 switch it off, as in
   the reference.  No porosity terms.  Bit-identical to the reference's one-thread results. */
int  csb_matrix_set_face_cog(csb_matrix_t *m, const csb_real_t *i_face_cog,
                             const csb_real_t *b_face_cog);
int  csb_gradient_scalar_hyd(const csb_matrix_t *m, int imrgra, int inc, int hyd_p_flag,
                             const csb_real_t *f_ext, const csb_real_t *pvar,
                             const csb_real_t *coefa, const csb_real_t *coefb, csb_real_t *grad);
int  csb_diffusion_potential_hyd(const csb_matrix_t *m, int irecmf, int init, int inc,
                                 int iphydp, const csb_real_t *frcxt, const csb_real_t *pvar,
                                 const csb_real_t *coefa, const csb_real_t *coefb,
                                 const csb_real_t *cofaf, const csb_real_t *cofbf,
                                 const csb_real_t *i_visc, const csb_real_t *b_visc,
                                 const csb_real_t *visel, csb_real_t *diverg);
int  csb_face_diffusion_potential_hyd(const csb_matrix_t *m, int ircflu, int init, int inc,
                                      int iphydp, const csb_real_t *frcxt,
                                      const csb_real_t *pvar, const csb_real_t *coefa,
                                      const csb_real_t *coefb, const csb_real_t *cofaf,
                                      const csb_real_t *cofbf, const csb_real_t *i_visc,
                                      const csb_real_t *b_visc, const csb_real_t *visel,
                                      csb_real_t *i_massflux, csb_real_t *b_massflux);

/* Copy of the MSR coefficients held in HBM (cs_matrix_get_msr_arrays, cs_matrix.cpp:6199,
   lends the host arrays; these live on the device).  Either pointer may be NULL. */
int  csb_matrix_get_coefficients(const csb_matrix_t *m, csb_real_t *d_val, csb_real_t *x_val);

int  csb_matrix_set_halo(csb_matrix_t *m, csb_halo_t *halo);

void csb_matrix_destroy(csb_matrix_t **m);

csb_lnum_t csb_matrix_get_n_rows(const csb_matrix_t *m);     /* cs_matrix.cpp:4976 */
csb_lnum_t csb_matrix_get_n_columns(const csb_matrix_t *m);  /* cs_matrix.cpp:4957 */
csb_lnum_t csb_matrix_get_n_entries(const csb_matrix_t *m);  /* cs_matrix.cpp:4998 */

/* y = A.x  -- cs_matrix_vector_multiply (src/alge/cs_matrix.cpp:6405);
   x has n_cols_ext entries (ghosts are updated through the halo first),
   y has n_rows.  Host or device pointers. */
int  csb_matrix_vector_multiply(const csb_matrix_t *m, csb_real_t *x, csb_real_t *y);

/* y = (A - D).x -- cs_matrix_vector_multiply_partial(CS_MATRIX_SPMV_E)
   (src/alge/cs_matrix.cpp:6593) */
int  csb_matrix_vector_multiply_exdiag(const csb_matrix_t *m, csb_real_t *x, csb_real_t *y);

/* Algorithmic HBM bytes of one SpMV launch: 12*nnz + 4*(n+1) + 8n + 8n_cols + 8n
   (SURVEY.md section 8(d)). */
double csb_matrix_spmv_bytes(const csb_matrix_t *m);

/*----------------------------------------------------------------------------
 * BLAS-1 reductions (cs_dot, cs_dot_xx, cs_dot_xx_xy_yz;
 * src/alge/cs_blas.cpp:1629-1905); device-resident 2-stage deterministic sums.
 *----------------------------------------------------------------------------*/

double csb_dot(csb_lnum_t n, const csb_real_t *x, const csb_real_t *y);
double csb_dot_xx(csb_lnum_t n, const csb_real_t *x);
int    csb_dot_xx_xy_yz(csb_lnum_t n, const csb_real_t *x, const csb_real_t *y,
                        const csb_real_t *z, double *xx, double *xy, double *yz);

/*----------------------------------------------------------------------------
 * Halo (cs_halo_t, src/base/cs_halo.h:74-171): ghost cells exchanged with
 * neighbour ranks.  One process per GPU; exchange = grouped ncclSend/ncclRecv
 * (src/base/cs_halo.cpp:518-587) on the library's communication stream.
 *----------------------------------------------------------------------------*/

/* Communicator bootstrap (ncclGetUniqueId on rank 0, broadcast of the 128-byte
   id by the caller, ncclCommInitRank on every rank;
   src/base/cs_base_cuda.cu:201-262). NCCL is resolved with dlopen. */
int         csb_comm_get_unique_id(void *id_128_bytes);
void       *csb_comm_init_rank(const void *id_128_bytes, int n_ranks, int rank);
void        csb_comm_destroy(void *nccl_comm);

/* nccl_comm: an ncclComm_t (as void *) spanning all ranks, or NULL for a
   single-rank (periodic/self) halo.
   c_domain_rank[n_c_domains]; send_index[n_c_domains+1] into send_list;
   index[n_c_domains+1] = ghost sections (offsets from n_local_elts). */
csb_halo_t *csb_halo_create(void              *nccl_comm,
                            int                local_rank,
                            int                n_c_domains,
                            const int         *c_domain_rank,
                            csb_lnum_t         n_local_elts,
                            const csb_lnum_t  *send_index,
                            const csb_lnum_t  *send_list,
                            const csb_lnum_t  *index);
void        csb_halo_destroy(csb_halo_t **halo);

/* cs_halo_sync_var (pack + start + wait, src/base/cs_halo.cpp:1739,:1957,:2130)
   on a device array of n_local_elts + n_ghost doubles. */
int         csb_halo_sync_var(const csb_halo_t *halo, csb_real_t *var_d);

/*----------------------------------------------------------------------------
 * Peer-memory communicator: what the solvers use instead of MPI / NCCL calls
 * for the ghost exchange (cs_halo_sync_var, src/base/cs_halo.cpp:1739-2220),
 * the global sums of dot products (MPI_Allreduce, src/alge/cs_sles_it.cpp:
 * 372-378; ncclAllReduce, src/alge/cs_sles_it_cuda.cu:1000-1006) and the
 * exchanges of the hierarchy construction (src/alge/cs_grid.cpp:802-910,
 * :1366-1374, :6561-6581).  Each rank owns an arena in its GPU's memory that
 * every other rank of the NVSwitch domain maps; kernels store into peers'
 * arenas and publish counters (see csrc/comm.cu).  n_ranks <= 16.
 *
 * One process per GPU (the reference's model, src/base/cs_base_cuda.cu:752):
 *     c = csb_pcomm_create(n_ranks, rank, arena_bytes, window_bytes);
 *     csb_pcomm_get_ipc_handle(c, h);   MPI_Allgather / all_gather the 64 bytes
 *     csb_pcomm_open_peers(c, all_handles);
 * Ranks as host threads of one process (each thread calls csb_init first):
 *     g = csb_pcomm_group_create(n_ranks);               once
 *     c = csb_pcomm_create_in_group(g, rank, ...);       every thread (collective)
 * arena_bytes: room for the solvers' ghost buffers (16 bytes per ghost row and
 * level is enough); window_bytes: staging window, >= 24 bytes per ghost row.
 *----------------------------------------------------------------------------*/

csb_pcomm_t *csb_pcomm_create(int n_ranks, int rank, size_t arena_bytes, size_t window_bytes);
int          csb_pcomm_get_ipc_handle(csb_pcomm_t *c, void *handle_64_bytes);
int          csb_pcomm_open_peers(csb_pcomm_t *c, const void *all_handles /* n_ranks x 64 */);
csb_pcomm_group_t *csb_pcomm_group_create(int n_ranks);
void         csb_pcomm_group_destroy(csb_pcomm_group_t **group);
csb_pcomm_t *csb_pcomm_create_in_group(csb_pcomm_group_t *group, int rank, size_t arena_bytes,
                                       size_t window_bytes);
void         csb_pcomm_destroy(csb_pcomm_t **c);
int          csb_pcomm_rank(const csb_pcomm_t *c);
int          csb_pcomm_n_ranks(const csb_pcomm_t *c);
/* collective: in-place sum of n <= 8 host doubles over all ranks (rank order,
   identical result everywhere); csb_pcomm_barrier = the same with no payload */
int          csb_pcomm_allreduce_sum(csb_pcomm_t *c, double *vals, int n);
int          csb_pcomm_barrier(csb_pcomm_t *c);

/* Attach a halo to the communicator (collective over the neighbour ranks):
   c_domain_rank values are ranks of `c`.  csb_halo_sync_var and the solvers
   then exchange ghosts through peer memory instead of NCCL. */
int          csb_halo_set_pcomm(csb_halo_t *halo, csb_pcomm_t *c);

/*----------------------------------------------------------------------------
 * Iterative solvers (cs_sles_it_t)
 *----------------------------------------------------------------------------*/

/* cs_sles_it_create (src/alge/cs_sles_it.cpp:4659): poly_degree 0 = Jacobi preconditioner,
   1 / >= 2 = polynomial of degree 1 / 2 (PCG, FCG, IPCG; :4683-4690), < 0 = none, or the one
   handed over with csb_sles_it_transfer_pc. */
csb_sles_it_t *csb_sles_it_create(csb_sles_it_type_t  solver_type,
                                  int                 poly_degree,
                                  int                 n_max_iter,
                                  int                 update_stats);

/* cs_sles_it_set_fallback_threshold (src/alge/cs_sles_it.cpp:5791): when a solve ends in a
   state < threshold, csb_sles_it_solve runs the fallback solver -- GCR + the same Jacobi
   preconditioner, the reference's choice on a device (:5383-5385) -- from the current vx, or
   from zero when the first solver diverged (_fallback :4464-4528), and returns its state, the
   iteration counts added.  Defaults as cs_sles_it_create (:4736-4749):
   CSB_SLES_MAX_ITERATION for BiCGStab (it falls back on breakdown and divergence),
   CSB_SLES_DIVERGED (never) for the other solvers; n_iter_max = the solver's own. */
void csb_sles_it_set_fallback_threshold(csb_sles_it_t *c, csb_sles_convergence_state_t threshold,
                                        int n_iter_max);
/* cs_sles_it_set_restart_interval (src/alge/cs_sles_it.cpp:5810): GCR restart, default 20 */
void csb_sles_it_set_restart_interval(csb_sles_it_t *c, int interval);
/* 1 if the last csb_sles_it_solve of this context went through the fallback solver */
int  csb_sles_it_used_fallback(const csb_sles_it_t *c);

/* Global-sum communicator for dot products (cs_sles_it_set_mpi_reduce_comm,
   src/alge/cs_sles_it.cpp:5669): ncclComm_t as void *, NULL = single rank. */
int  csb_sles_it_set_reduce_comm(csb_sles_it_t *c, void *nccl_comm);
/* Same role with the peer-memory communicator (PCG: the whole distributed
   solve then runs as one persistent kernel per rank). */
int  csb_sles_it_set_pcomm(csb_sles_it_t *c, csb_pcomm_t *comm);

/* cs_sles_setup_t (src/alge/cs_sles.h:91-101) -> cs_sles_it_setup (:4964) */
void csb_sles_it_setup(void *context, const char *name, const csb_matrix_t *a,
                       int verbosity);

/* cs_sles_solve_t (src/alge/cs_sles.h:121-142) -> cs_sles_it_solve (:5194).
   vx_ini == vx: nonzero initial guess in vx; vx_ini == NULL: start from 0. */
csb_sles_convergence_state_t
csb_sles_it_solve(void                *context,
                  const char          *name,
                  const csb_matrix_t  *a,
                  int                  verbosity,
                  double               precision,
                  double               r_norm,
                  int                 *n_iter,
                  double              *residual,
                  const csb_real_t    *rhs,
                  csb_real_t          *vx_ini,
                  csb_real_t          *vx,
                  size_t               aux_size,
                  void                *aux_vectors);

void csb_sles_it_free(void *context);                 /* cs_sles_it_free :5436 */
void csb_sles_it_destroy(void **context);             /* cs_sles_it_destroy :4764 */
double csb_sles_it_get_last_initial_residual(const csb_sles_it_t *c);  /* :5505 */

/* Residual norm after each convergence test of the last solve (the data the
   reference plots through cs_time_plot, src/alge/cs_sles_it.cpp:264-274).
   Returns the number of values written (<= max_n). */
int  csb_sles_it_get_residual_history(const csb_sles_it_t *c, double *hist, int max_n);

/*----------------------------------------------------------------------------
 * Multigrid (cs_multigrid_t)
 *----------------------------------------------------------------------------*/

/* cs_multigrid_create (src/alge/cs_multigrid.cpp:4644) with the V-cycle
   defaults of :4693-4764. */
csb_multigrid_t *csb_multigrid_create(csb_multigrid_type_t mg_type);

/* cs_multigrid_set_solver_options (src/alge/cs_multigrid.cpp:5032) */
void csb_multigrid_set_solver_options(csb_multigrid_t    *mg,
                                      csb_sles_it_type_t  descent_smoother_type,
                                      csb_sles_it_type_t  ascent_smoother_type,
                                      csb_sles_it_type_t  coarse_solver_type,
                                      int                 n_max_cycles,
                                      int                 n_max_iter_descent,
                                      int                 n_max_iter_ascent,
                                      int                 n_max_iter_coarse,
                                      int                 poly_degree_descent,
                                      int                 poly_degree_ascent,
                                      int                 poly_degree_coarse,
                                      double              precision_mult_descent,
                                      double              precision_mult_ascent,
                                      double              precision_mult_coarse);

/* cs_multigrid_set_coarsening_options (src/alge/cs_multigrid.cpp:4945);
   coarsening_type 0 (default) / 1 = CS_GRID_COARSENING_SPD_DX; 3 = CS_GRID_COARSENING_SPD_PW
   (pairwise aggregation; needs p0p1_relax <= 0, i.e. Galerkin coarse matrices, one rank).
   p0p1_relax <= 0 selects Galerkin coarse matrices, which need no mesh association. */
void csb_multigrid_set_coarsening_options(csb_multigrid_t  *mg,
                                          int               aggregation_limit,
                                          int               coarsening_type,
                                          int               n_max_levels,
                                          unsigned long long min_g_rows,
                                          double            p0p1_relax,
                                          int               postprocess_block_size);

/* Number of row blocks the greedy aggregation is confined to; restates the
   reference's per-OpenMP-thread blocks (src/alge/cs_grid.cpp:4374-4393).
   n_blocks == reference run with that many OpenMP threads (maps are bit-identical per
   count); 0 (default): the OpenMP thread count of this process (OMP_NUM_THREADS), which
   is what the reference uses. */
void csb_multigrid_set_aggregation_blocks(csb_multigrid_t *mg, int n_blocks);

/* cs_multigrid_set_merge_options (src/alge/cs_multigrid.cpp:6039-6051): coarse levels whose
   mean row count per rank falls under rows_mean_threshold, or whose global row count falls
   under rows_glob_threshold, are merged onto fewer ranks (cs_grid_merge_ranks,
   src/alge/cs_grid_rank_merge.cpp:1359; trigger cs_grid.cpp:8355-8372).  On this path
   rank_stride >= the number of ranks (everything onto rank 0) is available; rank_stride <= 1
   (default, as in the reference) disables merging.  The merged levels keep the aggregates
   built on the partitioned levels (the reference re-aggregates across the former rank
   boundaries). */
void csb_multigrid_set_merge_options(csb_multigrid_t *mg, int rank_stride,
                                     int rows_mean_threshold,
                                     unsigned long long rows_glob_threshold);
/* level at which the last setup merged (-1: none) */
int  csb_multigrid_get_merge_level(const csb_multigrid_t *mg);

int  csb_multigrid_set_reduce_comm(csb_multigrid_t *mg, void *nccl_comm);
/* Communicator of a distributed hierarchy: required when the matrix has a halo
   (the reference takes cs_glob_mpi_comm, src/alge/cs_multigrid.cpp:4720). */
int  csb_multigrid_set_pcomm(csb_multigrid_t *mg, csb_pcomm_t *comm);

void csb_multigrid_setup(void *context, const char *name, const csb_matrix_t *a,
                         int verbosity);               /* cs_multigrid_setup :5275 */

csb_sles_convergence_state_t
csb_multigrid_solve(void                *context,      /* cs_multigrid_solve :5362 */
                    const char          *name,
                    const csb_matrix_t  *a,
                    int                  verbosity,
                    double               precision,
                    double               r_norm,
                    int                 *n_iter,
                    double              *residual,
                    const csb_real_t    *rhs,
                    csb_real_t          *vx_ini,
                    csb_real_t          *vx,
                    size_t               aux_size,
                    void                *aux_vectors);

void csb_multigrid_free(void *context);                /* cs_multigrid_free :5604 */
void csb_multigrid_destroy(void **context);            /* cs_multigrid_destroy :4793 */

/* Hierarchy inspection (what cs_grid_get_info / cs_grid_dump expose,
   src/alge/cs_grid.cpp:7740, :8839): used by the parity tests. */
int  csb_multigrid_get_n_levels(const csb_multigrid_t *mg);
/* 1 when the last csb_multigrid_setup completed (setup returns void like
   cs_multigrid_setup; on failure the message is in csb_last_error()) */
int  csb_multigrid_is_setup(const csb_multigrid_t *mg);
int  csb_multigrid_get_n_cycles(const csb_multigrid_t *mg);
/* info[0] = n_rows, info[1] = n_cols_ext, info[2] = n_entries (off-diagonal) */
int  csb_multigrid_get_level_info(const csb_multigrid_t *mg, int level, csb_lnum_t info[3]);
/* Halo of a level of a distributed hierarchy (cs_grid_get_halo, src/alge/cs_grid.cpp:
   7724): sizes[0] = n_c_domains, sizes[1] = n_send, sizes[2] = n_ghost; any array
   pointer may be NULL.  Returns -1 when the level has no halo. */
int  csb_multigrid_get_level_halo(const csb_multigrid_t *mg, int level, int sizes[3],
                                  int *c_domain_rank, csb_lnum_t *send_index,
                                  csb_lnum_t *send_list, csb_lnum_t *index);
/* Copy level data to host arrays (any pointer may be NULL).  coarse_row has
   the parent level's n_rows entries and maps parent rows -> rows of `level`. */
int  csb_multigrid_get_level_arrays(const csb_multigrid_t *mg, int level,
                                    csb_lnum_t *coarse_row, csb_lnum_t *row_index,
                                    csb_lnum_t *col_id, csb_real_t *d_val,
                                    csb_real_t *x_val);
/* Fine-level residual norm at the start of each cycle of the last solve. */
int  csb_multigrid_get_cycle_residuals(const csb_multigrid_t *mg, double *res, int max_n);
/* Seconds spent in the last setup, split: [0] total, [1] aggregation (host),
   [2] coarse matrix assembly, [3] transfers. */
int  csb_multigrid_get_setup_times(const csb_multigrid_t *mg, double t[4]);

/* Device-side phase timers of the last solve (persistent V-cycle kernel):
   seconds[level*6 + phase], phase = 0 descent smoother, 1 residual, 2 restrict,
   3 coarsest solve, 4 prolong, 5 ascent smoother; the instrument the reference
   offers through its per-level timers (src/alge/cs_multigrid.cpp:170-205).
   Returns the number of levels written (<= max_levels). */
int  csb_multigrid_get_phase_times(const csb_multigrid_t *mg, double *seconds, int max_levels);
/* The same timers inside the PCG smoothers (descent + ascent + coarsest solve):
   seconds[level*3 + k], k = 0 residual SpMV (r = A.vx - rhs), 1 SpMV + 3 dots,
   2 fused vector updates. */
int  csb_multigrid_get_smoother_times(const csb_multigrid_t *mg, double *seconds, int max_levels);

/* Host-only pieces of the hierarchy construction (no device needed), exposed
   so that CPU-side tests can check them: the DX aggregation map
   (_automatic_aggregation_dx_msr, src/alge/cs_grid.cpp:4282) and the coarse MSR
   structure + coarse face numbering (_coarse_msr_struct :3020,
   _msr_face_adjacency :3751).  csb_host_coarse_structure returns the number of
   coarse off-diagonal entries; output arrays must hold at least the fine nnz. */
int  csb_host_aggregate_dx_msr(int level, csb_lnum_t n_rows, const csb_lnum_t *row_index,
                               const csb_lnum_t *col_id, const csb_real_t *d_val,
                               const csb_real_t *x_val, int max_aggregation, int n_blocks,
                               csb_lnum_t *f_c_row);
/* One coarsening step of CS_GRID_COARSENING_SPD_PW with Galerkin coarse values
   (_pairwise_msr, src/alge/cs_grid.cpp:1744-2020; recursion for aggregation limits 4 and 8,
   :8305-8346; values :6040-6090), entirely on the host.  Output arrays must hold n_rows + 1 /
   the fine nnz / n_rows / the fine nnz entries; returns the number of coarse rows. */
int  csb_host_coarsen_pairwise(int level, csb_lnum_t n_rows, const csb_lnum_t *row_index,
                               const csb_lnum_t *col_id, const csb_real_t *d_val,
                               const csb_real_t *x_val, int aggregation_limit,
                               csb_lnum_t *f_c_row, csb_lnum_t *c_row_index,
                               csb_lnum_t *c_col_id, csb_real_t *c_d_val, csb_real_t *c_x_val);
int  csb_host_coarse_structure(csb_lnum_t f_n_rows, const csb_lnum_t *f_row_index,
                               const csb_lnum_t *f_col_id, const csb_lnum_t *f_c_row,
                               csb_lnum_t c_n_rows, csb_lnum_t *c_row_index,
                               csb_lnum_t *c_col_id, csb_lnum_t *c_cell_face,
                               short *c_cell_face_sgn, csb_lnum_t *c_n_faces);

/*----------------------------------------------------------------------------
 * Multigrid as a preconditioner (cs_multigrid_pc_create, src/alge/cs_multigrid.cpp:5682;
 * _multigrid_pc_apply :1500-1600; _multigrid_v_cycle_pc :3569-4000) and the multigrid
 * smoothers (src/alge/cs_multigrid_smoother.h:77-165).
 *
 * The reference's production pressure solver is FCG preconditioned by one V-cycle:
 *   c  = csb_sles_it_create(CSB_SLES_FCG, -1, n_max_iter, 1);
 *   pc = csb_multigrid_pc_create(CSB_MULTIGRID_V_CYCLE);
 *   csb_sles_it_transfer_pc(c, &pc);              // cs_sles_it_transfer_pc, cs_sles_it.h:379
 * csb_multigrid_solve with n_max_cycles == 1 runs the same single cycle without
 * convergence tests (cs_multigrid_solve, cs_multigrid.cpp:5484-5499) and returns
 * CSB_SLES_ITERATING like the reference.
 *----------------------------------------------------------------------------*/

/* V-cycle with the reference's device settings for a preconditioner: RJ3 smoothers
   (3 sweeps), PCG coarse solver (500), one cycle, precision_mult (-1, -1, 1). */
csb_multigrid_t *csb_multigrid_pc_create(csb_multigrid_type_t mg_type);

/* The solver takes ownership of the preconditioner; *pc is set to NULL.
   CSB_SLES_FCG, _IPCG or _PCG created with poly_degree < 0 only; with no preconditioner transferred such
   a solver runs unpreconditioned (identity).  Returns 0, or -1 (see csb_last_error). */
int  csb_sles_it_transfer_pc(csb_sles_it_t *c, csb_multigrid_t **pc);
csb_multigrid_t *csb_sles_it_get_pc(csb_sles_it_t *c);   /* cs_sles_it_get_pc */

/* cs_multigrid_smoother_create / _setup / _solve (cs_multigrid_smoother.cpp:1726, :1825,
   :2008): n_iter sweeps (rounded up to whole schedules for RJ2 / RJ3 at setup), no
   convergence test, *residual = -1, returns CSB_SLES_MAX_ITERATION.  Free / destroy / log /
   copy: csb_sles_it_free, _destroy, _log, _copy (the reference uses cs_sles_it_destroy). */
csb_sles_it_t *csb_multigrid_smoother_create(csb_sles_it_type_t smoother_type,
                                             int poly_degree, int n_iter);
void csb_multigrid_smoother_setup(void *context, const char *name, const csb_matrix_t *a,
                                  int verbosity);
csb_sles_convergence_state_t
csb_multigrid_smoother_solve(void                *context,
                             const char          *name,
                             const csb_matrix_t  *a,
                             int                  verbosity,
                             double               precision,
                             double               r_norm,
                             int                 *n_iter,
                             double              *residual,
                             const csb_real_t    *rhs,
                             csb_real_t          *vx_ini,
                             csb_real_t          *vx,
                             size_t               aux_size,
                             void                *aux_vectors);

/*----------------------------------------------------------------------------
 * cs_sles_t registry boundary (src/alge/cs_sles.cpp:1385-1425): these are the
 * six callbacks a maintainer passes to cs_sles_define(); see INTEGRATION.md.
 *   setup  : csb_multigrid_setup   / csb_sles_it_setup
 *   solve  : csb_multigrid_solve   / csb_sles_it_solve
 *   free   : csb_multigrid_free    / csb_sles_it_free
 *   log    : csb_multigrid_log     / csb_sles_it_log
 *   copy   : csb_multigrid_copy    / csb_sles_it_copy
 *   destroy: csb_multigrid_destroy / csb_sles_it_destroy
 *----------------------------------------------------------------------------*/

void  csb_sles_it_log(const void *context, int log_type);
void *csb_sles_it_copy(const void *context);
void  csb_multigrid_log(const void *context, int log_type);
void *csb_multigrid_copy(const void *context);

#ifdef __cplusplus
}
#endif

#endif /* CSB200_H */
