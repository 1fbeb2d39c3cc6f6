/* oracle/cs_oracle.h -- TEST INFRASTRUCTURE (see cs_oracle.c header).
 * CPU restatement of the reference's src/alge hot path; parity pinned against
 * the reference's own compiled code (oracle/_ref). Never used by the product. */
#ifndef CS_ORACLE_H
#define CS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* cs_sles_convergence_state_t (src/alge/cs_sles.h:57-68) */
#define ORA_DIVERGED      -3
#define ORA_BREAKDOWN     -2
#define ORA_MAX_ITERATION -1
#define ORA_ITERATING      0
#define ORA_CONVERGED      1

/* One level of the hierarchy: subset of cs_grid_t (src/alge/cs_grid_priv.h:53-162) */
typedef struct {
  int      level, n_rows, nnz, n_faces;
  int     *row_index, *col_id;
  double  *d_val, *x_val;
  int     *coarse_row;      /* parent rows -> rows of this level (level > 0) */
  int     *cell_face;       /* per off-diagonal entry */
  short   *cell_face_sgn;
  double  *cell_cen, *cell_vol;
  double  *face_normal;     /* level > 0 */
  double  *face_u_normal, *face_surf;  /* level 0 */
  double  *xa0, *xa0ij;     /* level > 0 */
  int      n_cols_ext;      /* n_rows + ghost rows (== n_rows on one rank) */
  int     *face_cell;       /* hierarchy of a matrix mapped from native arrays   */
  double  *xa;              /* (cs_oracle_native.c): faces and per-face values    */
} ora_grid_t;

typedef struct {
  int          n_levels;
  ora_grid_t **grids;
} ora_mg_t;

typedef struct {
  int     n_max_cycles;
  int     n_max_iter[3];      /* descent, ascent, coarse */
  double  precision_mult[3];
} ora_mg_opt_t;

/* cs_sles_it_type_t values used on this path (src/alge/cs_sles_it.h:55-87) */
#define ORA_SLES_PCG        0
#define ORA_SLES_JACOBI     3
#define ORA_SLES_P_GAUSS_SEIDEL      8   /* cs_sles_it_type_t values (cs_sles_it.h:59-86) */
#define ORA_SLES_P_SYM_GAUSS_SEIDEL  9
#define ORA_SLES_TS_F_GAUSS_SEIDEL  17
#define ORA_SLES_TS_B_GAUSS_SEIDEL  18
#define ORA_SLES_L1_JACOBI 13
#define ORA_SLES_R_JACOBI  14
#define ORA_SLES_RJ2       15
#define ORA_SLES_RJ3       16

/* options of the single-cycle (preconditioner) form of the V-cycle */
typedef struct {
  int     type[3];            /* descent, ascent, coarse: cs_sles_it_type_t */
  int     n_max_iter[3];
  double  precision_mult[3];  /* < 0: cs_multigrid_smoother (no convergence test) */
} ora_mg_pc_opt_t;

void   ora_thread_range(int n, int type_size, int t_id, int n_t, int *s_id, int *e_id);

void   ora_spmv_msr(int n_rows, const int *row_index, const int *col_id,
                    const double *d_val, const double *x_val, int exclude_diag,
                    const double *x, double *y);

void   ora_spmv_csr(int n_rows, const int *row_index, const int *col_id, const double *val,
                    int exclude_diag, const double *x, double *y);
double ora_dot(int n, const double *x, const double *y);
double ora_dot_xx(int n, const double *x);
void   ora_dot_xx_xy(int n, const double *x, const double *y, double *xx, double *xy);
void   ora_dot_xy_yz(int n, const double *x, const double *y, const double *z,
                     double *xy, double *yz);
void   ora_dot_xx_xy_yz(int n, const double *x, const double *y, const double *z,
                        double *xx, double *xy, double *yz);

int    ora_pcg_sr_p0(int n_rows, const int *row_index, const int *col_id,
                     const double *d_val, const double *x_val,
                     int n_max_iter, double precision, double r_norm,
                     const double *rhs, double *vx, int use_vx_ini,
                     double *out, double *history, int history_size);
int    ora_fcg(int n_rows, const int *row_index, const int *col_id,
               const double *d_val, const double *x_val,
               int n_max_iter, double precision, double r_norm,
               const double *rhs, double *vx, int use_vx_ini,
               double *out, double *history, int history_size);
int    ora_ipcg(int n_rows, const int *row_index, const int *col_id,
                const double *d_val, const double *x_val,
                int n_max_iter, double precision, double r_norm,
                const double *rhs, double *vx, int use_vx_ini,
                double *out, double *history, int history_size);
int    ora_pcg_sr_npc(int n_rows, const int *row_index, const int *col_id,
                      const double *d_val, const double *x_val,
                      int n_max_iter, double precision, double r_norm,
                      const double *rhs, double *vx, int use_vx_ini,
                      double *out, double *history, int history_size);
int    ora_jacobi(int n_rows, const int *row_index, const int *col_id,
                  const double *d_val, const double *x_val,
                  int n_max_iter, double precision, double r_norm,
                  const double *rhs, double *vx, int use_vx_ini,
                  double *out, double *history, int history_size);
/* _p_gauss_seidel_msr (sym 0) / _p_sym_gauss_seidel_msr (sym 1), cs_sles_it.cpp:3940, :4124 */
int    ora_gauss_seidel(int sym, int n_rows, const int *row_index, const int *col_id,
                        const double *d_val, const double *x_val,
                        int n_max_iter, double precision, double r_norm,
                        const double *rhs, double *vx, int use_vx_ini,
                        double *out, double *history, int history_size);
void   ora_spmv_msr_b3(int n_rows, const int *row_index, const int *col_id,
                       const double *d_val, const double *x_val, int exclude_diag,
                       const double *x, double *y);
void   ora_spmv_msr_bn(int n_rows, int db, const int *row_index, const int *col_id,
                       const double *d_val, const double *x_val, int exclude_diag,
                       const double *x, double *y);
/* _gcr (cs_sles_it.cpp:3272-3480): Jacobi-preconditioned (jacobi != 0) GCR with restart */
int    ora_gcr(int n_rows, int db, const int *row_index, const int *col_id,
               const double *d_val, const double *x_val, int jacobi, int restart,
               int n_max_iter, double precision, double r_norm,
               const double *rhs, double *vx, int use_vx_ini,
               double *out, double *history, int history_size);
/* cs_multigrid_smoother on block-diagonal matrices: block Jacobi (type 3), block relaxed
   Jacobi (14, 15, 16); cs_multigrid_smoother.cpp:390-790 */
int    ora_smoother_block(int type, int n_rows, int db, const int *row_index, const int *col_id,
                          const double *d_val, const double *x_val, int n_max_iter,
                          const double *rhs, double *vx, int use_vx_ini, double *out);
int    ora_block_jacobi(int n_rows, int db, const int *row_index, const int *col_id,
                        const double *d_val, const double *x_val,
                        int n_max_iter, double precision, double r_norm,
                        const double *rhs, double *vx, int use_vx_ini,
                        double *out, double *history, int history_size);
int    ora_bicgstab_b(int n_rows, int db, const int *row_index, const int *col_id,
                      const double *d_val, const double *x_val,
                      int n_max_iter, double precision, double r_norm,
                      const double *rhs, double *vx, int use_vx_ini,
                      double *out, double *history, int history_size);
int    ora_bicgstab(int n_rows, const int *row_index, const int *col_id,
                    const double *d_val, const double *x_val,
                    int n_max_iter, double precision, double r_norm,
                    const double *rhs, double *vx, int use_vx_ini,
                    double *out, double *history, int history_size);

int    ora_aggregation_dx_msr(int level, int f_n_rows, const int *row_index,
                              const int *col_id, const double *d_val,
                              const double *x_val, int max_aggregation,
                              int n_threads, int *f_c_row);

ora_grid_t *ora_grid_create_fine(int n_rows, const int *row_index, const int *col_id,
                                 const double *d_val, const double *x_val,
                                 int n_faces, const int *c2f, const short *c2f_sgn,
                                 const double *cell_cen, const double *cell_vol,
                                 const double *face_u_normal, const double *face_surf);
ora_grid_t *ora_grid_create_fine_ext(int n_rows, int n_cols_ext, const int *row_index,
                                     const int *col_id, const double *d_val,
                                     const double *x_val, int n_faces, const int *c2f,
                                     const short *c2f_sgn, const double *cell_cen,
                                     const double *cell_vol, const double *face_u_normal,
                                     const double *face_surf);
ora_grid_t *ora_grid_coarsen_begin(const ora_grid_t *f, const int *f_c_row, int c_n_rows,
                                   int c_n_cols_ext);
void        ora_grid_coarsen_finish(ora_grid_t *c, double relax_param);
ora_grid_t *ora_grid_coarsen(const ora_grid_t *f, int max_aggregation,
                             double relax_param, int n_threads);
void        ora_grid_free(ora_grid_t *g);

ora_mg_t   *ora_mg_setup(ora_grid_t *fine, int aggregation_limit, int n_levels_max,
                         int n_g_rows_min, double p0p1_relax, int n_threads);
void        ora_mg_free(ora_mg_t *mg, int free_fine);

/* cs_diffusion_potential without flux reconstruction: cs_oracle_rhs.c */
void   ora_diffusion_potential(int n_cells, int n_i_faces, const int *i_face_cells,
                               int n_b_faces, const int *b_face_cells, int init, int inc,
                               const double *pvar, const double *cofaf, const double *cofbf,
                               const double *i_visc, const double *b_visc, double *diverg);

void   ora_gradient_lsq(int n_cells, const int *row_index, const int *col_id, int n_i_faces,
                        const int *i_face_cells, int n_b_faces, const int *b_face_cells,
                        const double *cell_cen, const double *b_face_u_normal,
                        const double *b_dist, int inc, const double *pvar, const double *coefa,
                        const double *coefb, double *grad);
void   ora_diffusion_potential_rc(int n_cells, int n_i_faces, const int *i_face_cells,
                                  int n_b_faces, const int *b_face_cells, const double *i_geo,
                                  const double *i_face_surf, int irecmf, int init, int inc,
                                  const double *pvar, const double *grad, const double *cofaf,
                                  const double *cofbf, const double *i_visc,
                                  const double *b_visc, const double *visel, double *diverg);

/* forms with II' of the boundary faces (diipb, 3 per boundary face; NULL = the plain forms):
   value at I' per boundary face (cs_gradient_boundary_iprime_lsq_s), then the gradient and
   the boundary flux with it */
void   ora_b_iprime_lsq(int n_cells, const int *row_index, const int *col_id, int n_b_faces,
                        const int *b_face_cells, const double *cell_cen,
                        const double *b_face_u_normal, const double *b_dist,
                        const double *diipb, double clip_coeff, const double *pvar,
                        const double *coefa, const double *coefb, double *val_ip);
void   ora_gradient_lsq_b(int n_cells, const int *row_index, const int *col_id, int n_i_faces,
                          const int *i_face_cells, int n_b_faces, const int *b_face_cells,
                          const double *cell_cen, const double *b_face_u_normal,
                          const double *b_dist, const double *diipb, const double *val_ip,
                          int inc, const double *pvar, const double *coefa,
                          const double *coefb, double *grad);
void   ora_diffusion_potential_rc_b(int n_cells, int n_i_faces, const int *i_face_cells,
                                    int n_b_faces, const int *b_face_cells, const double *i_geo,
                                    const double *i_face_surf, const double *val_ip, int irecmf,
                                    int init, int inc, const double *pvar, const double *grad,
                                    const double *cofaf, const double *cofbf,
                                    const double *i_visc, const double *b_visc,
                                    const double *visel, double *diverg);

/* Green-Gauss gradient with least-squares face values (imrgra = 4): reconstruction pass over
   the least-squares gradient r_grad (_reconstruct_scalar_gradient) */
void   ora_gradient_green_reconstruct(int n_cells, int n_i_faces, const int *i_face_cells,
                                      int n_b_faces, const int *b_face_cells,
                                      const double *weight, const double *dofij,
                                      const double *i_face_surf, const double *i_face_u_normal,
                                      const double *b_face_surf, const double *b_face_u_normal,
                                      const double *cell_vol, const double *val_ip, int inc,
                                      const double *pvar, const double *coefa,
                                      const double *coefb, const double *r_grad,
                                      const double *corr_grad_lin, double *grad);

/* hydrostatic pressure treatment (hyd_p_flag = 1, exterior force f_ext [n_cells][3]): values at
   I' and face values of the boundary faces, least-squares gradient, Green-Gauss pass */
void   ora_b_face_values_hyd(int n_cells, const int *row_index, const int *col_id,
                             const int *c2f, int n_b_faces, const int *b_face_cells,
                             const double *cell_cen, const double *b_face_u_normal,
                             const double *b_dist, const double *diipb,
                             const double *i_face_cog, const double *b_face_cog,
                             const double *f_ext, double clip_coeff, int inc,
                             const double *pvar, const double *coefa, const double *coefb,
                             double *val_ip, double *val_f);
void   ora_gradient_lsq_hyd(int n_cells, const int *row_index, const int *col_id, int n_i_faces,
                            const int *i_face_cells, int n_b_faces, const int *b_face_cells,
                            const double *cell_cen, const double *b_face_u_normal,
                            const double *b_dist, const double *diipb, const double *i_face_cog,
                            const double *b_face_cog, const double *f_ext, const double *val_f,
                            const double *pvar, double *grad);
void   ora_gradient_green_reconstruct_hyd(int n_cells, int n_i_faces, const int *i_face_cells,
                                          int n_b_faces, const int *b_face_cells,
                                          const double *weight, const double *dofij,
                                          const double *i_face_surf,
                                          const double *i_face_u_normal,
                                          const double *b_face_surf,
                                          const double *b_face_u_normal, const double *cell_vol,
                                          const double *cell_cen, const double *i_face_cog,
                                          const double *f_ext, const double *val_f,
                                          const double *pvar, const double *r_grad,
                                          const double *corr_grad_lin, double *grad);

/* gradient limiter (eqp->imligr 0 / 1 / 2 with the factor climgr), applied to grad in place */
void   ora_gradient_clip(int n_cells, const int *row_index, const int *col_id, const int *c2f,
                         const short *c2f_sgn, int n_b_faces, const int *b_face_cells,
                         const double *cell_cen, const double *i_geo, int clip_mode,
                         double climgp, const double *pvar, const double *val_f, double *grad);

/* iterative Green-Gauss gradient (imrgra = 0); returns the number of sweeps */
int    ora_gradient_green_iter(int n_cells, int n_i_faces, const int *i_face_cells,
                               int n_b_faces, const int *b_face_cells, const double *weight,
                               const double *dofij, const double *i_face_surf,
                               const double *i_face_u_normal, const double *b_face_surf,
                               const double *b_face_u_normal, const double *cell_vol,
                               const double *diipb, int nswrgp, double epsrgp, int inc,
                               const double *pvar, const double *coefa, const double *coefb,
                               double *grad);

/* cs_divergence and cs_face_diffusion_potential (grad NULL: no reconstruction) */
void   ora_divergence(int n_cells, int n_i_faces, const int *i_face_cells, int n_b_faces,
                      const int *b_face_cells, int init, const double *i_massflux,
                      const double *b_massflux, double *diverg);
void   ora_face_diffusion_potential(int n_cells, int n_i_faces, const int *i_face_cells,
                                    int n_b_faces, const int *b_face_cells, const double *i_geo,
                                    const double *i_face_surf, const double *val_ip, int init,
                                    int inc, const double *pvar, const double *grad,
                                    const double *cofaf, const double *cofbf,
                                    const double *i_visc, const double *b_visc,
                                    const double *visel, double *i_massflux,
                                    double *b_massflux);

/* matrix assembler on one rank, MSR target: cs_oracle_assembler.c */
int    ora_assembler_structure(int n_rows, int n, const long long *row_g_id,
                               const long long *col_g_id, long long l_range_0,
                               int *row_index, int *col_id);
int    ora_assembler_values_add(int n_rows, const int *row_index, const int *col_id, int n,
                                const long long *row_g_id, const long long *col_g_id,
                                long long l_range_0, const double *val, double *d_val,
                                double *x_val);

/* matrix mapped from native (face-based) arrays: cs_oracle_native.c */
ora_grid_t *ora_grid_coarsen_galerkin_map(const ora_grid_t *f, const int *f_c_row, int c_n_rows);
int    ora_aggregation_fc(int level, int f_n_cells, int f_n_cells_ext, int f_n_faces,
                          const int *face_cell, const double *da, const double *xa,
                          int max_aggregation, int *f_c_cell);
int    ora_coarsen_face_cell(int f_n_faces, const int *f_face_cell, const int *coarse_row,
                             int c_n_cells, int *coarse_face, int *c_face_cell);
ora_grid_t *ora_grid_create_fine_native(int n_rows, int n_faces, const int *face_cell,
                                        const double *da, const double *xa,
                                        const double *cell_cen, const double *cell_vol,
                                        const double *face_u_normal, const double *face_surf);
ora_grid_t *ora_grid_coarsen_native(const ora_grid_t *f, int max_aggregation,
                                    double relax_param, int n_threads);
ora_mg_t   *ora_mg_setup_native(ora_grid_t *fine, int aggregation_limit, int n_levels_max,
                                int n_g_rows_min, double p0p1_relax, int n_threads);

void   ora_restrict(int f_n_rows, int c_n_rows, const int *coarse_row,
                    const double *f_var, double *c_var);
void   ora_prolong_inc(int f_n_rows, const int *coarse_row,
                       const double *c_var, double *f_var);

int    ora_mg_solve(const ora_mg_t *mg, const ora_mg_opt_t *opt,
                    double precision, double r_norm,
                    const double *rhs, double *vx, int use_vx_ini,
                    double *out, double *cycle_res);

void   ora_matrix_compute_coeffs_sym(int n_rows, const int *row_index, const int *cell_i_faces,
                                     int n_b_faces, const int *b_face_cells,
                                     int idiffp, int ndircp, double thetap, double relaxp,
                                     const double *cofbfp, const double *rovsdt,
                                     const double *i_visc, const double *b_visc,
                                     double *da, double *ea);
void   ora_matrix_compute_coeffs(int n_rows, const int *row_index, const int *col_id,
                                 const int *cell_i_faces, const short *cell_i_faces_sgn,
                                 int n_b_faces, const int *b_face_cells,
                                 int iconvp, int idiffp, int ndircp, double thetap,
                                 double relaxp, const double *coefbp, const double *cofbfp,
                                 const double *rovsdt, const double *i_massflux,
                                 const double *b_massflux, const double *i_visc,
                                 const double *b_visc, const double *xcpp,
                                 double *da, double *ea);
void   ora_matrix_compute_coeffs_v3(int n_rows, const int *row_index,
                                    const int *cell_i_faces, const short *cell_i_faces_sgn,
                                    int n_b_faces, const int *b_face_cells,
                                    int iconvp, int idiffp, int ndircp, double thetap,
                                    double relaxp, const double *coefbp, const double *cofbfp,
                                    const double *fimp, const double *i_massflux,
                                    const double *b_massflux, const double *i_visc,
                                    const double *b_visc, double *da, double *ea);
int    ora_pairwise_msr(int level, int f_n_rows, const int *row_index, const int *col_id,
                        const double *d_val, const double *x_val, int *f_c_row);
ora_mg_t *ora_mg_setup_pw(ora_grid_t *fine, int aggregation_limit, int n_levels_max,
                          int n_g_rows_min);
void   ora_gauss_inverse_33(const double a[9], double b[9]);
int    ora_block3_jacobi(int n_rows, const int *row_index, const int *col_id,
                         const double *d_val, const double *x_val,
                         int n_max_iter, double precision, double r_norm,
                         const double *rhs, double *vx, int use_vx_ini,
                         double *out, double *history, int history_size);
int    ora_smoother(int type, int n_rows, const int *row_index, const int *col_id,
                    const double *d_val, const double *x_val, int n_max_iter,
                    const double *rhs, double *vx, int use_vx_ini, double *out);
int    ora_mg_cycle_pc(const ora_mg_t *mg, const ora_mg_pc_opt_t *opt,
                       double precision, double r_norm,
                       const double *rhs, double *vx, int use_vx_ini, double *out);
int    ora_it_mgpc(int solver, const ora_mg_t *mg, const ora_mg_pc_opt_t *opt,
                   int n_max_iter, double precision, double r_norm,
                   const double *rhs, double *vx, int use_vx_ini,
                   double *out, double *history, int history_size);
int    ora_it_poly(int solver, int n_rows, const int *row_index, const int *col_id,
                   const double *d_val, const double *x_val, int poly_degree,
                   int n_max_iter, double precision, double r_norm,
                   const double *rhs, double *vx, int use_vx_ini,
                   double *out, double *history, int history_size);
int    ora_mg_solve_sm(const ora_mg_t *mg, const ora_mg_pc_opt_t *opt, int n_max_cycles,
                       double precision, double r_norm,
                       const double *rhs, double *vx, int use_vx_ini,
                       double *out, double *cycle_res);
int    ora_mg_k_cycle(const ora_mg_t *mg, const ora_mg_pc_opt_t *opt, double k_cycle_threshold,
                      double precision, double r_norm, const double *rhs, double *vx,
                      double *out);
int    ora_it_kpc(int solver, const ora_mg_t *mg, const ora_mg_pc_opt_t *opt,
                  double k_cycle_threshold, int n_max_iter, double precision, double r_norm,
                  const double *rhs, double *vx, int use_vx_ini,
                  double *out, double *history, int history_size);
int    ora_fcg_mgpc(const ora_mg_t *mg, const ora_mg_pc_opt_t *opt,
                    int n_max_iter, double precision, double r_norm,
                    const double *rhs, double *vx, int use_vx_ini,
                    double *out, double *history, int history_size);

#ifdef __cplusplus
}
#endif
#endif
