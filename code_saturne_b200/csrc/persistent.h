/* persistent.h -- descriptors of the persistent cooperative kernels (persistent.cu). */
#ifndef CSB_PERSISTENT_H
#define CSB_PERSISTENT_H

#include "csb_internal.h"

namespace csb {

/* One level of the hierarchy as seen by the device-side V-cycle. */
struct MgLevelDesc {
  int           n_rows;
  int           n_slabs;       /* ceil(n_rows / slab_rows) */
  int           slab_rows;     /* <= 256 rows streamed per CTA step */
  int           n_rows_next;
  int           level_id;      /* index of the level (slot of its exchange counter) */
  int           gather_depth;  /* 2, or 3 when no row has more than 8 entries (persistent.cu) */
  int           row_len_max;   /* longest row (entries) */
  int           r_slab_rows;   /* slab plan of the restriction to the next level (rows = rows of
                                  the next level, entries = this level's rows); 0: none */
  int           r_n_slabs;
  const int    *row_index;
  const int    *col_id;
  const double *d_val;
  const double *x_val;
  const double *ad_inv;
  const double *l1_inv;        /* L1 Jacobi smoother only (cs_sles_it_priv.cpp:231-262) */
  double       *rhs;           /* levels > 0 */
  double       *vx;            /* levels > 0 */
  const int    *coarse_row;    /* rows of this level -> rows of the next level */
  const int    *c_f_row_index; /* rows of the next level -> rows of this level */
  const int    *c_f_row_ids;
  ItState      *st_descent;    /* also used for the coarsest solve */
  ItState      *st_ascent;
  double        n_g_rows;      /* rows of this level over all ranks */
  const DistLevel *dist;       /* ghost exchange of this level; nullptr on one rank */
};

struct PersistentCommon {
  double   *work;          /* 5 * work_stride doubles: rk dk gk zk sk */
  size_t    work_stride;
  double   *partials;      /* 12 * grid doubles */
  unsigned *barrier;       /* 2 words, zero-initialised */
  const CommDev *comm;     /* nullptr on one rank */
};

constexpr int PERSISTENT_S_RED = 136;   /* doubles of reduction scratch ahead of the slab */

struct MgSolveParams {
  int                n_levels;
  const MgLevelDesc *levels;     /* device array */
  MgState           *ms;
  const double      *rhs;
  double            *vx;
  double            *wr;
  int                vx_ini_nonzero;
  int                n_max_cycles;
  int                n_max_iter[3];
  double             precision_mult[3];
  double             precision, r_norm;
  int                single_rows;  /* levels up to this size run in one CTA */
  int                small_rows;   /* coarsest level up to this size: shared-memory PCG */
};

/* Sweep schedule of a Jacobi-family cs_multigrid_smoother (cs_multigrid_smoother.cpp:143-373,
   :804-905): mode = SmoothMode, n_iter sweeps, relaxation weight w[it % wk_m]. */
struct SmoothPlanDev {
  int    mode, wk_m, n_iter;
  double w[3];
};

/* FCG preconditioned by one V-cycle (cs_sles_it FCG + cs_multigrid_pc, the reference's
   production pressure-solver layout), whole solve in one launch: k_fcg_mg_persistent. */
struct FcgMgParams {
  int                n_levels;
  const MgLevelDesc *levels;     /* device array */
  MgState           *ms;         /* phase timers */
  ItState           *st;         /* FCG state, residual history */
  const double      *rhs;
  double            *vx;
  int                vx_ini_nonzero;
  int                n_max_iter;           /* FCG */
  double             precision, r_norm;
  SmoothPlanDev      sm[2];                /* descent, ascent smoothers */
  int                coarse_n_max_iter;    /* coarsest-level PCG */
  double             coarse_precision_mult;
  int                single_rows;          /* levels up to this size run in one CTA */
  int                small_rows;           /* coarsest level up to this size: shared-memory PCG */
  double            *wr, *tmp;             /* residual; second sweep buffer (level-0 sized) */
  double            *fwork;                /* FCG vectors rk vk wk dk qk */
  size_t             fstride;
  const MergeDev    *merge;                /* coarse levels merged onto rank 0, or nullptr */
  int                merge_level;          /* first merged level (-1: none) */
};

size_t persistent_smem_bytes(size_t max_slab_entries);
int    persistent_grid(int which, size_t smem);   /* 0: PCG, 1: multigrid, 2: FCG, 3: IPCG,
                                                      4: FCG + multigrid kernel */

bool launch_pcg_persistent(const MgLevelDesc &L, const double *rhs, double *vx,
                           bool vx_ini_nonzero, int n_max_iter, double precision,
                           double r_norm, const PersistentCommon &P, ItState *st,
                           int grid, size_t smem, bool single, bool cooperative);

bool launch_mg_persistent(const MgSolveParams &S, const PersistentCommon &P, int grid,
                          size_t smem, bool cooperative);

/* flexible CG (cs_sles_it.cpp:832-985): same arguments as the PCG kernel */
bool launch_fcg_persistent(const MgLevelDesc &L, const double *rhs, double *vx,
                           bool vx_ini_nonzero, int n_max_iter, double precision,
                           double r_norm, const PersistentCommon &P, ItState *st,
                           int grid, size_t smem, bool single, bool cooperative,
                           bool ipcg = false);   /* true: _conjugate_gradient_ip (:1016-1203) */

bool launch_fcg_mg_persistent(const FcgMgParams &S, const PersistentCommon &P, int grid,
                              size_t smem, bool cooperative);

/* y = A.x (or E.x) through the bulk-copy pipeline of the persistent kernels, as a
   stand-alone launch (cs_matrix_vector_multiply).  Returns false with no error set when
   the pipeline cannot take these buffers (alignment, odd length): the caller falls back
   to k_spmv_stream. */
bool launch_spmv_pipeline(const _csb_matrix_t *m, const double *x, double *y, bool exclude_diag,
                          bool *launched);

/* gather depth of the SpMV phases for a matrix with these row lengths (CSB_GATHER_DEPTH) */
int plan_gather_depth(const std::vector<csb_lnum_t> &row_index, int n_rows, int *row_len_max);

/* choose the slab height of a level so that a slab's products fit in ~40 KB */
void plan_slabs(const std::vector<csb_lnum_t> &row_index, int n_rows, int &slab_rows,
                int &n_slabs, size_t &max_entries);

} // namespace csb

#endif
