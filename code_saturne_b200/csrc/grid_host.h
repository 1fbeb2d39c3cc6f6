/* grid_host.h -- host-side hierarchy construction helpers (see grid_host.cpp). */
#ifndef CSB_GRID_HOST_H
#define CSB_GRID_HOST_H

#include <vector>

namespace csb {

/* Fine -> coarse row map of the reference's DX aggregation for a symmetric
   scalar MSR matrix (-2 never remains; -1 = penalised row).  n_blocks > 1
   confines aggregates to independent row blocks, as the reference does per
   OpenMP thread.  Returns the number of coarse rows. */
int aggregate_dx_msr(int level, int n_rows, const int *row_index, const int *col_id,
                     const double *d_val, const double *x_val, int max_aggregation,
                     int n_blocks, int *f_c_row);

struct CoarseStructure {
  std::vector<int>   c_f_row_index, c_f_row_ids;  /* coarse -> fine rows */
  std::vector<int>   c_row_index, c_col_id;       /* coarse MSR structure */
  std::vector<int>   c_cell_face;                 /* per coarse entry */
  std::vector<short> c_cell_face_sgn;
  int                c_n_faces = 0;
};

void build_coarse_structure(int f_n_rows, const int *f_row_index, const int *f_col_id,
                            const int *f_c_row, int c_n_rows, CoarseStructure &cs);

} // namespace csb

#endif
