/* grid_host.h -- host-side hierarchy construction helpers (see grid_host.cpp). */
#ifndef CSB_GRID_HOST_H
#define CSB_GRID_HOST_H

#include <cstddef>
#include <memory>
#include <utility>
#include <vector>

namespace csb {

/* std::vector whose resize() leaves new elements uninitialised: the structure arrays of a
   16 M-row level are hundreds of megabytes, every element is written before it is read, and
   a zero fill by one thread costs as much as the sweep that follows.
   Storage comes from the library's host pool (runtime.cu): page-locked blocks that are
   cached when a hierarchy is freed and handed out again by the next setup -- the reference
   rebuilds the hierarchy at every pressure solve, and a block that is already pinned and
   already faulted in moves to and from the device at the full PCIe rate. */
extern "C" void *csb_host_pool_alloc(size_t bytes);
extern "C" void  csb_host_pool_free(void *p, size_t bytes);

template <typename T>
struct NoInitAlloc {
  using value_type = T;
  template <typename U> struct rebind { using other = NoInitAlloc<U>; };
  NoInitAlloc() = default;
  template <typename U> NoInitAlloc(const NoInitAlloc<U> &) {}
  T *allocate(size_t n) { return static_cast<T *>(csb_host_pool_alloc(n*sizeof(T))); }
  void deallocate(T *p, size_t n) { csb_host_pool_free(p, n*sizeof(T)); }
  template <typename U> bool operator==(const NoInitAlloc<U> &) const { return true; }
  template <typename U> bool operator!=(const NoInitAlloc<U> &) const { return false; }
  template <typename U, typename... A> void construct(U *p, A &&... a)
  {
    if constexpr (sizeof...(A) == 0) ::new ((void *)p) U;
    else ::new ((void *)p) U(std::forward<A>(a)...);
  }
};
template <typename T> using HostVec = std::vector<T, NoInitAlloc<T>>;

/* Fine -> coarse row map of the reference's DX aggregation for a symmetric
   scalar MSR matrix (-2 never remains; -1 = penalised row).  n_blocks > 1
   confines aggregates to independent row blocks, as the reference does per
   OpenMP thread.  Returns the number of coarse rows. */
int aggregate_dx_msr(int level, int n_rows, const int *row_index, const int *col_id,
                     const double *d_val, const double *x_val, int max_aggregation,
                     int n_blocks, int *f_c_row);

struct CoarseStructure {
  std::vector<int>   c_f_row_index;               /* coarse -> fine rows */
  HostVec<int>       c_f_row_ids;
  std::vector<int>   c_row_index;                 /* coarse MSR structure */
  HostVec<int>       c_col_id;
  HostVec<int>       c_cell_face;                 /* per coarse entry */
  HostVec<short>     c_cell_face_sgn;
  int                c_n_faces = 0;
};

void build_coarse_structure(int f_n_rows, const int *f_row_index, const int *f_col_id,
                            const int *f_c_row, int c_n_rows, CoarseStructure &cs);

/* Matrices mapped from native (face-based) arrays (grid_host.cpp, last section) */
bool msr_structure_from_faces(int n_rows, int n_faces, const int *face_cell,
                              std::vector<int> &row_index, HostVec<int> &col_id,
                              HostVec<int> &c2f, HostVec<short> &c2f_sgn);
int aggregate_fc(int level, int n_rows, int n_cols_ext, int n_faces, const int *face_cell,
                 const double *da, const double *xa, int max_aggregation, int *f_c_cell);
void number_coarse_faces_native(int f_n_faces, const int *f_face_cell, const int *f_c_row,
                                int c_n_rows, CoarseStructure &cs, HostVec<int> &c_face_cell);

/* Fine -> coarse row map of the reference's pairwise aggregation (_pairwise_msr,
   cs_grid.cpp:1744-2020; beta = 0.25, diagonal-dominance threshold 5 at level 0).
   Returns the number of coarse rows. */
int aggregate_pw_msr(int level, int n_rows, const int *row_index, const int *col_id,
                     const double *d_val, const double *x_val, int *f_c_row);

struct HostMsr {
  int                 n_rows = 0;
  std::vector<int>    row_index;
  HostVec<int>        col_id;
  HostVec<double>     d_val, x_val;
};

/* One coarsening step of CS_GRID_COARSENING_SPD_PW with Galerkin values: pairs, and for an
   aggregation limit of 4 or 8 pairs of pairs (recursion on the intermediate Galerkin matrix,
   cs_grid.cpp:8305-8346).  f_c_row receives the map from the fine rows to the rows of
   `coarse`.  Returns the number of coarse rows. */
int coarsen_pairwise_galerkin(int level, int n_rows, const int *row_index, const int *col_id,
                              const double *d_val, const double *x_val, int aggregation_limit,
                              int *f_c_row, HostMsr &coarse);

} // namespace csb

#endif
