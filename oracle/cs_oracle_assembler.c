/*============================================================================
 * oracle/cs_oracle_assembler.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the reference's matrix assembler on one rank, MSR target
 * (separate diagonal), scalar coefficients:
 *
 *   structure  cs_matrix_assembler_add_g_ids / _compute     cs_matrix_assembler.cpp:2673, :2734
 *              (_matrix_assembler_compute_local :1780, _sort_and_compact_local :230):
 *              per row the sorted, duplicate-free list of its columns, diagonal apart
 *   values     cs_matrix_assembler_values_init / _add_g     cs_matrix.cpp:5635, :2346;
 *              cs_matrix_assembler.cpp:3276 -> _msr_assembler_values_add cs_matrix.cpp:2422:
 *              zero, then `+=` of every coefficient in arrival order (one thread)
 *
 * Pinned bit for bit against oracle/_ref/libcsref.so (tests/test_oracle.py,
 * tests/golden/make_golden_assembler.py).
 *============================================================================*/

#include <stdlib.h>
#include <string.h>

#include "cs_oracle.h"

static int _cmp_i(const void *a, const void *b)
{ return (*(const int *)a > *(const int *)b) - (*(const int *)a < *(const int *)b); }

/* row_index: n_rows + 1 entries; col_id: at least n entries (the caller sizes it for the
   worst case); returns the number of extra-diagonal entries */
int
ora_assembler_structure(int n_rows, int n, const long long *row_g_id, const long long *col_g_id,
                        long long l_range_0, int *row_index, int *col_id)
{
  int *count = calloc(n_rows + 1, sizeof(int));
  for (int i = 0; i < n; i++)
    if (row_g_id[i] != col_g_id[i]) count[row_g_id[i] - l_range_0 + 1]++;
  for (int r = 0; r < n_rows; r++) count[r+1] += count[r];
  int *wide = malloc(sizeof(int)*(count[n_rows] > 0 ? count[n_rows] : 1));
  int *fill = malloc(sizeof(int)*(n_rows > 0 ? n_rows : 1));
  for (int r = 0; r < n_rows; r++) fill[r] = count[r];
  for (int i = 0; i < n; i++)
    if (row_g_id[i] != col_g_id[i])
      wide[fill[row_g_id[i] - l_range_0]++] = (int)(col_g_id[i] - l_range_0);
  int nnz = 0;
  row_index[0] = 0;
  for (int r = 0; r < n_rows; r++) {
    int *b = wide + count[r];
    const int m = count[r+1] - count[r];
    qsort(b, m, sizeof(int), _cmp_i);
    for (int k = 0; k < m; k++)
      if (k == 0 || b[k] != b[k-1]) col_id[nnz++] = b[k];
    row_index[r+1] = nnz;
  }
  free(count); free(wide); free(fill);
  return nnz;
}

/* add n coefficients to (d_val, x_val), which the caller zeroes first (values_init);
   returns the number of coefficients that have no entry in the structure */
int
ora_assembler_values_add(int n_rows, const int *row_index, const int *col_id, int n,
                         const long long *row_g_id, const long long *col_g_id,
                         long long l_range_0, const double *val, double *d_val, double *x_val)
{
  int n_missing = 0;
  for (int i = 0; i < n; i++) {
    const int r = (int)(row_g_id[i] - l_range_0), c = (int)(col_g_id[i] - l_range_0);
    if (r == c) { d_val[r] += val[i]; continue; }
    int lo = row_index[r], hi = row_index[r+1] - 1, k = -1;
    while (lo <= hi) {
      const int mid = lo + (hi - lo)/2;
      if (col_id[mid] < c) lo = mid + 1;
      else if (col_id[mid] > c) hi = mid - 1;
      else { k = mid; break; }
    }
    if (k < 0) n_missing++;
    else x_val[k] += val[i];
  }
  return n_missing;
}
