/*============================================================================
 * grid_host.cpp -- host side of the multigrid hierarchy construction.
 *
 * The reference always aggregates on the host, even in its GPU builds
 * (src/alge/cs_grid.cpp:8035-8041, :8089-8097): the greedy DX aggregation is
 * an order-dependent sweep whose fine->coarse maps must be reproduced
 * bit-exactly, so it stays a host sweep here too.  The integer structure of
 * the coarse level (coarse->fine adjacency, sorted coarse columns, coarse
 * face numbering) is built here with OpenMP over coarse rows; all floating
 * point assembly of coarse values runs on the device (kernels.cu).
 *
 * Reference behaviour reproduced (not copied): _automatic_aggregation_dx_msr
 * (cs_grid.cpp:4282-4767), _coarse_to_fine_adjacency_msr (:2860),
 * _coarse_msr_struct (:3020), _msr_face_adjacency (:3751); pairwise aggregation
 * _pairwise_msr (:1744-2020) with its recursion for aggregation limits 4 and 8
 * (:8143-8150, :8305-8346) over host Galerkin matrices (:6040-6090).
 *============================================================================*/

#include "grid_host.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <omp.h>

namespace csb {

namespace {

/* Row range of block b out of nb, cache-line aligned like
   cs_parall_thread_range (src/base/cs_parall.h:571-590) with 8-byte items. */
inline void block_range(int n, int b, int nb, int &s, int &e)
{
  const int per = (n + nb - 1) / nb;
  const int cl = 64 / 8;
  auto up = [cl](long v) { return (int)(((v + cl - 1) / cl) * cl); };
  s = up((long)b * per);
  e = up((long)(b + 1) * per);
  if (e > n) e = n;
  if (s > n) s = n;
}

/* Greedy pairwise-growing aggregation of one independent block of rows. */
class BlockAggregator {
public:
  BlockAggregator(int level, const int *ri, const int *ci, const double *d,
                  const double *x, int max_aggregation, int s, int e,
                  int *f_c_row, int *queue_storage, short *card, short *count)
    : level_(level), ri_(ri), ci_(ci), d_(d), x_(x), limit_(max_aggregation),
      s_(s), e_(e), map_(f_c_row), card_(card), count_(count)
  {
    queue_ = queue_storage + (size_t)ri[s] * 2;
  }

  int run()
  {
    prepare();
    first_sweep();
    for (int pass = 2; pass <= 10; pass++) {
      int pending = 0;
      for (int i = s_; i < e_; i++)
        if (map_[i] < -1) pending++;
      if (pending == 0 || (long)(n_coarse_ + pending) * 8 < (e_ - s_) || q_len_ == 0)
        break;
      if (cur_limit_ < limit_) cur_limit_++;
      neighbour_bonus_ += 1;
      queue_sweep();
    }
    for (int i = s_; i < e_; i++)
      if (map_[i] < -1) map_[i] = n_coarse_++;
    return n_coarse_;
  }

private:
  void prepare()
  {
    for (int i = s_; i < e_; i++) {
      card_[i] = (short)(ri_[i+1] - ri_[i]);
      count_[i] = 1;
    }
    if (level_ == 0) {
      /* rows whose diagonal dwarfs the off-diagonal mass are left out
         ("penalised", threshold 1e4, cs_grid.cpp:165, :4409-4422) */
      for (int i = s_; i < e_; i++) {
        double mass = 0.0;
        for (int k = ri_[i]; k < ri_[i+1]; k++) {
          const double v = x_[k];
          if (v < 0.) mass -= v; else mass += v;
        }
        if (d_[i] > 1e4 * mass) {
          map_[i] = -1;
          count_[i] = 0;
        }
      }
    }
  }

  /* Returns true when edge (i, j) stored at entry k led to a merge. */
  inline bool try_edge(int i, int j, int k, int bonus)
  {
    const int ci = map_[i], cj = map_[j];
    const double inv = (double)((card_[i] + bonus) * (card_[j] + bonus))
                       / (d_[i] * d_[j]);
    const double sq = x_[k] * x_[k];     /* symmetric: a_ij * a_ji */
    const double crit = inv * sq;
    if (!(crit > 1. - 1.e-6)) return false;
    if (ci > -1 && cj < -1) {
      if (count_[s_ + ci] < cur_limit_ + 1) {
        map_[j] = ci; count_[s_ + ci] += 1; return true;
      }
    }
    else if (ci < -1 && cj > -1) {
      if (count_[s_ + cj] < cur_limit_ + 1) {
        map_[i] = cj; count_[s_ + cj] += 1; return true;
      }
    }
    else if (ci < -1 && cj < -1) {
      map_[i] = n_coarse_; map_[j] = n_coarse_;
      count_[s_ + n_coarse_] += 1;
      n_coarse_++;
      return true;
    }
    return false;
  }

  void first_sweep()
  {
    if (cur_limit_ < limit_) cur_limit_++;
    q_len_ = 0;
    for (int i = s_; i < e_; i++) {
      for (int k = ri_[i]; k < ri_[i+1]; k++) {
        const int j = ci_[k];
        if (j < i || j >= e_) continue;          /* upper triangle, inside block */
        const int ci = map_[i], cj = map_[j];
        if (ci == -1 || cj == -1 || (ci >= 0 && cj >= 0)) continue;
        if (!try_edge(i, j, k, 0)) {
          queue_[2*q_len_] = i; queue_[2*q_len_+1] = k;
          q_len_++;
        }
      }
    }
  }

  void queue_sweep()
  {
    long kept = 0;
    for (long q = 0; q < q_len_; q++) {
      const int i = queue_[2*q], k = queue_[2*q+1];
      const int j = ci_[k];
      if (map_[i] >= 0 && map_[j] >= 0) continue;
      if (!try_edge(i, j, k, neighbour_bonus_)) {
        queue_[2*kept] = i; queue_[2*kept+1] = k;
        kept++;
      }
    }
    q_len_ = kept;
  }

  int level_;
  const int *ri_, *ci_;
  const double *d_, *x_;
  int limit_, s_, e_;
  int *map_;
  short *card_, *count_;
  int *queue_ = nullptr;
  long q_len_ = 0;
  int n_coarse_ = 0;
  int cur_limit_ = 1;
  int neighbour_bonus_ = 0;
};

/* Renumber a block's coarse rows by first appearance (cs_grid.cpp:1638). */
void renumber_first_seen(int n, int n_coarse, int *map, int *scratch)
{
  std::fill(scratch, scratch + n_coarse, -1);
  int next = 0;
  for (int i = 0; i < n; i++) {
    const int c = map[i];
    if (c < 0) continue;
    if (scratch[c] < 0) scratch[c] = next++;
    map[i] = scratch[c];
  }
}

} // anonymous namespace

int aggregate_dx_msr(int level, int n_rows, const int *row_index, const int *col_id,
                     const double *d_val, const double *x_val, int max_aggregation,
                     int n_blocks, int *f_c_row)
{
  int nb = n_blocks;
  if (n_rows / 128 < nb) nb = n_rows / 128;      /* CS_THR_MIN, cs_parall.h:549 */
  if (nb < 1) nb = 1;

  const long nnz = row_index[n_rows];
  HostVec<short> card(std::max(n_rows, 1)), count(std::max(n_rows, 1));
  HostVec<int> queue(std::max<long>(std::max<long>(2*nnz, n_rows), 1));
  std::vector<int> n_coarse_b(nb, 0), rs(nb), re(nb);

  std::fill(f_c_row, f_c_row + n_rows, -2);

  #pragma omp parallel for schedule(static, 1) if (nb > 1)
  for (int b = 0; b < nb; b++) {
    block_range(n_rows, b, nb, rs[b], re[b]);
    BlockAggregator ag(level, row_index, col_id, d_val, x_val, max_aggregation,
                       rs[b], re[b], f_c_row, queue.data(), card.data(), count.data());
    n_coarse_b[b] = ag.run();
  }

  /* the queue storage is free again: reuse it as renumbering scratch */
  #pragma omp parallel for schedule(static, 1) if (nb > 1)
  for (int b = 0; b < nb; b++)
    renumber_first_seen(re[b] - rs[b], n_coarse_b[b], f_c_row + rs[b],
                        queue.data() + rs[b]);

  int total = 0;
  for (int b = 0; b < nb; b++) {
    const int shift = total;
    total += n_coarse_b[b];
    if (shift != 0)
      for (int i = rs[b]; i < re[b]; i++)
        if (f_c_row[i] > -1) f_c_row[i] += shift;
  }
  return total;
}

/*----------------------------------------------------------------------------
 * Pairwise aggregation (CS_GRID_COARSENING_SPD_PW)
 *----------------------------------------------------------------------------*/

namespace {

/* Rows bucketed by their number of still-free strong neighbours ("degree"); every bucket is
   a doubly linked list whose head is the most recently inserted row, so that the row taken
   next -- the head of the lowest non-empty bucket -- is the one the reference's
   _graph_m_ptr lists (cs_grid.cpp:1544-1620) would give. */
class DegreeBuckets {
public:
  DegreeBuckets(int n_rows, int max_degree)
    : head_(max_degree + 1, -1), next_(std::max(n_rows, 1), -1), prev_(std::max(n_rows, 1), -1),
      lo_(0), hi_(max_degree) {}

  int lo() const { return lo_; }
  int hi() const { return hi_; }
  int head(int d) const { return head_[d]; }
  void set_empty() { lo_ = -1; hi_ = -1; }
  void skip_empty_low() { while (lo_ < hi_ && head_[lo_] < 0) lo_++; }

  /* initial fill: called for rows in DEcreasing order, so that lists start ascending */
  void push_front_initial(int d, int row)
  {
    const int h = head_[d];
    head_[d] = row;
    next_[row] = h;
    if (h > -1) prev_[h] = row;
  }

  void remove(int d, int row)
  {
    const int p = prev_[row], n = next_[row];
    if (p > -1) {
      next_[p] = n;
      if (n > -1) prev_[n] = p;
    }
    else {
      head_[d] = n;
      if (n > -1) prev_[n] = -1;
      if (n < 0) {                     /* bucket d is now empty: tighten the bounds */
        if (d == lo_)
          skip_empty_low();
        else if (d == hi_) {
          hi_ -= 1;
          while (hi_ > lo_ && head_[hi_] < 0) hi_--;
        }
      }
    }
    prev_[row] = -1;
    next_[row] = -1;
  }

  void insert(int d, int row)
  {
    const int h = head_[d];
    if (h > -1) {
      prev_[h] = row;
      next_[row] = h;
    }
    else if (d < lo_)
      lo_ = d;
    else if (d > hi_) {
      hi_ = d;
      if (head_[lo_] < 0) lo_ = d;
    }
    head_[d] = row;
  }

private:
  std::vector<int> head_, next_, prev_;
  int lo_, hi_;
};

} // anonymous namespace

int aggregate_pw_msr(int level, int n_rows, const int *row_index, const int *col_id,
                     const double *d_val, const double *x_val, int *f_c_row)
{
  const double beta = 0.25;                                /* cs_grid.cpp:2049 */
  const double dd_threshold = (level == 0) ? 5. : -1.;     /* _dd_threshold_pw, :172, :2050 */

  std::vector<short> degree(std::max(n_rows, 1));
  std::vector<double> strongest(std::max(n_rows, 1));      /* max_k(-a_ik) over a_ik < 0 */
  std::fill(f_c_row, f_c_row + n_rows, -2);

  int max_degree = -1;
  for (int i = 0; i < n_rows; i++) {
    const int ks = row_index[i], ke = row_index[i+1];
    double abs_sum = 0., big = 0.;
    for (int k = ks; k < ke; k++) {
      const double v = x_val[k];
      abs_sum += std::fabs(v);
      if (v < 0) big = std::max(big, -v);
    }
    strongest[i] = big;
    short deg = 0;
    if (dd_threshold > 0 && d_val[i] > dd_threshold * abs_sum) {
      deg = -1;                       /* strongly diagonally dominant: left out */
      f_c_row[i] = -1;
    }
    else if (dd_threshold > 0) {
      for (int k = ke - 1; k >= ks; k--)
        if (col_id[k] < n_rows && x_val[k] < beta * abs_sum) deg++;
    }
    else {
      for (int k = ke - 1; k >= ks; k--)
        if (col_id[k] < n_rows) deg++;
    }
    degree[i] = deg;
    if (max_degree < deg) max_degree = deg;
  }
  if (max_degree < 0) return 0;       /* every row left out (cs_grid.cpp:1850) */

  DegreeBuckets buckets(n_rows, max_degree);
  for (int i = n_rows - 1; i >= 0; i--)
    if (degree[i] >= 0) buckets.push_front_initial(degree[i], i);
  buckets.skip_empty_low();

  int n_coarse = 0;
  while (buckets.lo() > -1) {
    const int i = buckets.head(buckets.lo());
    int pair[2] = {i, -1};
    f_c_row[i] = n_coarse;
    /* free neighbour with the most negative coefficient, kept if strong enough */
    int j = -1;
    double a_min = HUGE_VAL;
    for (int k = row_index[i]; k < row_index[i+1]; k++) {
      const int c = col_id[k];
      if (c < n_rows && f_c_row[c] == -2 && x_val[k] < a_min) { a_min = x_val[k]; j = c; }
    }
    if (row_index[i+1] > row_index[i] && !(a_min >= -beta * strongest[i])) {
      f_c_row[j] = n_coarse;
      pair[1] = j;
    }
    n_coarse++;

    for (int ip = 0; ip < 2; ip++) {
      const int r = pair[ip];
      if (r < 0 || degree[r] < 0) continue;
      buckets.remove(degree[r], r);
      degree[r] = -1;
      for (int k = row_index[r+1] - 1; k >= row_index[r]; k--) {
        const int c = col_id[k];
        if (c >= n_rows) continue;
        const short dc = degree[c];
        if (dc >= 0) {
          buckets.remove(dc, c);
          if (dc > 0) buckets.insert(dc - 1, c);
          degree[c] = dc - 1;
        }
      }
    }
    if (buckets.lo() >= buckets.hi() && buckets.head(buckets.lo()) < 0)
      buckets.set_empty();
  }

  for (int i = 0; i < n_rows; i++)
    if (f_c_row[i] < -1) f_c_row[i] = n_coarse++;
  return n_coarse;
}

/* Galerkin coarse values on the host, same sums in the same order as k_coarse_galerkin
   (kernels.cu) and the reference's row-parallel lambda (cs_grid.cpp:6040-6090). */
void galerkin_values_host(int f_n_rows, const int *f_row_index, const int *f_col_id,
                          const double *f_d_val, const double *f_x_val, const int *f_c_row,
                          int c_n_rows, const CoarseStructure &cs, HostVec<double> &c_d_val,
                          HostVec<double> &c_x_val)
{
  (void)f_n_rows;
  c_d_val.assign(std::max(c_n_rows, 1), 0.);
  c_x_val.assign(std::max(cs.c_row_index[c_n_rows], 1), 0.);
  #pragma omp parallel for schedule(static)
  for (int ic = 0; ic < c_n_rows; ic++) {
    const int s_id = cs.c_row_index[ic];
    const int *cols = cs.c_col_id.data() + s_id;
    const int n_cols = cs.c_row_index[ic+1] - s_id;
    double *row = c_x_val.data() + s_id;
    double d = 0.;
    for (int r = cs.c_f_row_index[ic]; r < cs.c_f_row_index[ic+1]; r++) {
      const int ii = cs.c_f_row_ids[r];
      d += f_d_val[ii];
      for (int k = f_row_index[ii]; k < f_row_index[ii+1]; k++) {
        const int jc = f_c_row[f_col_id[k]];
        if (jc < 0) continue;
        if (jc != ic)
          row[std::lower_bound(cols, cols + n_cols, jc) - cols] += f_x_val[k];
        else
          d += f_x_val[k];
      }
    }
    c_d_val[ic] = d;
  }
}

int coarsen_pairwise_galerkin(int level, int n_rows, const int *row_index, const int *col_id,
                              const double *d_val, const double *x_val, int aggregation_limit,
                              int *f_c_row, HostMsr &coarse)
{
  const int c_n_rows = aggregate_pw_msr(level, n_rows, row_index, col_id, d_val, x_val, f_c_row);
  CoarseStructure cs;
  build_coarse_structure(n_rows, row_index, col_id, f_c_row, c_n_rows, cs);
  HostMsr c;
  c.n_rows = c_n_rows;
  galerkin_values_host(n_rows, row_index, col_id, d_val, x_val, f_c_row, c_n_rows, cs,
                       c.d_val, c.x_val);
  c.row_index.swap(cs.c_row_index);
  c.col_id.swap(cs.c_col_id);
  if (aggregation_limit > 2) {
    /* pair the pairs (limit / 2), project the map back to the fine rows and keep only the
       coarsest matrix, whose values are sums of the intermediate matrix's */
    std::vector<int> c_cc(std::max(c_n_rows, 1));
    HostMsr cc;
    const int cc_n_rows
      = coarsen_pairwise_galerkin(level + 1, c_n_rows, c.row_index.data(), c.col_id.data(),
                                  c.d_val.data(), c.x_val.data(), aggregation_limit / 2,
                                  c_cc.data(), cc);
    for (int i = 0; i < n_rows; i++)
      if (f_c_row[i] >= 0) f_c_row[i] = c_cc[f_c_row[i]];
    coarse = std::move(cc);
    return cc_n_rows;
  }
  coarse = std::move(c);
  return c_n_rows;
}

void build_coarse_structure(int f_n_rows, const int *f_row_index, const int *f_col_id,
                            const int *f_c_row, int c_n_rows, CoarseStructure &cs)
{
  const bool tm = (getenv("CSB_SETUP_TIMES") != nullptr);
  double t_[8]; int ti_ = 0;
  auto mark = [&]() { if (tm) t_[ti_++] = omp_get_wtime(); };
  mark();
  /* coarse -> fine rows, ascending fine ids inside each coarse row */
  cs.c_f_row_index.assign((size_t)c_n_rows + 1, 0);
  cs.c_f_row_ids.resize(std::max(f_n_rows, 1));
  for (int i = 0; i < f_n_rows; i++)
    if (f_c_row[i] > -1) cs.c_f_row_index[f_c_row[i] + 1]++;
  for (int c = 0; c < c_n_rows; c++)
    cs.c_f_row_index[c+1] += cs.c_f_row_index[c];
  {
    std::vector<int> fill(cs.c_f_row_index.begin(), cs.c_f_row_index.end() - 1);
    for (int i = 0; i < f_n_rows; i++) {
      const int c = f_c_row[i];
      if (c > -1) cs.c_f_row_ids[fill[c]++] = i;
    }
  }

  mark();
  /* sorted, unique coarse columns per coarse row, in one pass: every coarse row first
     writes its columns into an oversized array (room for all entries of its fine rows, the
     reference's c_row_index_0, cs_grid.cpp:2966-2980), sorts and deduplicates them there;
     a scan of the counts and a parallel copy compact the rows */
  cs.c_row_index.assign((size_t)c_n_rows + 1, 0);
  const int *cfi = cs.c_f_row_index.data(), *cfr = cs.c_f_row_ids.data();
  std::vector<long> ub((size_t)c_n_rows + 1, 0);
  #pragma omp parallel for schedule(static)
  for (int c = 0; c < c_n_rows; c++) {
    long n = 0;
    for (int r = cfi[c]; r < cfi[c+1]; r++) n += f_row_index[cfr[r]+1] - f_row_index[cfr[r]];
    ub[(size_t)c+1] = n;
  }
  for (int c = 0; c < c_n_rows; c++) ub[(size_t)c+1] += ub[c];
  mark();
  HostVec<int> wide((size_t)std::max<long>(ub[c_n_rows], 1));
  #pragma omp parallel for schedule(static)
  for (int c = 0; c < c_n_rows; c++) {
    int *buf = wide.data() + ub[c];
    int n = 0;
    for (int r = cfi[c]; r < cfi[c+1]; r++) {
      const int i = cfr[r];
      for (int k = f_row_index[i]; k < f_row_index[i+1]; k++) {
        const int cc = f_c_row[f_col_id[k]];
        if (cc > -1 && cc != c) buf[n++] = cc;
      }
    }
    std::sort(buf, buf + n);
    cs.c_row_index[(size_t)c+1] = (int)(std::unique(buf, buf + n) - buf);
  }
  mark();
  for (int c = 0; c < c_n_rows; c++)
    cs.c_row_index[c+1] += cs.c_row_index[c];
  const int c_nnz = cs.c_row_index[c_n_rows];
  cs.c_col_id.resize(std::max(c_nnz, 1));
  #pragma omp parallel for schedule(static)
  for (int c = 0; c < c_n_rows; c++)
    std::copy(wide.data() + ub[c], wide.data() + ub[c] + (cs.c_row_index[c+1] - cs.c_row_index[c]),
              cs.c_col_id.data() + cs.c_row_index[c]);
  { HostVec<int>().swap(wide); }
  mark();

  /* coarse faces = upper-triangle entries numbered in (row, column) order;
     lower-triangle entries point to the same face with sign -1 */
  cs.c_cell_face.resize(std::max(c_nnz, 1));
  cs.c_cell_face_sgn.resize(std::max(c_nnz, 1));
  std::vector<int> upper_before((size_t)c_n_rows + 1, 0);
  #pragma omp parallel for schedule(static)
  for (int c = 0; c < c_n_rows; c++) {
    int n_up = 0;
    for (int k = cs.c_row_index[c]; k < cs.c_row_index[c+1]; k++)
      if (cs.c_col_id[k] >= c) n_up++;
    upper_before[c+1] = n_up;
  }
  for (int c = 0; c < c_n_rows; c++)
    upper_before[c+1] += upper_before[c];
  cs.c_n_faces = upper_before[c_n_rows];
  #pragma omp parallel for schedule(static)
  for (int c = 0; c < c_n_rows; c++) {
    int f = upper_before[c];
    for (int k = cs.c_row_index[c]; k < cs.c_row_index[c+1]; k++) {
      if (cs.c_col_id[k] < c) { cs.c_cell_face[k] = -1; cs.c_cell_face_sgn[k] = -1; }
      else { cs.c_cell_face[k] = f++; cs.c_cell_face_sgn[k] = 1; }
    }
  }
  #pragma omp parallel for schedule(static)
  for (int c = 0; c < c_n_rows; c++) {
    for (int k = cs.c_row_index[c]; k < cs.c_row_index[c+1]; k++) {
      if (cs.c_cell_face_sgn[k] > 0) continue;
      const int a = cs.c_col_id[k];
      const int *b = cs.c_col_id.data() + cs.c_row_index[a];
      const int *e = cs.c_col_id.data() + cs.c_row_index[a+1];
      const int *p = std::lower_bound(b, e, c);
      cs.c_cell_face[k] = cs.c_cell_face[p - cs.c_col_id.data()];
    }
  }
  mark();
  if (tm) {
    fprintf(stderr, "[csb200] coarse structure %d -> %d rows:", f_n_rows, c_n_rows);
    for (int i = 1; i < ti_; i++) fprintf(stderr, " %.3f", t_[i] - t_[i-1]);
    fprintf(stderr, " s (c->f adjacency, bounds, gather+sort, compact, faces)\n");
  }
}

/*----------------------------------------------------------------------------
 * Matrices mapped from native (face-based) arrays: the entry of a caller that holds
 * edges, da, xa (cs_matrix_set_coefficients, cs_matrix.cpp:5315).  The reference then
 * aggregates over the face list when it runs one thread (_automatic_aggregation_fc,
 * cs_grid.cpp:3997-4263), numbers the coarse faces in the order the fine faces reach them
 * (_coarsen_face_cell, :637-788) and sums the coarse values in face order
 * (_compute_coarse_quantities_native, :5144-5558; on the device here, kernels.cu).
 *----------------------------------------------------------------------------*/

/* MSR structure of a face list: both directions of every face, columns ascending in each
   row, and for every entry its face and side (+1: the row is the face's first cell).
   What cs_matrix_structure_create(CS_MATRIX_MSR) + _matrix_from_native's entry -> face
   map give (cs_grid.cpp:7146-7186).  Returns false when two faces join the same cells. */
bool msr_structure_from_faces(int n_rows, int n_faces, const int *face_cell,
                              std::vector<int> &row_index, HostVec<int> &col_id,
                              HostVec<int> &c2f, HostVec<short> &c2f_sgn)
{
  row_index.assign((size_t)n_rows + 1, 0);
  for (int f = 0; f < n_faces; f++) {
    const int a = face_cell[2*f], b = face_cell[2*f+1];
    if (a < n_rows) row_index[(size_t)a+1]++;
    if (b < n_rows) row_index[(size_t)b+1]++;
  }
  for (int i = 0; i < n_rows; i++) row_index[(size_t)i+1] += row_index[i];
  const int nnz = row_index[n_rows];
  col_id.resize(std::max(nnz, 1));
  c2f.resize(std::max(nnz, 1));
  c2f_sgn.resize(std::max(nnz, 1));
  /* (column, signed face) pairs are dropped in face order, then every row is sorted by
     column; the face rides along in the upper half of a 64-bit key */
  HostVec<int64_t> key((size_t)std::max(nnz, 1));
  {
    std::vector<int> fill(row_index.begin(), row_index.end() - 1);
    for (int f = 0; f < n_faces; f++) {
      const int a = face_cell[2*f], b = face_cell[2*f+1];
      if (a < n_rows) key[(size_t)fill[a]++] = ((int64_t)b << 32) | (uint32_t)(2*f);
      if (b < n_rows) key[(size_t)fill[b]++] = ((int64_t)a << 32) | (uint32_t)(2*f + 1);
    }
  }
  bool unique = true;
  #pragma omp parallel for schedule(static) reduction(&&:unique)
  for (int i = 0; i < n_rows; i++) {
    int64_t *b = key.data() + row_index[i], *e = key.data() + row_index[(size_t)i+1];
    std::sort(b, e);
    for (int64_t *p = b; p < e; p++) {
      const size_t k = (size_t)(p - key.data());
      col_id[k] = (int)(*p >> 32);
      const uint32_t fs = (uint32_t)(*p & 0xffffffff);
      c2f[k] = (int)(fs >> 1);
      c2f_sgn[k] = (fs & 1u) ? (short)-1 : (short)1;
      if (p > b && col_id[k] == col_id[k-1]) unique = false;
    }
  }
  return unique;
}

/* Face-list form of the DX aggregation.  The faces still worth trying are kept in one
   queue compacted in place behind the read position (a face is re-queued at or before the
   slot it was read from).  f_c_cell: n_cols_ext entries, all written. */
int aggregate_fc(int level, int n_rows, int n_cols_ext, int n_faces, const int *face_cell,
                 const double *da, const double *xa, int max_aggregation, int *f_c_cell)
{
  const int n_ext = std::max(n_cols_ext, 1);
  HostVec<int> queue((size_t)std::max(n_faces, 1)), n_nei((size_t)n_ext), n_in((size_t)n_ext);
  for (int i = 0; i < n_cols_ext; i++) { f_c_cell[i] = -2; n_nei[i] = -1; n_in[i] = 1; }
  for (int f = 0; f < n_faces; f++) {
    queue[f] = f;
    n_nei[face_cell[2*f]] += 1;
    n_nei[face_cell[2*f+1]] += 1;
  }
  if (level == 0) {
    /* cells whose diagonal dwarfs their extra-diagonal mass stay out (threshold 1e4) */
    HostVec<double> mass((size_t)n_ext);
    for (int i = 0; i < n_cols_ext; i++) mass[i] = 0.;
    for (int f = 0; f < n_faces; f++) {
      const double v = std::fabs(xa[f]);
      mass[face_cell[2*f]] += v;
      mass[face_cell[2*f+1]] += v;
    }
    for (int i = 0; i < n_rows; i++)
      if (da[i] > 1e4 * mass[i]) f_c_cell[i] = -1;
  }

  int n_coarse = 0, q_len = n_faces, limit = 1;
  for (int pass = 1; pass <= 10; pass++) {
    limit = std::min(limit + 1, max_aggregation);
    const int bonus = pass;
    const int n_try = q_len;
    q_len = 0;
    for (int q = 0; q < n_try; q++) {
      const int f = queue[q];
      const int i = face_cell[2*f], j = face_cell[2*f+1];
      if (i >= n_rows || j >= n_rows) continue;          /* never across a rank boundary */
      const int ci = f_c_cell[i], cj = f_c_cell[j];
      if (!(ci < -1 || cj < -1)) continue;
      bool merged = false;
      const double inv = (double)((n_nei[i] + bonus) * (n_nei[j] + bonus)) / (da[i] * da[j]);
      const double crit = inv * (xa[f] * xa[f]);
      if (crit > 1. - 1.e-6) {
        if (ci > -1 && cj < -1) {
          if (n_in[ci] < limit + 1) { f_c_cell[j] = ci; n_in[ci] += 1; merged = true; }
        }
        else if (ci < -1 && cj > -1) {
          if (n_in[cj] < limit + 1) { f_c_cell[i] = cj; n_in[cj] += 1; merged = true; }
        }
        else if (ci < -1 && cj < -1) {
          f_c_cell[i] = n_coarse; f_c_cell[j] = n_coarse;
          n_in[n_coarse] += 1;
          n_coarse++;
          merged = true;
        }
      }
      if (!merged && (f_c_cell[i] < -1 || f_c_cell[j] < -1)) queue[q_len++] = f;
    }
    long pending = 0;
    #pragma omp parallel for reduction(+:pending) schedule(static)
    for (int i = 0; i < n_rows; i++)
      if (f_c_cell[i] < -1) pending++;
    if (pending == 0 || ((long)n_coarse + pending) * 8 < n_rows || q_len == 0) break;
  }
  for (int i = 0; i < n_rows; i++)
    if (f_c_cell[i] < -1) f_c_cell[i] = n_coarse++;
  return n_coarse;
}

/* Coarse faces of the native path: the coarse MSR structure of `cs` (sorted coarse
   neighbours, build_coarse_structure) is kept, its faces are renumbered in the order in
   which the fine faces first reach them; c_face_cell[2*cf] < c_face_cell[2*cf+1]. */
void number_coarse_faces_native(int f_n_faces, const int *f_face_cell, const int *f_c_row,
                                int c_n_rows, CoarseStructure &cs, HostVec<int> &c_face_cell)
{
  const int c_nnz = cs.c_row_index[c_n_rows];
  int *face = cs.c_cell_face.data();
  short *sgn = cs.c_cell_face_sgn.data();
  #pragma omp parallel for schedule(static)
  for (int k = 0; k < c_nnz; k++) face[k] = -1;
  c_face_cell.resize((size_t)std::max(c_nnz, 2));     /* 2 x (nnz / 2) faces */
  int n = 0;
  auto entry = [&cs](int row, int col) {
    const int *b = cs.c_col_id.data() + cs.c_row_index[row];
    const int *e = cs.c_col_id.data() + cs.c_row_index[(size_t)row+1];
    return (int)(std::lower_bound(b, e, col) - cs.c_col_id.data());
  };
  for (int f = 0; f < f_n_faces; f++) {
    int ic = f_c_row[f_face_cell[2*f]], jc = f_c_row[f_face_cell[2*f+1]];
    if (ic == jc || ic < 0 || jc < 0 || ic >= c_n_rows || jc >= c_n_rows) continue;
    if (ic > jc) std::swap(ic, jc);
    const int k = entry(ic, jc);
    if (face[k] > -1) continue;
    const int k2 = entry(jc, ic);
    face[k] = n; sgn[k] = 1;
    face[k2] = n; sgn[k2] = -1;
    c_face_cell[2*(size_t)n] = ic;
    c_face_cell[2*(size_t)n+1] = jc;
    n++;
  }
  cs.c_n_faces = n;
}

} // namespace csb
