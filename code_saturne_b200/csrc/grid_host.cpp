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
 * _coarse_msr_struct (:3020), _msr_face_adjacency (:3751).
 *============================================================================*/

#include "grid_host.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
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
  std::vector<short> card(std::max(n_rows, 1)), count(std::max(n_rows, 1));
  std::vector<int> queue(std::max<long>(std::max<long>(2*nnz, n_rows), 1));
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

void build_coarse_structure(int f_n_rows, const int *f_row_index, const int *f_col_id,
                            const int *f_c_row, int c_n_rows, CoarseStructure &cs)
{
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

  /* sorted, unique coarse columns per coarse row: count, scan, fill */
  cs.c_row_index.assign((size_t)c_n_rows + 1, 0);
  const int *cfi = cs.c_f_row_index.data(), *cfr = cs.c_f_row_ids.data();

  auto gather_row = [&](int c, std::vector<int> &buf) {
    buf.clear();
    for (int r = cfi[c]; r < cfi[c+1]; r++) {
      const int i = cfr[r];
      for (int k = f_row_index[i]; k < f_row_index[i+1]; k++) {
        const int cc = f_c_row[f_col_id[k]];
        if (cc > -1 && cc != c) buf.push_back(cc);
      }
    }
    std::sort(buf.begin(), buf.end());
    buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
  };

  #pragma omp parallel
  {
    std::vector<int> buf;
    buf.reserve(64);
    #pragma omp for schedule(static)
    for (int c = 0; c < c_n_rows; c++) {
      gather_row(c, buf);
      cs.c_row_index[c+1] = (int)buf.size();
    }
  }
  for (int c = 0; c < c_n_rows; c++)
    cs.c_row_index[c+1] += cs.c_row_index[c];
  const int c_nnz = cs.c_row_index[c_n_rows];
  cs.c_col_id.resize(std::max(c_nnz, 1));
  #pragma omp parallel
  {
    std::vector<int> buf;
    buf.reserve(64);
    #pragma omp for schedule(static)
    for (int c = 0; c < c_n_rows; c++) {
      gather_row(c, buf);
      std::copy(buf.begin(), buf.end(), cs.c_col_id.begin() + cs.c_row_index[c]);
    }
  }

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
}

} // namespace csb
