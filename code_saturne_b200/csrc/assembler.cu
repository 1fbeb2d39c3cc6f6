/*============================================================================
 * assembler.cu -- matrix assembler (cs_matrix_assembler_t / cs_matrix_assembler_values_t,
 * src/alge/cs_matrix_assembler.cpp; matrix side src/alge/cs_matrix.cpp:2346-2560, :5635)
 * for matrices whose graph is given as (row, column) pairs of global ids rather than by
 * mesh faces: CDO systems, internal coupling.  One rank (every id inside l_range).
 *
 * What the reference defines: the structure is the sorted, duplicate-free column list of
 * every row with the diagonal kept apart (MSR); the values start at zero and every
 * coefficient given to cs_matrix_assembler_values_add(_g) is ADDED to its entry, in the
 * order of the calls (one thread), so an entry is a sum in arrival order.
 *
 * Here nothing is added one coefficient at a time.  The coefficients that arrive are given a
 * slot (row's diagonal or position of the extra-diagonal entry) by a parallel binary search
 * and parked; when the values are finalized a stable bucket pass lists, for every slot, its
 * contributions in arrival order and ONE kernel folds them (k_assembler_fold): same sums as
 * the reference, bit for bit, without atomics.  For the usual time loop -- the same pairs in
 * the same order at every step -- the list is made once from the pairs declared with
 * csb_matrix_assembler_add_g_ids and kept on the device (csb_matrix_assembler_values_add_declared):
 * a re-assembly is then that one kernel reading the caller's value array, which may live in
 * HBM.
 *============================================================================*/

#include "csb_internal.h"

#include <algorithm>
#include <omp.h>

using namespace csb;

struct _csb_matrix_assembler_t {
  long long  l_range[2] = {0, 0};
  bool       separate_diag = true;
  bool       computed = false;
  std::vector<long long> g_rc;          /* declared pairs, in order: row, column */
  csb_lnum_t n_rows = 0;
  std::vector<csb_lnum_t> r_idx;        /* MSR structure: sorted unique columns, no diagonal */
  HostVec<csb_lnum_t>     c_id;
  /* the declared sequence as a fold plan (made on first use) */
  bool       plan_ready = false;
  DevBuf<csb_lnum_t> plan_index, plan_src;
};

struct _csb_matrix_assembler_values_t {
  csb_matrix_t                  *m = nullptr;
  const _csb_matrix_assembler_t *ma = nullptr;
  /* parked coefficients: slot (row: diagonal of that row; n_rows + k: extra-diagonal entry
     k) and value, in arrival order */
  HostVec<csb_lnum_t> slot;
  HostVec<double>     val;
  size_t              n_parked = 0;
};

namespace {

/* slot of (row, column), -1 when the pair is not in the structure */
inline csb_lnum_t slot_of(const _csb_matrix_assembler_t *ma, long long g_r, long long g_c)
{
  if (g_r < ma->l_range[0] || g_r >= ma->l_range[1] || g_c < ma->l_range[0]
      || g_c >= ma->l_range[1]) return -1;
  const csb_lnum_t r = (csb_lnum_t)(g_r - ma->l_range[0]), c = (csb_lnum_t)(g_c - ma->l_range[0]);
  if (r == c) return r;
  const csb_lnum_t *b = ma->c_id.data() + ma->r_idx[r], *e = ma->c_id.data() + ma->r_idx[(size_t)r+1];
  const csb_lnum_t *p = std::lower_bound(b, e, c);
  if (p == e || *p != c) return -1;
  return ma->n_rows + (csb_lnum_t)(p - ma->c_id.data());
}

/* contributions of every slot in arrival order: index[n_slots + 1], src[n] */
void bucket_by_slot(size_t n, const csb_lnum_t *slot, size_t n_slots,
                    HostVec<csb_lnum_t> &index, HostVec<csb_lnum_t> &src)
{
  index.resize(n_slots + 1);
  src.resize(std::max<size_t>(n, 1));
  for (size_t s = 0; s <= n_slots; s++) index[s] = 0;
  for (size_t i = 0; i < n; i++) index[(size_t)slot[i] + 1]++;
  for (size_t s = 0; s < n_slots; s++) index[s+1] += index[s];
  HostVec<csb_lnum_t> fill(index.begin(), index.end() - 1);
  for (size_t i = 0; i < n; i++) src[(size_t)fill[slot[i]]++] = (csb_lnum_t)i;
}

} // anonymous namespace

extern "C" {

csb_matrix_assembler_t *
csb_matrix_assembler_create(const csb_gnum_t l_range[2], int separate_diag)
{
  if (!ctx().ready) { set_error("csb_matrix_assembler_create: csb_init not called"); return nullptr; }
  if (l_range == nullptr || l_range[1] < l_range[0]
      || l_range[1] - l_range[0] > 0x7fffffffLL) {
    set_error("csb_matrix_assembler_create: invalid row range"); return nullptr;
  }
  if (!separate_diag) {
    set_error("csb_matrix_assembler_create: the matrices of this library are MSR: "
              "separate_diag must be true");
    return nullptr;
  }
  csb_matrix_assembler_t *ma = new csb_matrix_assembler_t;
  ma->l_range[0] = l_range[0]; ma->l_range[1] = l_range[1];
  ma->n_rows = (csb_lnum_t)(l_range[1] - l_range[0]);
  return ma;
}

void csb_matrix_assembler_destroy(csb_matrix_assembler_t **ma)
{
  if (ma != nullptr && *ma != nullptr) { delete *ma; *ma = nullptr; }
}

int csb_matrix_assembler_add_g_ids(csb_matrix_assembler_t *ma, csb_lnum_t n,
                                   const csb_gnum_t *row_g_id, const csb_gnum_t *col_g_id)
{
  if (ma == nullptr || n < 0 || (n > 0 && (row_g_id == nullptr || col_g_id == nullptr))) {
    set_error("csb_matrix_assembler_add_g_ids: bad arguments"); return -1;
  }
  if (ma->computed) {
    set_error("csb_matrix_assembler_add_g_ids: structure already computed"); return -1;
  }
  for (csb_lnum_t i = 0; i < n; i++) {
    if (row_g_id[i] < ma->l_range[0] || row_g_id[i] >= ma->l_range[1]
        || col_g_id[i] < ma->l_range[0] || col_g_id[i] >= ma->l_range[1]) {
      set_error("csb_matrix_assembler_add_g_ids: pair (%lld, %lld) outside the local range "
                "[%lld, %lld): rows and columns of other ranks are not handled",
                (long long)row_g_id[i], (long long)col_g_id[i], ma->l_range[0], ma->l_range[1]);
      return -1;
    }
  }
  const size_t old = ma->g_rc.size();
  ma->g_rc.resize(old + 2*(size_t)n);
  for (csb_lnum_t i = 0; i < n; i++) {
    ma->g_rc[old + 2*(size_t)i] = row_g_id[i];
    ma->g_rc[old + 2*(size_t)i + 1] = col_g_id[i];
  }
  return 0;
}

/* _matrix_assembler_compute_local + _sort_and_compact_local (cs_matrix_assembler.cpp:1780,
   :230): per row, the sorted set of its declared columns, the diagonal apart */
int csb_matrix_assembler_compute(csb_matrix_assembler_t *ma)
{
  if (ma == nullptr) { set_error("csb_matrix_assembler_compute: null assembler"); return -1; }
  if (ma->computed) return 0;
  const csb_lnum_t n_rows = ma->n_rows;
  const size_t n = ma->g_rc.size() / 2;
  std::vector<long long> count((size_t)n_rows + 1, 0);
  for (size_t i = 0; i < n; i++)
    if (ma->g_rc[2*i] != ma->g_rc[2*i+1]) count[(size_t)(ma->g_rc[2*i] - ma->l_range[0]) + 1]++;
  for (csb_lnum_t r = 0; r < n_rows; r++) count[(size_t)r+1] += count[r];
  if (count[n_rows] > 0x7fffffffLL) {
    set_error("csb_matrix_assembler_compute: more than 2^31 declared pairs"); return -1;
  }
  HostVec<csb_lnum_t> wide((size_t)std::max<long long>(count[n_rows], 1));
  {
    std::vector<long long> fill(count.begin(), count.end() - 1);
    for (size_t i = 0; i < n; i++) {
      const long long r = ma->g_rc[2*i] - ma->l_range[0], c = ma->g_rc[2*i+1] - ma->l_range[0];
      if (r != c) wide[(size_t)fill[r]++] = (csb_lnum_t)c;
    }
  }
  ma->r_idx.assign((size_t)n_rows + 1, 0);
  #pragma omp parallel for schedule(static)
  for (csb_lnum_t r = 0; r < n_rows; r++) {
    csb_lnum_t *b = wide.data() + count[r], *e = wide.data() + count[(size_t)r+1];
    std::sort(b, e);
    ma->r_idx[(size_t)r+1] = (csb_lnum_t)(std::unique(b, e) - b);
  }
  for (csb_lnum_t r = 0; r < n_rows; r++) ma->r_idx[(size_t)r+1] += ma->r_idx[r];
  ma->c_id.resize((size_t)std::max(ma->r_idx[n_rows], 1));
  #pragma omp parallel for schedule(static)
  for (csb_lnum_t r = 0; r < n_rows; r++)
    std::copy(wide.data() + count[r], wide.data() + count[r] + (ma->r_idx[(size_t)r+1] - ma->r_idx[r]),
              ma->c_id.data() + ma->r_idx[r]);
  ma->computed = true;
  return 0;
}

csb_lnum_t csb_matrix_assembler_get_n_rows(const csb_matrix_assembler_t *ma)
{ return ma != nullptr ? ma->n_rows : 0; }

/* copies of cs_matrix_assembler_get_row_index / _get_col_ids; either pointer may be NULL;
   returns the number of extra-diagonal entries, -1 before csb_matrix_assembler_compute */
csb_lnum_t csb_matrix_assembler_get_structure(const csb_matrix_assembler_t *ma,
                                              csb_lnum_t *row_index, csb_lnum_t *col_id)
{
  if (ma == nullptr || !ma->computed) return -1;
  if (row_index != nullptr)
    memcpy(row_index, ma->r_idx.data(), sizeof(csb_lnum_t)*((size_t)ma->n_rows + 1));
  if (col_id != nullptr)
    memcpy(col_id, ma->c_id.data(), sizeof(csb_lnum_t)*(size_t)ma->r_idx[ma->n_rows]);
  return ma->r_idx[ma->n_rows];
}

/* cs_matrix_structure_create_from_assembler(CS_MATRIX_MSR, ma) + cs_matrix_create
   (cs_matrix.cpp:4468, :4682): the structure of the assembler, values zero */
csb_matrix_t *csb_matrix_create_from_assembler(const csb_matrix_assembler_t *ma)
{
  Context &c = ctx();
  if (ma == nullptr || !ma->computed) {
    set_error("csb_matrix_create_from_assembler: csb_matrix_assembler_compute first");
    return nullptr;
  }
  csb_matrix_t *m = new csb_matrix_t;
  m->n_rows = ma->n_rows;
  m->n_cols_ext = ma->n_rows;
  m->symmetric = false;            /* cs_matrix_assembler_values_init: _set_fill_info(false) */
  m->h_row_index = ma->r_idx;
  m->nnz = ma->r_idx[ma->n_rows];
  bool ok = m->row_index.upload(ma->r_idx.data(), (size_t)ma->n_rows + 1)
         && m->col_id.upload(ma->c_id.data(), (size_t)m->nnz)
         && m->d_val.alloc((size_t)m->n_rows) && m->x_val.alloc((size_t)m->nnz)
         && CSB_CUDA(cudaMemsetAsync(m->d_val.p, 0, sizeof(double)*(size_t)std::max(m->n_rows, 1), c.stream))
         && CSB_CUDA(cudaMemsetAsync(m->x_val.p, 0, sizeof(double)*(size_t)std::max(m->nnz, 1), c.stream))
         && CSB_CUDA(cudaStreamSynchronize(c.stream));
  if (!ok) { delete m; return nullptr; }
  matrix_plan_spmv(m);
  return m;
}

/* cs_matrix_assembler_values_init (cs_matrix.cpp:5635; _msr_assembler_values_init :2346):
   the values are zeroed */
csb_matrix_assembler_values_t *
csb_matrix_assembler_values_init(csb_matrix_t *m, const csb_matrix_assembler_t *ma,
                                 int diag_block_size, int extra_diag_block_size)
{
  Context &c = ctx();
  if (m == nullptr || ma == nullptr || !ma->computed || m->n_rows != ma->n_rows
      || m->nnz != ma->r_idx[ma->n_rows] || m->db != 1 || m->csr != nullptr || m->from_native) {
    set_error("csb_matrix_assembler_values_init: the matrix must come from "
              "csb_matrix_create_from_assembler on this assembler");
    return nullptr;
  }
  if (diag_block_size != 1 || extra_diag_block_size != 1) {
    set_error("csb_matrix_assembler_values_init: scalar coefficients only");
    return nullptr;
  }
  if ((size_t)m->n_rows + (size_t)m->nnz > 0x7fffffffULL) {
    set_error("csb_matrix_assembler_values_init: more than 2^31 matrix entries");
    return nullptr;
  }
  if (!CSB_CUDA(cudaMemsetAsync(m->d_val.p, 0, sizeof(double)*(size_t)std::max(m->n_rows, 1), c.stream))
      || !CSB_CUDA(cudaMemsetAsync(m->x_val.p, 0, sizeof(double)*(size_t)std::max(m->nnz, 1), c.stream)))
    return nullptr;
  m->symmetric = false;
  m->h_vals_valid = false;
  m->b_d_val = nullptr; m->b_x_val = nullptr;
  csb_matrix_assembler_values_t *mav = new csb_matrix_assembler_values_t;
  mav->m = m; mav->ma = ma;
  return mav;
}

/* cs_matrix_assembler_values_add_g (cs_matrix_assembler.cpp:3276): n coefficients by global
   row and column id.  The coefficients are parked in arrival order (host or device arrays
   are accepted, they are read before the call returns). */
int csb_matrix_assembler_values_add_g(csb_matrix_assembler_values_t *mav, csb_lnum_t n,
                                      const csb_gnum_t *row_g_id, const csb_gnum_t *col_g_id,
                                      const csb_real_t *val)
{
  if (mav == nullptr || n < 0
      || (n > 0 && (row_g_id == nullptr || col_g_id == nullptr || val == nullptr))) {
    set_error("csb_matrix_assembler_values_add_g: bad arguments"); return -1;
  }
  if (n == 0) return 0;
  const size_t old = mav->n_parked, tot = old + (size_t)n;
  if (tot > 0x7fffffffULL) {
    set_error("csb_matrix_assembler_values_add_g: more than 2^31 coefficients between two "
              "csb_matrix_assembler_values_done calls");
    return -1;
  }
  if (tot > mav->slot.size()) {
    const size_t cap = std::max<size_t>(tot, 2*mav->slot.size());
    HostVec<csb_lnum_t> s2(cap); HostVec<double> v2(cap);
    if (old > 0) {
      memcpy(s2.data(), mav->slot.data(), sizeof(csb_lnum_t)*old);
      memcpy(v2.data(), mav->val.data(), sizeof(double)*old);
    }
    mav->slot.swap(s2); mav->val.swap(v2);
  }
  if (!CSB_CUDA(cudaMemcpy(mav->val.data() + old, val, sizeof(double)*(size_t)n, cudaMemcpyDefault)))
    return -1;
  const _csb_matrix_assembler_t *ma = mav->ma;
  csb_lnum_t *slot = mav->slot.data() + old;
  long bad = -1;
  #pragma omp parallel for schedule(static) if (n > 4096)
  for (csb_lnum_t i = 0; i < n; i++) {
    slot[i] = slot_of(ma, row_g_id[i], col_g_id[i]);
    if (slot[i] < 0) {
      #pragma omp critical
      { if (bad < 0 || i < bad) bad = i; }
    }
  }
  if (bad >= 0) {
    set_error("csb_matrix_assembler_values_add_g: coefficient %ld, (%lld, %lld), is not in the "
              "assembler's structure", bad, (long long)row_g_id[bad], (long long)col_g_id[bad]);
    return -1;
  }
  mav->n_parked = tot;
  return 0;
}

/* cs_matrix_assembler_values_done (:3510): the parked coefficients are folded into the
   matrix, every entry taking its contributions in arrival order */
int csb_matrix_assembler_values_done(csb_matrix_assembler_values_t *mav)
{
  if (mav == nullptr) { set_error("csb_matrix_assembler_values_done: null"); return -1; }
  if (mav->n_parked == 0) return 0;
  Context &c = ctx();
  csb_matrix_t *m = mav->m;
  const size_t n = mav->n_parked, n_slots = (size_t)m->n_rows + (size_t)m->nnz;
  HostVec<csb_lnum_t> index, src;
  bucket_by_slot(n, mav->slot.data(), n_slots, index, src);
  DevBuf<csb_lnum_t> d_index, d_src;
  DevBuf<double> d_val;
  bool ok = d_index.upload(index.data(), n_slots + 1) && d_src.upload(src.data(), n)
         && d_val.upload(mav->val.data(), n)
         && launch_assembler_fold(m->n_rows, m->nnz, d_index.p, d_src.p, d_val.p, m->d_val.p,
                                  m->x_val.p)
         && CSB_CUDA(cudaStreamSynchronize(c.stream));
  mav->n_parked = 0;
  m->h_vals_valid = false;
  return ok ? 0 : -1;
}

/* The coefficients of ALL the pairs declared with csb_matrix_assembler_add_g_ids, in the
   order they were declared (val[k] belongs to the k-th declared pair; host or device array):
   what a time loop that calls cs_matrix_assembler_values_add_g with the same ids at every
   step adds.  The fold plan is built once per assembler and stays in HBM. */
int csb_matrix_assembler_values_add_declared(csb_matrix_assembler_values_t *mav,
                                             const csb_real_t *val)
{
  if (mav == nullptr || val == nullptr) {
    set_error("csb_matrix_assembler_values_add_declared: bad arguments"); return -1;
  }
  if (csb_matrix_assembler_values_done(mav) != 0) return -1;    /* keep the arrival order */
  Context &c = ctx();
  _csb_matrix_assembler_t *ma = const_cast<_csb_matrix_assembler_t *>(mav->ma);
  csb_matrix_t *m = mav->m;
  const size_t n = ma->g_rc.size() / 2, n_slots = (size_t)m->n_rows + (size_t)m->nnz;
  if (n == 0) return 0;
  if (!ma->plan_ready) {
    HostVec<csb_lnum_t> slot(n), index, src;
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++)
      slot[(size_t)i] = slot_of(ma, ma->g_rc[2*(size_t)i], ma->g_rc[2*(size_t)i+1]);
    bucket_by_slot(n, slot.data(), n_slots, index, src);
    if (!ma->plan_index.upload(index.data(), n_slots + 1) || !ma->plan_src.upload(src.data(), n)
        || !CSB_CUDA(cudaStreamSynchronize(c.stream))) return -1;
    ma->plan_ready = true;
  }
  DevBuf<double> staged;
  const double *d_val = val;
  if (!is_device_ptr(val)) {
    if (!staged.upload(val, n)) return -1;
    d_val = staged.p;
  }
  const bool ok = launch_assembler_fold(m->n_rows, m->nnz, ma->plan_index.p, ma->plan_src.p,
                                        d_val, m->d_val.p, m->x_val.p)
               && CSB_CUDA(cudaStreamSynchronize(c.stream));
  m->h_vals_valid = false;
  return ok ? 0 : -1;
}

/* cs_matrix_assembler_values_finalize (:3119) */
int csb_matrix_assembler_values_finalize(csb_matrix_assembler_values_t **mav)
{
  if (mav == nullptr || *mav == nullptr) return 0;
  const int r = csb_matrix_assembler_values_done(*mav);
  delete *mav;
  *mav = nullptr;
  return r;
}

} /* extern "C" */
