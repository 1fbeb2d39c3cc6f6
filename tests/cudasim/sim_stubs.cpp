/* tests/cudasim/sim_stubs.cpp -- TEST INFRASTRUCTURE (see cuda_runtime.h in this directory).
 * Entry points of persistent.cu / comm.cu / halo.cu, which need a real device: every one
 * reports an error, so the simulated library can only take the kernel-per-phase path. */
#include "csb_internal.h"
#include "persistent.h"

namespace csb {

static bool no(const char *what) { set_error("cudasim: %s needs a real device", what); return false; }

size_t persistent_smem_bytes(size_t) { return 1024; }
int    persistent_grid(int, size_t) { return 1; }
bool launch_pcg_persistent(const MgLevelDesc &, const double *, double *, bool, int, double,
                           double, const PersistentCommon &, ItState *, int, size_t, bool, bool)
{ return no("k_pcg_persistent"); }
bool launch_mg_persistent(const MgSolveParams &, const PersistentCommon &, int, size_t, bool)
{ return no("k_mg_persistent"); }
bool launch_fcg_persistent(const MgLevelDesc &, const double *, double *, bool, int, double,
                           double, const PersistentCommon &, ItState *, int, size_t, bool, bool,
                           bool)
{ return no("k_fcg_persistent"); }
bool launch_fcg_mg_persistent(const FcgMgParams &, const PersistentCommon &, int, size_t, bool)
{ return no("k_fcg_mg_persistent"); }
bool launch_spmv_pipeline(const _csb_matrix_t *, const double *, double *, bool, bool *launched)
{ *launched = false; return true; }
int plan_gather_depth(const std::vector<csb_lnum_t> &, int, int *m) { if (m) *m = 64; return 2; }
void plan_slabs(const std::vector<csb_lnum_t> &, int n_rows, int &slab_rows, int &n_slabs,
                size_t &max_entries)
{ slab_rows = 224; n_slabs = (n_rows + 223)/224; max_entries = 2432; }

void  *pcomm_arena_alloc(csb_pcomm_t *, size_t) { no("peer arena"); return nullptr; }
size_t pcomm_arena_mark(csb_pcomm_t *) { return 0; }
void   pcomm_arena_release(csb_pcomm_t *, size_t) {}
void   pcomm_host_barrier(csb_pcomm_t *) {}
bool   pcomm_allreduce(csb_pcomm_t *, double *, int) { return no("peer all-reduce"); }
bool   pcomm_allgather_count(csb_pcomm_t *, long long, long long *) { return no("peer all-gather"); }
bool   pcomm_gather_to_root(csb_pcomm_t *, int, const long long *, const void *, void *)
{ return no("peer gather"); }
bool   pcomm_mailbox(csb_pcomm_t *, int, const int *, const long long *, long long *)
{ return no("peer mailbox"); }
bool   pcomm_window_exchange(csb_pcomm_t *, const HaloDev &, int, int, const void *, void *)
{ return no("peer window exchange"); }
bool   halo_dev_view(const _csb_halo_t *, HaloDev &) { return no("halo"); }

} // namespace csb

extern "C" {
csb_halo_t *csb_halo_create(void *, int, int, const int *, csb_lnum_t, const csb_lnum_t *,
                            const csb_lnum_t *, const csb_lnum_t *)
{ csb::no("csb_halo_create"); return nullptr; }
void csb_halo_destroy(csb_halo_t **) {}
int  csb_halo_sync_var(const csb_halo_t *, csb_real_t *) { csb::no("csb_halo_sync_var"); return -1; }
int  csb_halo_set_pcomm(csb_halo_t *, csb_pcomm_t *) { csb::no("csb_halo_set_pcomm"); return -1; }
}
