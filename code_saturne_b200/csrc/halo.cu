/*============================================================================
 * halo.cu -- cs_halo_t ghost-cell exchange between GPUs (one process per GPU).
 *
 * Mirrors cs_halo_sync_pack / _start / _wait (src/base/cs_halo.cpp:1739,
 * :1957, :2130) and the reference's NCCL path (:518-587): gather the send
 * list into a device buffer, then one ncclSend + one ncclRecv per neighbour
 * rank inside a group, receiving straight into the ghost section of the
 * variable (ghosts are stored after the n_local_elts local values, one
 * contiguous section per neighbour, src/base/cs_halo.h:123-130).
 *
 * NCCL is resolved at run time with dlopen("libnccl.so.2") so that the
 * process uses the same NCCL that torch.distributed already loaded, and the
 * library has no link-time NCCL dependency for single-GPU use.
 *============================================================================*/

#include "csb_internal.h"

#include <dlfcn.h>

namespace {

typedef int ncclResult_t_;
typedef void *ncclComm_t_;
typedef struct { char internal[128]; } ncclUniqueId_;

struct NcclApi {
  bool  loaded = false;
  void *h = nullptr;
  ncclResult_t_ (*GetUniqueId)(ncclUniqueId_ *) = nullptr;
  ncclResult_t_ (*CommInitRank)(ncclComm_t_ *, int, ncclUniqueId_, int) = nullptr;
  ncclResult_t_ (*CommDestroy)(ncclComm_t_) = nullptr;
  ncclResult_t_ (*GroupStart)(void) = nullptr;
  ncclResult_t_ (*GroupEnd)(void) = nullptr;
  ncclResult_t_ (*Send)(const void *, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
  ncclResult_t_ (*Recv)(void *, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
  ncclResult_t_ (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t_,
                             cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t_) = nullptr;
};

NcclApi g_nccl;

const int NCCL_FLOAT64 = 8;   /* ncclDouble */
const int NCCL_SUM = 0;       /* ncclSum */

bool nccl_load()
{
  if (g_nccl.loaded) return true;
  const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] != nullptr && g_nccl.h == nullptr; i++)
    g_nccl.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (g_nccl.h == nullptr) {
    csb::set_error("NCCL library not found (dlopen libnccl.so.2): %s", dlerror());
    return false;
  }
#define CSB_SYM(field, name) \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.h, name); \
  if (g_nccl.field == nullptr) { csb::set_error("NCCL symbol %s missing", name); return false; }
  CSB_SYM(GetUniqueId, "ncclGetUniqueId")
  CSB_SYM(CommInitRank, "ncclCommInitRank")
  CSB_SYM(CommDestroy, "ncclCommDestroy")
  CSB_SYM(GroupStart, "ncclGroupStart")
  CSB_SYM(GroupEnd, "ncclGroupEnd")
  CSB_SYM(Send, "ncclSend")
  CSB_SYM(Recv, "ncclRecv")
  CSB_SYM(AllReduce, "ncclAllReduce")
  CSB_SYM(GetErrorString, "ncclGetErrorString")
#undef CSB_SYM
  g_nccl.loaded = true;
  return true;
}

bool nccl_ok(ncclResult_t_ r, const char *what)
{
  if (r == 0) return true;
  csb::set_error("%s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return false;
}

} // anonymous namespace

namespace csb {

/* in-place sum of n doubles over all ranks (dot products:
   ncclAllReduce(ncclDouble, ncclSum), src/alge/cs_sles_it_cuda.cu:1000-1006) */
bool comm_allreduce_sum(void *nccl_comm, double *buf_d, int n)
{
  if (nccl_comm == nullptr) return true;
  if (!nccl_load()) return false;
  return nccl_ok(g_nccl.AllReduce(buf_d, buf_d, (size_t)n, NCCL_FLOAT64, NCCL_SUM,
                                  (ncclComm_t_)nccl_comm, ctx().stream), "ncclAllReduce");
}

} // namespace csb

using namespace csb;

extern "C" {

/* Communicator bootstrap helpers (cs_base_cuda.cu:201-262 does the same with
   an MPI broadcast of the unique id; here the caller broadcasts the 128 bytes
   with torch.distributed). */
int csb_comm_get_unique_id(void *id_128_bytes)
{
  if (!nccl_load()) return -1;
  return nccl_ok(g_nccl.GetUniqueId((ncclUniqueId_ *)id_128_bytes), "ncclGetUniqueId") ? 0 : -1;
}

void *csb_comm_init_rank(const void *id_128_bytes, int n_ranks, int rank)
{
  if (!nccl_load()) return nullptr;
  ncclUniqueId_ id;
  memcpy(&id, id_128_bytes, sizeof(id));
  ncclComm_t_ comm = nullptr;
  if (!nccl_ok(g_nccl.CommInitRank(&comm, n_ranks, id, rank), "ncclCommInitRank"))
    return nullptr;
  return comm;
}

void csb_comm_destroy(void *nccl_comm)
{
  if (nccl_comm != nullptr && nccl_load()) g_nccl.CommDestroy((ncclComm_t_)nccl_comm);
}

csb_halo_t *
csb_halo_create(void *nccl_comm, int local_rank, int n_c_domains, const int *c_domain_rank,
                csb_lnum_t n_local_elts, const csb_lnum_t *send_index,
                const csb_lnum_t *send_list, const csb_lnum_t *index)
{
  if (!ctx().ready) { set_error("csb_halo_create: csb_init not called"); return nullptr; }
  if (n_c_domains < 0 || (n_c_domains > 0 && (c_domain_rank == nullptr || send_index == nullptr
                                              || index == nullptr))) {
    set_error("csb_halo_create: invalid arguments");
    return nullptr;
  }
  csb_halo_t *h = new csb_halo_t;
  h->nccl_comm = nccl_comm;
  h->local_rank = local_rank;
  h->n_c_domains = n_c_domains;
  h->n_local_elts = n_local_elts;
  h->c_domain_rank.assign(c_domain_rank, c_domain_rank + n_c_domains);
  h->send_index.assign(send_index, send_index + n_c_domains + 1);
  h->index.assign(index, index + n_c_domains + 1);
  h->n_send = (n_c_domains > 0) ? h->send_index[n_c_domains] : 0;
  h->n_ghost = (n_c_domains > 0) ? h->index[n_c_domains] : 0;
  h->h_send_list.resize((size_t)h->n_send);
  if (h->n_send > 0) {
    if (is_device_ptr(send_list)) {
      if (!CSB_CUDA(cudaMemcpy(h->h_send_list.data(), send_list, sizeof(csb_lnum_t)*(size_t)h->n_send,
                               cudaMemcpyDeviceToHost))) { delete h; return nullptr; }
    }
    else
      memcpy(h->h_send_list.data(), send_list, sizeof(csb_lnum_t)*(size_t)h->n_send);
  }
  if (!h->send_list.upload(send_list, (size_t)h->n_send) ||
      !h->send_buf.alloc((size_t)h->n_send) ||
      !CSB_CUDA(cudaStreamSynchronize(ctx().stream))) {
    delete h;
    return nullptr;
  }
  return h;
}

void csb_halo_destroy(csb_halo_t **halo)
{
  if (halo == nullptr || *halo == nullptr) return;
  delete *halo;
  *halo = nullptr;
}

/* Collective over the neighbour ranks: each side learns where its section
   starts among the neighbour's ghosts (the offset the reference's receiver
   passes to MPI_Irecv, src/base/cs_halo.cpp:1985-1996). */
int csb_halo_set_pcomm(csb_halo_t *h, csb_pcomm_t *comm)
{
  if (h == nullptr || comm == nullptr) { set_error("csb_halo_set_pcomm: null argument"); return -1; }
  const int nd = h->n_c_domains;
  for (int i = 0; i < nd; i++)
    if (h->c_domain_rank[i] < 0 || h->c_domain_rank[i] >= comm->h.n_ranks
        || h->c_domain_rank[i] == comm->h.rank) {
      set_error("csb_halo_set_pcomm: neighbour rank %d invalid for rank %d of %d",
                h->c_domain_rank[i], comm->h.rank, comm->h.n_ranks);
      return -1;
    }
  std::vector<long long> snd((size_t)nd*8 + 8, 0), rcv((size_t)nd*8 + 8, 0);
  for (int i = 0; i < nd; i++) snd[(size_t)i*8] = h->index[i];
  if (!pcomm_mailbox(comm, nd, h->c_domain_rank.data(), snd.data(), rcv.data())) return -1;
  h->peer_ghost_index.resize(nd);
  for (int i = 0; i < nd; i++) h->peer_ghost_index[i] = (int)rcv[(size_t)i*8];
  h->pcomm = comm;
  return 0;
}

int csb_halo_sync_var(const csb_halo_t *halo, csb_real_t *var_d)
{
  if (halo == nullptr) return 0;
  Context &c = ctx();
  csb_halo_t *h = const_cast<csb_halo_t *>(halo);
  if (h->pcomm != nullptr) {
    /* peer-memory exchange: gather + remote stores + flags in one kernel */
    HaloDev v;
    if (!halo_dev_view(h, v)) return -1;
    return pcomm_window_exchange(h->pcomm, v, 8, 1, var_d, var_d + h->n_local_elts) ? 0 : -1;
  }
  if (h->n_c_domains == 0) return 0;
  if (!launch_halo_pack(h->n_send, h->send_list.p, var_d, h->send_buf.p)) return -1;

  bool need_nccl = false;
  for (int i = 0; i < h->n_c_domains; i++)
    if (h->c_domain_rank[i] != h->local_rank) need_nccl = true;
  if (need_nccl) {
    if (h->nccl_comm == nullptr) {
      set_error("csb_halo_sync_var: halo has remote neighbours but no communicator");
      return -1;
    }
    if (!nccl_load()) return -1;
    if (!nccl_ok(g_nccl.GroupStart(), "ncclGroupStart")) return -1;
  }
  double *ghost = var_d + h->n_local_elts;
  for (int i = 0; i < h->n_c_domains; i++) {
    const int r = h->c_domain_rank[i];
    const size_t n_s = (size_t)(h->send_index[i+1] - h->send_index[i]);
    const size_t n_r = (size_t)(h->index[i+1] - h->index[i]);
    if (r == h->local_rank) {   /* periodicity / self section: local copy */
      if (n_r > 0 &&
          !CSB_CUDA(cudaMemcpyAsync(ghost + h->index[i], h->send_buf.p + h->send_index[i],
                                    sizeof(double)*n_r, cudaMemcpyDeviceToDevice, c.stream)))
        return -1;
    }
    else {
      if (n_s > 0 &&
          !nccl_ok(g_nccl.Send(h->send_buf.p + h->send_index[i], n_s, NCCL_FLOAT64, r,
                               (ncclComm_t_)h->nccl_comm, c.stream), "ncclSend")) return -1;
      if (n_r > 0 &&
          !nccl_ok(g_nccl.Recv(ghost + h->index[i], n_r, NCCL_FLOAT64, r,
                               (ncclComm_t_)h->nccl_comm, c.stream), "ncclRecv")) return -1;
    }
  }
  if (need_nccl && !nccl_ok(g_nccl.GroupEnd(), "ncclGroupEnd")) return -1;
  return 0;
}

} /* extern "C" */
