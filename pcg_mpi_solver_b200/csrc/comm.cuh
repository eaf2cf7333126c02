// comm.cuh - NCCL communicator + interface ("halo") exchange plan.
//
// Replaces the mpi4py traffic of the hot path: the three allreduces per iteration
// (MPI_SUM, pcg_solver.py:622-628) and the neighbour exchange-add of the matvec
// (pcg_solver.py:303-334).  NCCL is bound at run time (dlopen) so that libpcgb200.so loads on
// machines without NCCL and picks up the copy torch already mapped into the process.
#pragma once
#include <dlfcn.h>
#include <vector>

#include "common.cuh"
#include "pcg_kernels.cuh"

namespace pcgb {

// minimal NCCL ABI (stable since NCCL 2.x): types and enum values from nccl.h
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int *) = nullptr;
};

inline int nccl_api(NcclApi **out) {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char *names[] = {getenv("PCGB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      if (!n || !*n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (api.handle) {
#define PCGB_SYM(field, name) *(void **)(&api.field) = dlsym(api.handle, name)
      PCGB_SYM(GetUniqueId, "ncclGetUniqueId");
      PCGB_SYM(CommInitRank, "ncclCommInitRank");
      PCGB_SYM(CommDestroy, "ncclCommDestroy");
      PCGB_SYM(AllReduce, "ncclAllReduce");
      PCGB_SYM(Send, "ncclSend");
      PCGB_SYM(Recv, "ncclRecv");
      PCGB_SYM(GroupStart, "ncclGroupStart");
      PCGB_SYM(GroupEnd, "ncclGroupEnd");
      PCGB_SYM(GetErrorString, "ncclGetErrorString");
      PCGB_SYM(GetVersion, "ncclGetVersion");
#undef PCGB_SYM
    }
  }
  if (!api.handle || !api.CommInitRank || !api.AllReduce || !api.Send || !api.Recv)
    return fail(PCGB_ERR_NCCL, "NCCL not available: dlopen(libnccl.so.2) failed (%s)", dlerror() ? dlerror() : "missing symbols");
  *out = &api;
  return PCGB_OK;
}

#define PCGB_NCCL(api, call)                                                                   \
  do {                                                                                         \
    int e_ = (call);                                                                           \
    if (e_ != ::pcgb::ncclSuccess)                                                             \
      return ::pcgb::fail(PCGB_ERR_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #call,          \
                          (api)->GetErrorString ? (api)->GetErrorString(e_) : "nccl error");   \
  } while (0)

}  // namespace pcgb

struct pcgb_comm_s {
  pcgb::NcclApi *api = nullptr;
  pcgb::ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
};

struct pcgb_halo_s {
  pcgb_comm_t comm = nullptr;
  int n_nbr = 0;
  std::vector<int> nbr_rank;
  std::vector<int64_t> nbr_ptr;
  int64_t m = 0;          // total shared entries (with multiplicity over neighbours)
  int64_t ndof = 0;       // distinct interface dofs
  int *d_idx = nullptr;   // [m]   gather index for the pack
  int *d_dof = nullptr;   // [ndof] distinct dofs
  int *d_ptr = nullptr;   // [ndof+1]
  int *d_pos = nullptr;   // [m]   positions in the receive buffer, neighbour order per dof
  double *d_send = nullptr, *d_recv = nullptr;
};

namespace pcgb {

inline int halo_exchange_add(pcgb_halo_t h, double *y, cudaStream_t st, int *launches = nullptr) {
  if (!h || h->m == 0) return PCGB_OK;
  NcclApi *api = h->comm->api;
  k_halo_pack<<<(unsigned)((h->m + 255) / 256), 256, 0, st>>>(h->m, h->d_idx, y, h->d_send);
  PCGB_CHECK_LAUNCH();
  PCGB_NCCL(api, api->GroupStart());
  for (int j = 0; j < h->n_nbr; ++j) {
    const int64_t o = h->nbr_ptr[j], c = h->nbr_ptr[j + 1] - o;
    PCGB_NCCL(api, api->Send(h->d_send + o, (size_t)c, ncclFloat64, h->nbr_rank[j], h->comm->comm, st));
    PCGB_NCCL(api, api->Recv(h->d_recv + o, (size_t)c, ncclFloat64, h->nbr_rank[j], h->comm->comm, st));
  }
  PCGB_NCCL(api, api->GroupEnd());
  k_halo_unpack_add<<<(unsigned)((h->ndof + 255) / 256), 256, 0, st>>>(h->ndof, h->d_dof, h->d_ptr, h->d_pos, h->d_recv, y);
  PCGB_CHECK_LAUNCH();
  if (launches) *launches += 2;
  return PCGB_OK;
}

}  // namespace pcgb
