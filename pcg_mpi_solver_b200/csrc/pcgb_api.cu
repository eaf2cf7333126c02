// pcgb_api.cu - C ABI of libpcgb200.so (see include/pcgb200.h) and the host side of the PCG loop.
// sm_100a only.  No CPU fallback: every compute entry point requires a CUDA device.
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "spmv.cuh"
#include "pcg_kernels.cuh"
#include "comm.cuh"
#include "hexgen.cuh"
#include "ebe.cuh"
#include "ebe_color.cuh"
#include "assemble.cuh"

using namespace pcgb;

struct pcgb_csr_s {
  CsrPlan P;
};

struct pcgb_ebe_s {
  EbePlan P;
  int device = 0;
};

struct pcgb_asm_s {
  AsmPlan P;
};

struct pcgb_ebe2_s {
  EbeColorPlan C;
  int device = 0;
};

struct GraphKey {
  const void *minv, *w, *xb0, *resvec;
  int iters;
  int variant;   // what the captured iteration looks like: bit 0 peer transport, bit 1 interface-first split, bit 2 forked unpack
  bool operator==(const GraphKey &o) const {
    return minv == o.minv && w == o.w && xb0 == o.xb0 && resvec == o.resvec && iters == o.iters && variant == o.variant;
  }
};

struct pcgb_solver_s {
  pcgb_csr_t A = nullptr;
  pcgb_ebe_t E = nullptr;       // opt-in matrix-free operator (exactly one of A / E is set)
  pcgb_halo_t halo = nullptr;
  pcgb_comm_t comm = nullptr;
  int64_t n = 0;
  double *r = nullptr, *p = nullptr, *q = nullptr, *xalt = nullptr, *xown = nullptr;
  double *partials = nullptr;   // [5][kMaxVecGrid]
  double *stage = nullptr;      // first-level sums of the SpMV dot partials
  int stage_cap = 0;
  double *red = nullptr;        // [8] device scalars: [0]=pq, [1..5]=pp,xx,rr,rz,ninf ; [6..7] scratch
  PcgCtrl *d_ctrl = nullptr;
  PcgCtrl *h_ctrl = nullptr;    // pinned
  double *h_red = nullptr;      // pinned [8]
  cudaStream_t own = nullptr;   // the solve runs here: the caller's stream may be the legacy stream, which cannot be captured
  cudaStream_t side = nullptr;  // fork of the iteration: the halo unpack-add runs beside the p.q all-reduce
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr, ev_l0 = nullptr, ev_l1 = nullptr, ev_s0 = nullptr, ev_s1 = nullptr;
  std::vector<cudaEvent_t> ev_k;  // SpMV brackets (time_kernels)
  cudaGraphExec_t gexec = nullptr;
  GraphKey gkey{nullptr, nullptr, nullptr, nullptr, 0, 0};
  int launches = 0;             // launches issued outside graphs (running counter per solve)
  int launches_per_iter = 0;
  bool fork_halo = true;        // PCGB_FORK=0 keeps the unpack on the main stream
};

static int require_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(PCGB_ERR_NODEVICE, "no CUDA device visible: libpcgb200 has no CPU fallback");
  }
  return PCGB_OK;
}

namespace {
template <typename T>
cudaError_t upload(T **d, const std::vector<T> &v) {
  cudaError_t e = cudaMalloc(d, std::max<size_t>(v.size(), 1) * sizeof(T));
  if (e != cudaSuccess || v.empty()) return e;
  return cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}
// blob of a halo plan: IPC handle | m | per source rank r: {offset of r's segment in my receive buffer (doubles, -1 = not a
// neighbour), my neighbour slot of r, number of shared entries}
struct HaloBlobHead { unsigned char handle[PCGB_IPC_BLOB_BYTES]; int64_t m; };
struct HaloBlobRank { int64_t off, slot, cnt; };
}  // namespace

// Constant-memory slots of the 24-dof pattern matrices (c_ebe_ke24) are a per-device resource shared by every EBE
// operator of the process: slots are de-duplicated by content and reference-counted, so a second operator can never
// overwrite the matrices of a live one; when all slots are taken the group runs on the warp kernel (Ke from global memory).
namespace {
struct EbeSlot { int refs = 0; double ke[576]; };
struct EbeSlotTable { EbeSlot slot[kEbeMaxSlots]; };
std::mutex g_ebe_slot_mutex;
EbeSlotTable *ebe_slot_table(int device) {
  static std::vector<EbeSlotTable *> tabs;
  if (device < 0) return nullptr;
  if ((int)tabs.size() <= device) tabs.resize((size_t)device + 1, nullptr);
  if (!tabs[(size_t)device]) tabs[(size_t)device] = new EbeSlotTable();
  return tabs[(size_t)device];
}
// returns the slot holding `ke` (acquiring a reference) or -1 when the table is full
int ebe_slot_acquire(const double *ke) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  std::lock_guard<std::mutex> lock(g_ebe_slot_mutex);
  EbeSlotTable *t = ebe_slot_table(dev);
  int free_slot = -1;
  for (int k = 0; k < kEbeMaxSlots; ++k) {
    if (t->slot[k].refs > 0 && memcmp(t->slot[k].ke, ke, sizeof(t->slot[k].ke)) == 0) { t->slot[k].refs += 1; return k; }
    if (t->slot[k].refs == 0 && free_slot < 0) free_slot = k;
  }
  if (free_slot < 0) return -1;
  if (cudaMemcpyToSymbol(c_ebe_ke24, ke, 576 * sizeof(double), (size_t)free_slot * 576 * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  memcpy(t->slot[free_slot].ke, ke, sizeof(t->slot[free_slot].ke));
  t->slot[free_slot].refs = 1;
  return free_slot;
}
void ebe_slot_release(int device, int slot) {
  if (slot < 0 || slot >= kEbeMaxSlots) return;
  std::lock_guard<std::mutex> lock(g_ebe_slot_mutex);
  EbeSlotTable *t = ebe_slot_table(device);
  if (t && t->slot[slot].refs > 0) t->slot[slot].refs -= 1;
}
}  // namespace

extern "C" {

int pcgb_version(void) { return PCGB_VERSION; }
void pcgb_abi_sizes(int32_t out[4]) {
  out[0] = (int32_t)sizeof(pcgb_options); out[1] = (int32_t)sizeof(pcgb_result);
  out[2] = (int32_t)sizeof(pcgb_hex_box); out[3] = (int32_t)sizeof(pcgb_ebe_group);
}
const char *pcgb_last_error(void) { return last_error().c_str(); }
int pcgb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// ------------------------------------------------------------------------------------ CSR
int pcgb_csr_create(int64_t nrows, int64_t ncols, int64_t nnz, const void *d_rowptr, int rowptr_is_64, const int32_t *d_col,
                    const double *d_val, void *stream, pcgb_csr_t *out) {
  if (!out || nrows < 0 || ncols < 0 || nnz < 0 || !d_rowptr) return fail(PCGB_ERR_ARG, "pcgb_csr_create: bad argument");
  if (nrows > INT32_MAX - 8 || ncols > INT32_MAX - 8) return fail(PCGB_ERR_ARG, "pcgb_csr_create: more than 2^31 rows/cols");
  if (!rowptr_is_64 && nnz > INT32_MAX) return fail(PCGB_ERR_ARG, "pcgb_csr_create: nnz >= 2^31 needs 64-bit row offsets");
  PCGB_TRY(require_device());
  pcgb_csr_t A = new pcgb_csr_s();
  A->P.nrows = nrows; A->P.ncols = ncols; A->P.nnz = nnz;
  A->P.rowptr = d_rowptr; A->P.rp64 = rowptr_is_64 != 0; A->P.col = d_col; A->P.val = d_val;
  int rc = A->P.rp64 ? build_plan_t<int64_t>(A->P, (cudaStream_t)stream) : build_plan_t<int32_t>(A->P, (cudaStream_t)stream);
  if (rc != PCGB_OK) { free_plan(A->P); delete A; return rc; }
  *out = A;
  return PCGB_OK;
}

int pcgb_csr_destroy(pcgb_csr_t A) {
  if (!A) return PCGB_OK;
  free_plan(A->P);
  delete A;
  return PCGB_OK;
}

int pcgb_spmv(pcgb_csr_t A, const double *d_x, double *d_y, void *stream) {
  if (!A || !d_x || !d_y) return fail(PCGB_ERR_ARG, "pcgb_spmv: null argument");
  return spmv_launch(A->P, d_x, d_y, false, (cudaStream_t)stream);
}

static int csr_diag_compute(const CsrPlan &P, double *d_diag, cudaStream_t st) {
  if (P.nrows == 0) return PCGB_OK;
  if (!P.col) return fail(PCGB_ERR_ARG, "pcgb_csr_diag: the column array was released and no diagonal is cached");
  const unsigned grid = (unsigned)((P.nrows * 8 + 255) / 256);
  if (P.rp64) k_csr_diag<int64_t><<<grid, 256, 0, st>>>((const int64_t *)P.rowptr, P.col, P.val, P.nrows, d_diag);
  else k_csr_diag<int32_t><<<grid, 256, 0, st>>>((const int32_t *)P.rowptr, P.col, P.val, P.nrows, d_diag);
  PCGB_CHECK_LAUNCH();
  return PCGB_OK;
}

int pcgb_csr_diag(pcgb_csr_t A, double *d_diag, void *stream) {
  if (!A || !d_diag) return fail(PCGB_ERR_ARG, "pcgb_csr_diag: null argument");
  const CsrPlan &P = A->P;
  if (P.diag_cache) {
    PCGB_CUDA(cudaMemcpyAsync(d_diag, P.diag_cache, (size_t)P.nrows * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return PCGB_OK;
  }
  return csr_diag_compute(P, d_diag, (cudaStream_t)stream);
}

/* The selected SpMV kernel may not read d_col at all (the persistent staged-x kernel works on its own 16-bit index
 * stream).  pcgb_csr_release_col caches the diagonal (the one other consumer) and drops the library's reference, after
 * which the caller may free d_col (2 GB at 128^3, 16 GB at 256^3).  Returns PCGB_ERR_ARG if the plan still needs it. */
int pcgb_csr_release_col(pcgb_csr_t A, void *stream) {
  if (!A) return fail(PCGB_ERR_ARG, "pcgb_csr_release_col: null argument");
  CsrPlan &P = A->P;
  if (!P.col) return PCGB_OK;
  if (!P.persist) return fail(PCGB_ERR_ARG, "pcgb_csr_release_col: the selected SpMV kernel reads the column array");
  if (!P.diag_cache && P.nrows > 0) {
    PCGB_CUDA(cudaMalloc(&P.diag_cache, (size_t)P.nrows * sizeof(double)));
    int rc = csr_diag_compute(P, P.diag_cache, (cudaStream_t)stream);
    if (rc == PCGB_OK && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) rc = fail(PCGB_ERR_CUDA, "pcgb_csr_release_col: sync failed");
    if (rc != PCGB_OK) { cudaFree(P.diag_cache); P.diag_cache = nullptr; return rc; }
  }
  P.col = nullptr;
  return PCGB_OK;
}

int64_t pcgb_spmv_bytes(pcgb_csr_t A) {
  if (!A) return 0;
  const CsrPlan &P = A->P;
  return 12 * P.nnz + (P.rp64 ? 8 : 4) * (P.nrows + 1) + 8 * P.ncols + 8 * P.nrows;
}

/* bytes the selected kernel actually streams from HBM per SpMV: values 8 B per non-zero + the index stream
 * (4 B columns for the L1-gather kernel, 2 B staged positions, or 2/3 B with the column-triple index) + row offsets,
 * x, y and the tile / window tables */
int64_t pcgb_spmv_stream_bytes(pcgb_csr_t A) {
  if (!A) return 0;
  const CsrPlan &P = A->P;
  const int64_t idx_bytes = P.bsr ? (2 * P.nnz) / 9 : (P.persist && P.t3) ? (2 * P.nnz) / 3 : (P.staged ? 2 * P.nnz : 4 * P.nnz);
  return 8 * P.nnz + idx_bytes + (P.rp64 ? 8 : 4) * (P.nrows + 1) + 8 * P.ncols + 8 * P.nrows +
         (P.staged ? 8 * P.nwin + 8 * (int64_t)P.ntiles : 0) + (P.persist ? 32 : 12) * (int64_t)P.ntiles;
}

int pcgb_csr_plan_info(pcgb_csr_t A, int64_t info[16]) {
  if (!A || !info) return fail(PCGB_ERR_ARG, "pcgb_csr_plan_info: null argument");
  const CsrPlan &P = A->P;
  info[0] = P.ntiles; info[1] = P.tile_items; info[2] = P.lanes; info[3] = P.snap ? 1 : 0;
  info[4] = P.nfix; info[5] = P.staged ? P.smem_staged : P.smem_bytes; info[6] = P.max_row; info[7] = P.use_tma ? 1 : 0;
  info[8] = P.persist ? 2 : (P.staged ? 1 : 0); info[9] = P.nwin; info[10] = P.cap_x; info[11] = P.max_nw;
  if (P.persist) info[5] = P.smem_persist;
  info[12] = P.bsr ? 2 : (P.t3 ? 1 : 0); info[13] = P.desc_split ? P.nb_tiles : -1; info[14] = P.persist ? P.grid_persist : 0; info[15] = P.col ? 0 : 1;
  return PCGB_OK;
}

/* Interface-first split of the SpMV (multi-GPU overlap, SURVEY 8(e) "order boundary rows first, launch exchange, compute
 * interior rows"): d_rows = the distinct local rows that take part in the interface exchange.  The solver registers
 * them itself from its halo plan; the two entry points below exist so that the split can be tested on one GPU. */
int pcgb_csr_set_boundary_rows(pcgb_csr_t A, const int32_t *d_rows, int64_t count, void *stream) {
  if (!A || count < 0 || (count > 0 && !d_rows)) return fail(PCGB_ERR_ARG, "pcgb_csr_set_boundary_rows: bad argument");
  return spmv_set_boundary_rows(A->P, d_rows, count, (cudaStream_t)stream);
}
/* y = A x as two launches (interface tiles, then the rest); out_dot (device, may be NULL) receives x.y */
int pcgb_spmv_split(pcgb_csr_t A, const double *d_x, double *d_y, double *d_out_dot, void *stream) {
  if (!A || !d_x || !d_y) return fail(PCGB_ERR_ARG, "pcgb_spmv_split: null argument");
  const CsrPlan &P = A->P;
  if (!spmv_split_available(P)) return fail(PCGB_ERR_ARG, "pcgb_spmv_split: no split plan (persistent kernel + registered boundary rows needed)");
  cudaStream_t st = (cudaStream_t)stream;
  PCGB_TRY(spmv_launch_part(P, 0, d_x, d_y, d_out_dot != nullptr, st));
  PCGB_TRY(spmv_launch_part(P, 1, d_x, d_y, d_out_dot != nullptr, st));
  if (d_out_dot) {
    k_reduce<1, 0><<<1, 256, 0, st>>>(nullptr, P.dot_partials_split, spmv_split_dot_count(P), 0, d_out_dot, nullptr);
    PCGB_CHECK_LAUNCH();
  }
  return PCGB_OK;
}

// ------------------------------------------------------------------------------------ vector kernels
static int reduce_rows(const double *partials, int count, int nv, double *d_out, cudaStream_t st) {
  switch (nv) {
    case 1: k_reduce<1, 0><<<1, 256, 0, st>>>(nullptr, partials, count, kMaxVecGrid, d_out, nullptr); break;
    case 2: k_reduce<2, 0><<<1, 256, 0, st>>>(nullptr, partials, count, kMaxVecGrid, d_out, nullptr); break;
    default: k_reduce<5, 0><<<1, 256, 0, st>>>(nullptr, partials, count, kMaxVecGrid, d_out, nullptr); break;
  }
  PCGB_CHECK_LAUNCH();
  return PCGB_OK;
}

int pcgb_dot_w(int64_t n, const double *d_a, const double *d_b, const double *d_w, double *d_out, void *stream) {
  if (n < 0 || !d_out || (n > 0 && (!d_a || !d_b))) return fail(PCGB_ERR_ARG, "pcgb_dot_w: bad argument");
  PCGB_TRY(require_device());
  cudaStream_t st = (cudaStream_t)stream;
  // stream-ordered scratch on the CURRENT device: no state shared between calls, streams or devices
  double *scratch = nullptr;
  PCGB_CUDA(cudaMallocAsync(&scratch, kMaxVecGrid * sizeof(double), st));
  const int grid = vec_grid(n);
  k_dot_w<<<grid, kVecBlock, 0, st>>>(n, d_a, d_b, d_w, scratch);
  int rc = cudaGetLastError() == cudaSuccess ? reduce_rows(scratch, grid, 1, d_out, st) : fail(PCGB_ERR_CUDA, "pcgb_dot_w: launch failed");
  cudaFreeAsync(scratch, st);
  return rc;
}

int pcgb_axpby(int64_t n, double a, const double *d_x, double b, double *d_y, void *stream) {
  if (n < 0 || (n > 0 && (!d_x || !d_y))) return fail(PCGB_ERR_ARG, "pcgb_axpby: bad argument");
  if (n == 0) return PCGB_OK;
  k_axpby<<<vec_grid(n), kVecBlock, 0, (cudaStream_t)stream>>>(n, a, d_x, b, d_y);
  PCGB_CHECK_LAUNCH();
  return PCGB_OK;
}

int pcgb_mul(int64_t n, const double *d_x, const double *d_y, double *d_z, void *stream) {
  if (n < 0 || (n > 0 && (!d_x || !d_y || !d_z))) return fail(PCGB_ERR_ARG, "pcgb_mul: bad argument");
  if (n == 0) return PCGB_OK;
  k_mul<<<vec_grid(n), kVecBlock, 0, (cudaStream_t)stream>>>(n, d_x, d_y, d_z);
  PCGB_CHECK_LAUNCH();
  return PCGB_OK;
}

int pcgb_reciprocal(int64_t n, const double *d_d, double *d_out, void *stream) {
  if (n < 0 || (n > 0 && (!d_d || !d_out))) return fail(PCGB_ERR_ARG, "pcgb_reciprocal: bad argument");
  if (n == 0) return PCGB_OK;
  k_reciprocal<<<vec_grid(n), kVecBlock, 0, (cudaStream_t)stream>>>(n, d_d, d_out);
  PCGB_CHECK_LAUNCH();
  return PCGB_OK;
}

// ------------------------------------------------------------------------------------ communicator
int pcgb_comm_unique_id(unsigned char id[PCGB_UNIQUE_ID_BYTES]) {
  NcclApi *api = nullptr;
  PCGB_TRY(nccl_api(&api));
  ncclUniqueId uid;
  PCGB_NCCL(api, api->GetUniqueId(&uid));
  static_assert(sizeof(uid) == PCGB_UNIQUE_ID_BYTES, "ncclUniqueId size");
  memcpy(id, &uid, sizeof(uid));
  return PCGB_OK;
}

int pcgb_comm_create(int rank, int nranks, const unsigned char *id, pcgb_comm_t *out) {
  if (!out || rank < 0 || rank >= nranks) return fail(PCGB_ERR_ARG, "pcgb_comm_create: bad argument");
  if (!id && nranks > kMaxPeers) return fail(PCGB_ERR_ARG, "pcgb_comm_create: a peer-only communicator holds at most %d ranks", kMaxPeers);
  PCGB_TRY(require_device());
  pcgb_comm_t c = new pcgb_comm_s();
  c->rank = rank; c->nranks = nranks;
  cudaGetDevice(&c->device);
  int rc = PCGB_OK;
  if (id) {
    NcclApi *api = nullptr;
    rc = nccl_api(&api);
    if (rc == PCGB_OK) {
      ncclUniqueId uid;
      memcpy(&uid, id, sizeof(uid));
      c->api = api;
      int e = api->CommInitRank(&c->comm, nranks, uid, rank);
      if (e != ncclSuccess) rc = fail(PCGB_ERR_NCCL, "ncclCommInitRank -> %s", api->GetErrorString ? api->GetErrorString(e) : "error");
    }
  }
  // the peer window: zeroed before anyone can import it
  cudaError_t ce = cudaSuccess;
  if (rc == PCGB_OK && nranks <= kMaxPeers) {
    if ((ce = cudaMalloc(&c->win, kWinBytes)) == cudaSuccess && (ce = cudaMemset(c->win, 0, kWinBytes)) == cudaSuccess &&
        (ce = cudaMalloc(&c->d_status, sizeof(int))) == cudaSuccess && (ce = cudaMemset(c->d_status, 0, sizeof(int))) == cudaSuccess &&
        (ce = cudaMallocHost(&c->h_status, sizeof(int))) == cudaSuccess)
      ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) rc = fail(PCGB_ERR_CUDA, "pcgb_comm_create: window allocation -> %s", cudaGetErrorString(ce));
  }
  const char *tr = getenv("PCGB_COMM");
  c->transport = (tr && (tr[0] == 'n' || tr[0] == 'N')) ? PCGB_TRANSPORT_NCCL : PCGB_TRANSPORT_PEER;
  if (rc != PCGB_OK) { pcgb_comm_destroy(c); return rc; }
  *out = c;
  return PCGB_OK;
}

int pcgb_comm_destroy(pcgb_comm_t c) {
  if (!c) return PCGB_OK;
  for (int r = 0; r < kMaxPeers; ++r)
    if (c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
  cudaFree(c->win); cudaFree(c->d_status);
  if (c->h_status) cudaFreeHost(c->h_status);
  if (c->comm && c->api && c->api->CommDestroy) c->api->CommDestroy(c->comm);
  delete c;
  return PCGB_OK;
}

/* Peer windows: every rank exports the CUDA IPC handle of its window, the host code all-gathers the blobs over ANY
 * side channel (torch.distributed object collectives, mpi4py allgather, ...) and every rank imports them.  From then
 * on the all-reduces and the halo exchange of this communicator are the library's own kernels over NVLink peer memory
 * (csrc/peer.cuh) unless PCGB_COMM=nccl / pcgb_comm_set_transport(c, 0) keeps NCCL. */
int pcgb_comm_window_export(pcgb_comm_t c, unsigned char blob[PCGB_IPC_BLOB_BYTES]) {
  if (!c || !blob) return fail(PCGB_ERR_ARG, "pcgb_comm_window_export: null argument");
  if (!c->win) return fail(PCGB_ERR_ARG, "pcgb_comm_window_export: communicator has no window (more than %d ranks)", kMaxPeers);
  static_assert(sizeof(cudaIpcMemHandle_t) == PCGB_IPC_BLOB_BYTES, "cudaIpcMemHandle_t size");
  cudaIpcMemHandle_t h;
  PCGB_CUDA(cudaIpcGetMemHandle(&h, c->win));
  memcpy(blob, &h, sizeof(h));
  return PCGB_OK;
}

int pcgb_comm_window_import(pcgb_comm_t c, const unsigned char *blobs) {
  if (!c || !blobs) return fail(PCGB_ERR_ARG, "pcgb_comm_window_import: null argument");
  if (!c->win) return fail(PCGB_ERR_ARG, "pcgb_comm_window_import: communicator has no window");
  for (int r = 0; r < c->nranks; ++r) {
    if (r == c->rank || c->peer[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, blobs + (size_t)r * PCGB_IPC_BLOB_BYTES, sizeof(h));
    void *ptr = nullptr;
    PCGB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer[r] = static_cast<unsigned long long *>(ptr);
  }
  c->peer_ready = true;
  return PCGB_OK;
}

int pcgb_comm_transport(pcgb_comm_t c) { return c && c->use_peer() ? PCGB_TRANSPORT_PEER : PCGB_TRANSPORT_NCCL; }
int pcgb_comm_set_transport(pcgb_comm_t c, int transport) {
  if (!c || (transport != PCGB_TRANSPORT_NCCL && transport != PCGB_TRANSPORT_PEER)) return fail(PCGB_ERR_ARG, "pcgb_comm_set_transport: bad argument");
  if (transport == PCGB_TRANSPORT_PEER && !c->peer_ready) return fail(PCGB_ERR_ARG, "pcgb_comm_set_transport: peer windows have not been imported");
  if (transport == PCGB_TRANSPORT_NCCL && !c->api) return fail(PCGB_ERR_ARG, "pcgb_comm_set_transport: communicator was created without NCCL");
  c->transport = transport;
  return PCGB_OK;
}
/* 0 = healthy; non-zero after a peer kernel gave up waiting for another rank (sticky) */
int pcgb_comm_status(pcgb_comm_t c, void *stream) {
  if (!c || !c->d_status) return 0;
  if (cudaMemcpyAsync(c->h_status, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess ||
      cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess)
    return -1;
  return *c->h_status;
}

int pcgb_allreduce_sum(pcgb_comm_t c, double *d_buf, int count, void *stream) {
  if (!c || !d_buf || count < 0) return fail(PCGB_ERR_ARG, "pcgb_allreduce_sum: bad argument");
  return allreduce_sum(c, d_buf, count, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------ halo plan

int pcgb_halo_create(pcgb_comm_t c, int n_nbr, const int32_t *nbr_rank, const int64_t *nbr_ptr, const int64_t *idx_host,
                     int64_t nlocal, pcgb_halo_t *out) {
  if (!out || n_nbr < 0 || (n_nbr > 0 && (!c || !nbr_rank || !nbr_ptr))) return fail(PCGB_ERR_ARG, "pcgb_halo_create: bad argument");
  PCGB_TRY(require_device());
  pcgb_halo_t h = new pcgb_halo_s();
  h->comm = c; h->n_nbr = n_nbr;
  h->nbr_rank.assign(nbr_rank, nbr_rank + n_nbr);
  if (n_nbr > 0) h->nbr_ptr.assign(nbr_ptr, nbr_ptr + n_nbr + 1);
  else h->nbr_ptr.assign(1, 0);
  h->m = h->nbr_ptr.back();
  int rc = PCGB_OK;
  cudaError_t ce = cudaSuccess;
  if (h->m > INT32_MAX) rc = fail(PCGB_ERR_ARG, "pcgb_halo_create: too many shared dofs");
  for (int j = 0; j < n_nbr && rc == PCGB_OK; ++j)
    if (nbr_rank[j] < 0 || nbr_rank[j] >= c->nranks || nbr_rank[j] == c->rank || nbr_ptr[j + 1] < nbr_ptr[j])
      rc = fail(PCGB_ERR_ARG, "pcgb_halo_create: bad neighbour table at slot %d", j);
  std::vector<int> idx((size_t)h->m), ent_nbr((size_t)h->m), nptr((size_t)n_nbr + 1);
  if (rc == PCGB_OK && h->m > 0) {
    for (int j = 0; j < n_nbr; ++j)
      for (int64_t k = nbr_ptr[j]; k < nbr_ptr[j + 1]; ++k) ent_nbr[(size_t)k] = j;
    for (int64_t k = 0; k < h->m && rc == PCGB_OK; ++k) {
      if (idx_host[k] < 0 || idx_host[k] >= nlocal) rc = fail(PCGB_ERR_ARG, "pcgb_halo_create: index %lld out of range", (long long)idx_host[k]);
      else idx[(size_t)k] = (int)idx_host[k];
    }
  }
  if (rc == PCGB_OK) {
    for (int j = 0; j <= n_nbr; ++j) nptr[(size_t)j] = (int)h->nbr_ptr[(size_t)j];
    // group the receive positions by dof, neighbour order inside a dof (stable: k ascending = neighbour ascending)
    std::vector<int> order((size_t)h->m);
    for (int64_t k = 0; k < h->m; ++k) order[(size_t)k] = (int)k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return idx[a] < idx[b]; });
    std::vector<int> dof, ptr, pos((size_t)h->m);
    for (int64_t s = 0; s < h->m; ++s) {
      const int k = order[(size_t)s];
      if (s == 0 || idx[k] != idx[order[(size_t)s - 1]]) { dof.push_back(idx[k]); ptr.push_back((int)s); }
      pos[(size_t)s] = k;
    }
    ptr.push_back((int)h->m);
    h->ndof = (int64_t)dof.size();
    // receive block shared with the neighbours: [2][m] doubles + [n_nbr][2] flags (peer transport)
    h->blk_bytes = (size_t)2 * (size_t)h->m * sizeof(double) + (size_t)2 * (size_t)std::max(n_nbr, 1) * sizeof(unsigned long long);
    if ((ce = upload(&h->d_idx, idx)) == cudaSuccess && (ce = upload(&h->d_dof, dof)) == cudaSuccess &&
        (ce = upload(&h->d_ptr, ptr)) == cudaSuccess && (ce = upload(&h->d_pos, pos)) == cudaSuccess &&
        (ce = upload(&h->d_nbr_ptr, nptr)) == cudaSuccess && (ce = upload(&h->d_ent_nbr, ent_nbr)) == cudaSuccess &&
        (ce = cudaMalloc(&h->d_send, std::max<size_t>((size_t)h->m, 1) * sizeof(double))) == cudaSuccess &&
        (ce = cudaMalloc(&h->d_recv, std::max<size_t>((size_t)h->m, 1) * sizeof(double))) == cudaSuccess &&
        (ce = cudaMalloc(&h->blk, h->blk_bytes)) == cudaSuccess && (ce = cudaMemset(h->blk, 0, h->blk_bytes)) == cudaSuccess &&
        (ce = cudaMalloc(&h->d_epoch, 2 * sizeof(unsigned long long))) == cudaSuccess &&
        (ce = cudaMemset(h->d_epoch, 0, 2 * sizeof(unsigned long long))) == cudaSuccess)
      ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) rc = fail(PCGB_ERR_CUDA, "pcgb_halo_create: %s", cudaGetErrorString(ce));
  }
  if (rc != PCGB_OK) { pcgb_halo_destroy(h); return rc; }
  *out = h;
  return PCGB_OK;
}

int pcgb_halo_destroy(pcgb_halo_t h) {
  if (!h) return PCGB_OK;
  for (void *p : h->opened) cudaIpcCloseMemHandle(p);
  cudaFree(h->d_idx); cudaFree(h->d_dof); cudaFree(h->d_ptr); cudaFree(h->d_pos); cudaFree(h->d_send); cudaFree(h->d_recv);
  cudaFree(h->blk); cudaFree(h->d_nbr_ptr); cudaFree(h->d_ent_nbr); cudaFree(h->d_remote); cudaFree(h->d_remote_m);
  cudaFree(h->d_remote_flag); cudaFree(h->d_epoch);
  delete h;
  return PCGB_OK;
}

/* Peer transport of the halo: export / all-gather (host side channel) / import, like the communicator window.
 * Every rank of the communicator takes part, also one without neighbours. */
int64_t pcgb_halo_blob_bytes(pcgb_halo_t h) {
  if (!h || !h->comm) return 0;
  return (int64_t)(sizeof(HaloBlobHead) + (size_t)h->comm->nranks * sizeof(HaloBlobRank));
}

int pcgb_halo_export(pcgb_halo_t h, unsigned char *blob) {
  if (!h || !blob || !h->comm) return fail(PCGB_ERR_ARG, "pcgb_halo_export: null argument");
  HaloBlobHead head;
  memset(&head, 0, sizeof(head));
  cudaIpcMemHandle_t ih;
  PCGB_CUDA(cudaIpcGetMemHandle(&ih, h->blk));
  memcpy(head.handle, &ih, sizeof(ih));
  head.m = h->m;
  memcpy(blob, &head, sizeof(head));
  std::vector<HaloBlobRank> tab((size_t)h->comm->nranks, HaloBlobRank{-1, -1, 0});
  for (int j = 0; j < h->n_nbr; ++j) tab[(size_t)h->nbr_rank[j]] = HaloBlobRank{h->nbr_ptr[j], j, h->nbr_ptr[j + 1] - h->nbr_ptr[j]};
  memcpy(blob + sizeof(head), tab.data(), tab.size() * sizeof(HaloBlobRank));
  return PCGB_OK;
}

int pcgb_halo_import(pcgb_halo_t h, const unsigned char *blobs) {
  if (!h || !blobs || !h->comm) return fail(PCGB_ERR_ARG, "pcgb_halo_import: null argument");
  if (h->peer_ready) return PCGB_OK;
  const size_t stride = (size_t)pcgb_halo_blob_bytes(h);
  const int me = h->comm->rank;
  std::vector<double *> remote((size_t)h->n_nbr, nullptr);
  std::vector<int64_t> remote_m((size_t)h->n_nbr, 0);
  std::vector<unsigned long long *> remote_flag((size_t)h->n_nbr, nullptr);
  for (int j = 0; j < h->n_nbr; ++j) {
    const unsigned char *b = blobs + (size_t)h->nbr_rank[j] * stride;
    HaloBlobHead head;
    HaloBlobRank mine;
    memcpy(&head, b, sizeof(head));
    memcpy(&mine, b + sizeof(head) + (size_t)me * sizeof(HaloBlobRank), sizeof(mine));
    const int64_t cnt = h->nbr_ptr[j + 1] - h->nbr_ptr[j];
    if (mine.off < 0 || mine.cnt != cnt)
      return fail(PCGB_ERR_ARG, "pcgb_halo_import: rank %d shares %lld dofs with rank %d but that rank lists %lld for us (interface lists must match)",
                  me, (long long)cnt, h->nbr_rank[j], (long long)mine.cnt);
    cudaIpcMemHandle_t ih;
    memcpy(&ih, head.handle, sizeof(ih));
    void *ptr = nullptr;
    PCGB_CUDA(cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess));
    h->opened.push_back(ptr);
    remote[(size_t)j] = static_cast<double *>(ptr) + mine.off;
    remote_m[(size_t)j] = head.m;
    remote_flag[(size_t)j] = reinterpret_cast<unsigned long long *>(static_cast<unsigned char *>(ptr) + (size_t)2 * (size_t)head.m * sizeof(double)) + 2 * mine.slot;
  }
  PCGB_CUDA(upload(&h->d_remote, remote));
  PCGB_CUDA(upload(&h->d_remote_m, remote_m));
  PCGB_CUDA(upload(&h->d_remote_flag, remote_flag));
  h->peer_ready = true;
  return PCGB_OK;
}

int pcgb_halo_exchange_add(pcgb_halo_t h, double *d_y, void *stream) {
  if (!h || !d_y) return fail(PCGB_ERR_ARG, "pcgb_halo_exchange_add: null argument");
  return halo_exchange_add(h, d_y, (cudaStream_t)stream);
}

int64_t pcgb_halo_bytes(pcgb_halo_t h) { return h ? h->m * 8 : 0; }

// ------------------------------------------------------------------------------------ EBE operator (opt-in, f1)
int pcgb_ebe_create(int64_t n, int ngroups, const pcgb_ebe_group *groups, pcgb_ebe_t *out) {
  if (!out || n < 0 || ngroups < 0 || (ngroups > 0 && !groups)) return fail(PCGB_ERR_ARG, "pcgb_ebe_create: bad argument");
  if (n >= (1 << 30)) return fail(PCGB_ERR_ARG, "pcgb_ebe_create: more than 2^30 dofs");
  PCGB_TRY(require_device());
  pcgb_ebe_t E = new pcgb_ebe_s();
  cudaGetDevice(&E->device);
  EbePlan &P = E->P;
  P.n = n;
  std::vector<int> blk_group;
  std::vector<int64_t> blk_e0;
  P.bytes = 16 * n;
  int rc = PCGB_OK;
  cudaError_t ce = cudaSuccess;
  for (int g = 0; g < ngroups && rc == PCGB_OK; ++g) {
    const pcgb_ebe_group &src = groups[g];
    if (src.nd <= 0 || src.nd > 96 || src.ne < 0 || !src.ke_host || (src.ne > 0 && (!src.d_idx || !src.d_ck))) {
      rc = fail(PCGB_ERR_ARG, "pcgb_ebe_create: group %d: pattern size must be 1..96 and arrays non-null", g);
      break;
    }
    EbeGroup eg;
    eg.nd = src.nd; eg.ne = src.ne; eg.idx = src.d_idx; eg.sign = src.d_sign; eg.ck = src.d_ck;
    double *dke = nullptr;
    if ((ce = cudaMalloc(&dke, (size_t)src.nd * src.nd * sizeof(double))) != cudaSuccess ||
        (ce = cudaMemcpy(dke, src.ke_host, (size_t)src.nd * src.nd * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) {
      cudaFree(dke);
      rc = fail(PCGB_ERR_CUDA, "pcgb_ebe_create: %s", cudaGetErrorString(ce));
      break;
    }
    eg.ke = dke;
    if (src.nd == 24) eg.slot = ebe_slot_acquire(src.ke_host);
    if (eg.slot < 0)
      for (int64_t e0 = 0; e0 < src.ne; e0 += kEbeWarpsPerBlock) { blk_group.push_back((int)P.groups.size()); blk_e0.push_back(e0); }
    P.bytes += src.ne * ((int64_t)src.nd * (4 + (src.d_sign ? 1 : 0)) + 8);
    P.groups.push_back(eg);
  }
  if (rc == PCGB_OK) {
    P.nblk_warp = (int)blk_group.size();
    if ((ce = upload(&P.d_groups, P.groups)) != cudaSuccess || (ce = upload(&P.d_blk_group, blk_group)) != cudaSuccess ||
        (ce = upload(&P.d_blk_e0, blk_e0)) != cudaSuccess)
      rc = fail(PCGB_ERR_CUDA, "pcgb_ebe_create: %s", cudaGetErrorString(ce));
  }
  if (rc != PCGB_OK) { pcgb_ebe_destroy(E); return rc; }
  *out = E;
  return PCGB_OK;
}

int pcgb_ebe_destroy(pcgb_ebe_t E) {
  if (!E) return PCGB_OK;
  for (EbeGroup &g : E->P.groups) {
    cudaFree(const_cast<double *>(g.ke));
    ebe_slot_release(E->device, g.slot);
  }
  cudaFree(E->P.d_groups); cudaFree(E->P.d_blk_group); cudaFree(E->P.d_blk_e0);
  delete E;
  return PCGB_OK;
}

int pcgb_ebe_apply(pcgb_ebe_t E, const double *d_x, double *d_y, void *stream) {
  if (!E || !d_x || !d_y) return fail(PCGB_ERR_ARG, "pcgb_ebe_apply: null argument");
  return ebe_apply(E->P, d_x, d_y, (cudaStream_t)stream);
}

int64_t pcgb_ebe_bytes(pcgb_ebe_t E) { return E ? E->P.bytes : 0; }

// ------------------------------------------------------------------------------------ device assembly (f2)
int pcgb_assemble_symbolic(int64_t n, int ngroups, const pcgb_ebe_group *groups, int64_t *d_rowptr, int64_t *nnz_out, void *stream,
                           pcgb_asm_t *out) {
  if (!out || !d_rowptr || !nnz_out || n < 0 || ngroups < 0 || (ngroups > 0 && !groups)) return fail(PCGB_ERR_ARG, "pcgb_assemble_symbolic: bad argument");
  if (n >= (1ll << 31) - 8 || ngroups >= 65536) return fail(PCGB_ERR_ARG, "pcgb_assemble_symbolic: too many rows / pattern groups");
  PCGB_TRY(require_device());
  pcgb_asm_t a = new pcgb_asm_s();
  a->P.n = n;
  int rc = PCGB_OK;
  for (int g = 0; g < ngroups && rc == PCGB_OK; ++g) {
    const pcgb_ebe_group &src = groups[g];
    if (src.nd <= 0 || src.nd > 255 || src.ne < 0 || src.ne >= (1ll << 40) || !src.ke_host || (src.ne > 0 && (!src.d_idx || !src.d_ck))) {
      rc = fail(PCGB_ERR_ARG, "pcgb_assemble_symbolic: group %d: pattern size must be 1..255 and arrays non-null", g);
      break;
    }
    AsmGroup ag;
    ag.nd = src.nd; ag.ne = src.ne; ag.idx = src.d_idx; ag.sign = src.d_sign; ag.ck = src.d_ck;
    double *dke = nullptr;
    cudaError_t ce = cudaMalloc(&dke, (size_t)src.nd * src.nd * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMemcpy(dke, src.ke_host, (size_t)src.nd * src.nd * sizeof(double), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { cudaFree(dke); rc = fail(PCGB_ERR_CUDA, "pcgb_assemble_symbolic: %s", cudaGetErrorString(ce)); break; }
    ag.ke = dke;
    a->P.groups.push_back(ag);
  }
  if (rc == PCGB_OK) rc = asm_symbolic(a->P, d_rowptr, (cudaStream_t)stream);
  if (rc != PCGB_OK) { pcgb_assemble_destroy(a); return rc; }
  *nnz_out = a->P.nnz;
  *out = a;
  return PCGB_OK;
}

int pcgb_assemble_numeric(pcgb_asm_t a, const int64_t *d_rowptr, int32_t *d_col, double *d_val, void *stream) {
  if (!a || !d_rowptr || (a->P.nnz > 0 && (!d_col || !d_val))) return fail(PCGB_ERR_ARG, "pcgb_assemble_numeric: null argument");
  return asm_numeric(a->P, d_rowptr, d_col, d_val, (cudaStream_t)stream);
}

int pcgb_assemble_destroy(pcgb_asm_t a) {
  if (!a) return PCGB_OK;
  asm_free(a->P);
  delete a;
  return PCGB_OK;
}

// ------------------------------------------------------------------------------------ coloured EBE operator (operator level)
// groups[] holds one entry per (pattern group, colour), sorted by colour; phase[g] is the colour.  Pattern matrices of
// the 24-dof groups are de-duplicated into the constant-memory slots by content of ke_host (same pointer = same slot).
int pcgb_ebe2_create(int64_t n, int ngroups, const pcgb_ebe_group *groups, const int32_t *phase, pcgb_ebe2_t *out) {
  if (!out || n < 0 || ngroups < 0 || (ngroups > 0 && (!groups || !phase))) return fail(PCGB_ERR_ARG, "pcgb_ebe2_create: bad argument");
  if (n >= (1 << 30)) return fail(PCGB_ERR_ARG, "pcgb_ebe2_create: more than 2^30 dofs");
  PCGB_TRY(require_device());
  for (int g = 1; g < ngroups; ++g)
    if (phase[g] < phase[g - 1]) return fail(PCGB_ERR_ARG, "pcgb_ebe2_create: groups must be sorted by phase (colour)");
  pcgb_ebe2_t E = new pcgb_ebe2_s();
  cudaGetDevice(&E->device);
  EbePlan &P = E->C.P;
  P.n = n;
  P.bytes = 16 * n;
  std::vector<int> blk_group;
  std::vector<int64_t> blk_e0;
  std::vector<std::pair<const double *, double *>> ke_dev;   // host pointer -> device copy (shared between colours)
  int open_warp_phase = -1, open_warp_begin = 0;
  auto flush_warp = [&]() {
    if (open_warp_phase >= 0 && (int)blk_group.size() > open_warp_begin) E->C.launches.push_back({1, open_warp_begin, (int)blk_group.size()});
    open_warp_phase = -1;
  };
  for (int g = 0; g < ngroups; ++g) {
    const pcgb_ebe_group &src = groups[g];
    if (src.nd <= 0 || src.nd > 96 || src.ne < 0 || !src.ke_host || (src.ne > 0 && (!src.d_idx || !src.d_ck))) {
      delete E;
      return fail(PCGB_ERR_ARG, "pcgb_ebe2_create: group %d: pattern size must be 1..96 and arrays non-null", g);
    }
    EbeGroup eg;
    eg.nd = src.nd; eg.ne = src.ne; eg.idx = src.d_idx; eg.sign = src.d_sign; eg.ck = src.d_ck;
    double *dke = nullptr;
    for (auto &kv : ke_dev) if (kv.first == src.ke_host) dke = kv.second;
    if (!dke) {
      PCGB_CUDA(cudaMalloc(&dke, (size_t)src.nd * src.nd * sizeof(double)));
      PCGB_CUDA(cudaMemcpy(dke, src.ke_host, (size_t)src.nd * src.nd * sizeof(double), cudaMemcpyHostToDevice));
      ke_dev.push_back({src.ke_host, dke});
    }
    eg.ke = dke;
    bool t24 = false;
    if (src.nd == 24) {
      const int slot = ebe_slot_acquire(src.ke_host);   // shared per-device table: never overwrites a live operator's matrices
      if (slot >= 0) { eg.slot = slot; t24 = true; }
    }
    if (t24) {
      if (open_warp_phase >= 0 && open_warp_phase != phase[g]) flush_warp();
      if (src.ne > 0) E->C.launches.push_back({0, (int)P.groups.size(), 0});
    } else {
      if (open_warp_phase >= 0 && open_warp_phase != phase[g]) flush_warp();
      if (open_warp_phase < 0) { open_warp_phase = phase[g]; open_warp_begin = (int)blk_group.size(); }
      for (int64_t e0 = 0; e0 < src.ne; e0 += kEbeWarpsPerBlock) { blk_group.push_back((int)P.groups.size()); blk_e0.push_back(e0); }
    }
    P.bytes += src.ne * ((int64_t)src.nd * (4 + (src.d_sign ? 1 : 0)) + 8);
    P.groups.push_back(eg);
    E->C.nphases = phase[g] + 1;
  }
  flush_warp();
  // a t24 launch of phase c and a warp launch of the same phase may touch the same dofs only if the colouring was
  // done per pattern group; the Python side colours ALL elements of the subdomain together, so they cannot.
  if (!P.groups.empty()) {
    PCGB_CUDA(cudaMalloc(&P.d_groups, P.groups.size() * sizeof(EbeGroup)));
    PCGB_CUDA(cudaMemcpy(P.d_groups, P.groups.data(), P.groups.size() * sizeof(EbeGroup), cudaMemcpyHostToDevice));
  }
  P.nblk_warp = (int)blk_group.size();
  if (P.nblk_warp > 0) {
    PCGB_CUDA(cudaMalloc(&P.d_blk_group, blk_group.size() * sizeof(int)));
    PCGB_CUDA(cudaMalloc(&P.d_blk_e0, blk_e0.size() * sizeof(int64_t)));
    PCGB_CUDA(cudaMemcpy(P.d_blk_group, blk_group.data(), blk_group.size() * sizeof(int), cudaMemcpyHostToDevice));
    PCGB_CUDA(cudaMemcpy(P.d_blk_e0, blk_e0.data(), blk_e0.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
  }
  *out = E;
  return PCGB_OK;
}

int pcgb_ebe2_destroy(pcgb_ebe2_t E) {
  if (!E) return PCGB_OK;
  std::vector<const double *> freed;
  for (EbeGroup &g : E->C.P.groups) {
    bool done = false;
    for (const double *f : freed) done |= (f == g.ke);
    if (!done) { cudaFree(const_cast<double *>(g.ke)); freed.push_back(g.ke); }
    ebe_slot_release(E->device, g.slot);
  }
  cudaFree(E->C.P.d_groups); cudaFree(E->C.P.d_blk_group); cudaFree(E->C.P.d_blk_e0);
  delete E;
  return PCGB_OK;
}

int pcgb_ebe2_apply(pcgb_ebe2_t E, const double *d_x, double *d_y, void *stream) {
  if (!E || !d_x || !d_y) return fail(PCGB_ERR_ARG, "pcgb_ebe2_apply: null argument");
  return ebe_color_apply(E->C, d_x, d_y, (cudaStream_t)stream);
}

int pcgb_ebe2_launches(pcgb_ebe2_t E) { return E ? (int)E->C.launches.size() : 0; }

// ------------------------------------------------------------------------------------ solver
static int solver_create_common(pcgb_csr_t A, pcgb_ebe_t E, pcgb_halo_t halo, pcgb_comm_t comm, pcgb_solver_t *out) {
  if (halo && !comm) comm = halo->comm;
  PCGB_TRY(require_device());
  pcgb_solver_t s = new pcgb_solver_s();
  s->A = A; s->E = E; s->halo = halo; s->comm = comm; s->n = A ? A->P.nrows : E->P.n;
  s->fork_halo = env_int("PCGB_FORK", 1) != 0;
  const size_t nb = (size_t)(s->n > 0 ? s->n : 1) * sizeof(double);
  s->stage_cap = ((A ? A->P.ntiles : 0) + 4095) / 4096 + 1;
  cudaError_t ce = cudaSuccess;
  int rc = PCGB_OK;
  if ((ce = cudaMalloc(&s->r, nb)) == cudaSuccess && (ce = cudaMalloc(&s->p, nb)) == cudaSuccess &&
      (ce = cudaMalloc(&s->q, nb)) == cudaSuccess && (ce = cudaMalloc(&s->xalt, nb)) == cudaSuccess &&
      (ce = cudaMalloc(&s->xown, nb)) == cudaSuccess && (ce = cudaMemset(s->p, 0, nb)) == cudaSuccess &&
      (ce = cudaMalloc(&s->partials, 5 * kMaxVecGrid * sizeof(double))) == cudaSuccess &&
      (ce = cudaMalloc(&s->stage, (size_t)s->stage_cap * sizeof(double))) == cudaSuccess &&
      (ce = cudaMalloc(&s->red, 8 * sizeof(double))) == cudaSuccess && (ce = cudaMemset(s->red, 0, 8 * sizeof(double))) == cudaSuccess &&
      (ce = cudaMalloc(&s->d_ctrl, sizeof(PcgCtrl))) == cudaSuccess && (ce = cudaMallocHost(&s->h_ctrl, sizeof(PcgCtrl))) == cudaSuccess &&
      (ce = cudaMallocHost(&s->h_red, 8 * sizeof(double))) == cudaSuccess &&
      (ce = cudaStreamCreateWithFlags(&s->own, cudaStreamNonBlocking)) == cudaSuccess &&
      (ce = cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking)) == cudaSuccess &&
      (ce = cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming)) == cudaSuccess &&
      (ce = cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming)) == cudaSuccess &&
      (ce = cudaEventCreateWithFlags(&s->ev_in, cudaEventDisableTiming)) == cudaSuccess &&
      (ce = cudaEventCreateWithFlags(&s->ev_out, cudaEventDisableTiming)) == cudaSuccess &&
      (ce = cudaEventCreate(&s->ev_l0)) == cudaSuccess && (ce = cudaEventCreate(&s->ev_l1)) == cudaSuccess &&
      (ce = cudaEventCreate(&s->ev_s0)) == cudaSuccess)
    ce = cudaEventCreate(&s->ev_s1);
  if (ce != cudaSuccess) rc = fail(PCGB_ERR_CUDA, "pcgb_solver_create: %s", cudaGetErrorString(ce));
  // interface-first tile order: the tiles owning exchanged rows run first, their values travel while the rest computes
  // Opt-in (PCGB_OVERLAP=1): with the peer-store halo the exchange costs a few microseconds and is not worth a second
  // persistent launch (B200 x2, 128^3 per GPU: 1.118 ms/iteration split against 1.106 ms unsplit, profiles/bench_r2b_*).
  if (rc == PCGB_OK && A && halo && halo->ndof > 0 && env_int("PCGB_OVERLAP", 0) != 0)
    rc = spmv_set_boundary_rows(A->P, halo->d_dof, halo->ndof, s->own);
  if (rc != PCGB_OK) { pcgb_solver_destroy(s); return rc; }
  *out = s;
  return PCGB_OK;
}

int pcgb_solver_create(pcgb_csr_t A, pcgb_halo_t halo, pcgb_comm_t comm, pcgb_solver_t *out) {
  if (!A || !out) return fail(PCGB_ERR_ARG, "pcgb_solver_create: null argument");
  if (A->P.nrows != A->P.ncols) return fail(PCGB_ERR_ARG, "pcgb_solver_create: operator must be square");
  return solver_create_common(A, nullptr, halo, comm, out);
}

int pcgb_solver_create_ebe(pcgb_ebe_t E, pcgb_halo_t halo, pcgb_comm_t comm, pcgb_solver_t *out) {
  if (!E || !out) return fail(PCGB_ERR_ARG, "pcgb_solver_create_ebe: null argument");
  return solver_create_common(nullptr, E, halo, comm, out);
}

int pcgb_solver_destroy(pcgb_solver_t s) {
  if (!s) return PCGB_OK;
  if (s->gexec) cudaGraphExecDestroy(s->gexec);
  if (s->own) cudaStreamDestroy(s->own);
  if (s->side) cudaStreamDestroy(s->side);
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  if (s->ev_join) cudaEventDestroy(s->ev_join);
  if (s->ev_in) cudaEventDestroy(s->ev_in);
  if (s->ev_out) cudaEventDestroy(s->ev_out);
  if (s->ev_l0) cudaEventDestroy(s->ev_l0);
  if (s->ev_l1) cudaEventDestroy(s->ev_l1);
  if (s->ev_s0) cudaEventDestroy(s->ev_s0);
  if (s->ev_s1) cudaEventDestroy(s->ev_s1);
  for (cudaEvent_t e : s->ev_k) cudaEventDestroy(e);
  cudaFree(s->r); cudaFree(s->p); cudaFree(s->q); cudaFree(s->xalt); cudaFree(s->xown); cudaFree(s->partials); cudaFree(s->stage);
  cudaFree(s->red); cudaFree(s->d_ctrl);
  if (s->h_ctrl) cudaFreeHost(s->h_ctrl);
  if (s->h_red) cudaFreeHost(s->h_red);
  delete s;
  return PCGB_OK;
}

}  // extern "C"

// ---- host helpers of the solve -------------------------------------------------------------
namespace {

inline bool multi_rank(pcgb_solver_t s) { return s->comm && s->comm->nranks > 1; }
inline bool peer_path(pcgb_solver_t s) { return multi_rank(s) && s->comm->use_peer() && (!s->halo || s->halo->m == 0 || s->halo->use_peer()); }

// y = A x + interface sum  (calcMPFint, pcg_solver.py:339-342)
int op_apply(pcgb_solver_t s, const double *x, double *y, cudaStream_t st) {
  if (s->E) PCGB_TRY(ebe_apply(s->E->P, x, y, st, &s->launches));
  else PCGB_TRY(spmv_launch(s->A->P, x, y, false, st, &s->launches));
  if (s->halo) PCGB_TRY(halo_exchange_add(s->halo, y, st, &s->launches));
  return PCGB_OK;
}

// sum `nv` (<= 2) rows of s->partials over blocks and ranks into the device scalars s->red[dst .. dst+nv)
int reduce_to_device(pcgb_solver_t s, int count, int nv, int dst, cudaStream_t st) {
  PCGB_TRY(reduce_rows(s->partials, count, nv, s->red + dst, st));
  s->launches += 1;
  if (multi_rank(s)) {
    PCGB_TRY(allreduce_sum(s->comm, s->red + dst, nv, st));
    if (s->comm->use_peer()) s->launches += 1;
  }
  return PCGB_OK;
}

// q = A x ; r = b - q ; sum r*r*w over all ranks -> s->red[dst] (device)
int true_residual_dev(pcgb_solver_t s, const double *b, const double *w, const double *x, int dst, cudaStream_t st) {
  PCGB_TRY(op_apply(s, x, s->q, st));
  const int grid = vec_grid(s->n);
  k_residual<<<grid, kVecBlock, 0, st>>>(s->n, b, s->q, w, s->r, s->partials);
  PCGB_CHECK_LAUNCH();
  s->launches += 1;
  return reduce_to_device(s, grid, 1, dst, st);
}

// host copy of s->red[6..8) (one synchronisation)
int fetch_red(pcgb_solver_t s, cudaStream_t st) {
  PCGB_CUDA(cudaMemcpyAsync(s->h_red, s->red + 6, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  PCGB_CUDA(cudaStreamSynchronize(st));
  return PCGB_OK;
}

// r = b - A x ; returns sqrt(sum r*r*w) over all ranks
int true_residual(pcgb_solver_t s, const double *b, const double *w, const double *x, double *normr, cudaStream_t st) {
  PCGB_TRY(true_residual_dev(s, b, w, x, 6, st));
  PCGB_TRY(fetch_red(s, st));
  *normr = std::sqrt(s->h_red[0]);
  return PCGB_OK;
}

// (sum z.(r w), #inf z) of the current r into s->red[6..7] (device, all-reduced)
int rz_to_device(pcgb_solver_t s, const double *minv, const double *w, cudaStream_t st) {
  const int grid = vec_grid(s->n);
  k_rz<<<grid, kVecBlock, 0, st>>>(s->n, s->r, minv, w, s->partials);
  PCGB_CHECK_LAUNCH();
  s->launches += 1;
  return reduce_to_device(s, grid, 2, 6, st);
}

// one PCG iteration enqueued on st (pcg_solver.py:438-562); every kernel no-ops once the state is frozen.
// ev[0..kEvPerIter): optional events at the phase boundaries of this iteration (time_kernels):
//   0 start | 1 after p-update | 2 after SpMV (interface tiles when split) | 3 after halo pack | 4 after the interior tiles |
//   5 after the p.q reduction + all-reduce + alpha | 6 after halo unpack-add | 7 after the fused update | 8 after the norms reduction
constexpr int kEvPerIter = 9;
int enqueue_iteration(pcgb_solver_t s, const double *minv, const double *w, double *xb0, double *resvec, cudaStream_t st, int *nl,
                      cudaEvent_t *ev = nullptr) {
  const int64_t n = s->n;
  const int vg = vec_grid(n);
  const bool multi = multi_rank(s);
  const bool peer = peer_path(s);
#define PCGB_MARK(k) do { if (ev) PCGB_CUDA(cudaEventRecord(ev[k], st)); } while (0)
  PCGB_MARK(0);
  k_pupdate<<<vg, kVecBlock, 0, st>>>(s->d_ctrl, n, s->r, minv, s->p);
  PCGB_CHECK_LAUNCH();
  *nl += 1;
  PCGB_MARK(1);
  // p.q : per-tile partials -> (stage) -> scalar.  In the multi-GPU case the partials are those of the
  // UNASSEMBLED local product, whose rank sum equals the reference's weighted dot of the assembled q
  // (p is consistent on shared dofs and K = sum of the subdomain matrices); see DESIGN.md.
  const double *pq_src;
  int pq_cnt;
  bool halo_done = false;
  if (s->E) {  // matrix-free operator: product, then a separate unweighted dot of the local product
    PCGB_TRY(ebe_apply(s->E->P, s->p, s->q, st, nl, &s->d_ctrl->state));
    PCGB_MARK(2); PCGB_MARK(3); PCGB_MARK(4);
    k_dot_w<<<vg, kVecBlock, 0, st>>>(n, s->p, s->q, nullptr, s->partials);
    PCGB_CHECK_LAUNCH();
    *nl += 1;
    pq_src = s->partials; pq_cnt = vg;
  } else {
    const CsrPlan &P = s->A->P;
    if (peer && s->halo && s->halo->m > 0 && spmv_split_available(P)) {
      // interface tiles -> pack (peer stores fly over NVLink) -> interior tiles; the unpack follows the p.q all-reduce
      PCGB_TRY(spmv_launch_part(P, 0, s->p, s->q, true, st, nl, &s->d_ctrl->state));
      PCGB_MARK(2);
      PCGB_TRY(halo_pack(s->halo, s->q, st, nl));
      PCGB_MARK(3);
      PCGB_TRY(spmv_launch_part(P, 1, s->p, s->q, true, st, nl, &s->d_ctrl->state));
      PCGB_MARK(4);
      pq_src = P.dot_partials_split; pq_cnt = spmv_split_dot_count(P);
      halo_done = true;
    } else {
      PCGB_TRY(spmv_launch(P, s->p, s->q, true, st, nl, &s->d_ctrl->state));
      PCGB_MARK(2);
      if (peer && s->halo) { PCGB_TRY(halo_pack(s->halo, s->q, st, nl)); halo_done = true; }
      PCGB_MARK(3); PCGB_MARK(4);
      pq_src = P.dot_partials;
      pq_cnt = P.persist ? P.grid_persist : P.ntiles;
      if (pq_cnt > 8192) {
        const int sb = (pq_cnt + 4095) / 4096;
        k_stage_reduce<<<sb, 256, 0, st>>>(P.dot_partials, pq_cnt, s->stage);
        PCGB_CHECK_LAUNCH();
        *nl += 1;
        pq_src = s->stage; pq_cnt = sb;
      }
    }
  }
  if (peer) {
    // own kernels over peer memory: the all-reduce is fused into the reduction kernel, the scalar logic follows in the same CTA
    if (s->halo && !halo_done) PCGB_TRY(halo_pack(s->halo, s->q, st, nl));
    const PeerWin win = s->comm->window();
    // fork: the unpack-add of the interface values (waits for the neighbours' stores) runs beside the p.q all-reduce (waits
    // for the other ranks' sums); both only need the pack above, which never waits - so no order of execution can deadlock
    const bool fork = s->halo && s->halo->m > 0 && s->side != nullptr && s->fork_halo;
    if (fork) {
      PCGB_CUDA(cudaEventRecord(s->ev_fork, st));
      PCGB_CUDA(cudaStreamWaitEvent(s->side, s->ev_fork, 0));
      PCGB_TRY(halo_unpack(s->halo, s->q, s->side, nl));
      PCGB_CUDA(cudaEventRecord(s->ev_join, s->side));
    }
    k_reduce_ar<1, 1><<<1, 256, 0, st>>>(win, s->d_ctrl, pq_src, pq_cnt, 0, s->red, nullptr);
    PCGB_CHECK_LAUNCH();
    *nl += 1;
    PCGB_MARK(5);
    if (fork) PCGB_CUDA(cudaStreamWaitEvent(st, s->ev_join, 0));
    else if (s->halo) PCGB_TRY(halo_unpack(s->halo, s->q, st, nl));
    PCGB_MARK(6);
    k_update<<<vg, kVecBlock, 0, st>>>(s->d_ctrl, n, s->r, s->q, s->p, minv, w, xb0, s->xalt, s->partials);
    PCGB_CHECK_LAUNCH();
    PCGB_MARK(7);
    k_reduce_ar<5, 2><<<1, 256, 0, st>>>(win, s->d_ctrl, s->partials, vg, kMaxVecGrid, s->red + 1, resvec);
    PCGB_CHECK_LAUNCH();
    *nl += 2;
    PCGB_MARK(8);
    return PCGB_OK;
  }
  if (!multi) {
    k_reduce<1, 1><<<1, 256, 0, st>>>(s->d_ctrl, pq_src, pq_cnt, 0, s->red, nullptr);
    PCGB_CHECK_LAUNCH();
    *nl += 1;
    PCGB_MARK(5); PCGB_MARK(6);
  } else {
    // NCCL transport: reduction, ncclAllReduce and the scalar logic are three stream operations; the halo is serial
    k_reduce<1, 0><<<1, 256, 0, st>>>(s->d_ctrl, pq_src, pq_cnt, 0, s->red, nullptr);
    PCGB_CHECK_LAUNCH();
    PCGB_TRY(allreduce_sum(s->comm, s->red, 1, st));
    k_ctrl_alpha<<<1, 1, 0, st>>>(s->d_ctrl, s->red);
    PCGB_CHECK_LAUNCH();
    *nl += 2;
    PCGB_MARK(5);
    if (s->halo) PCGB_TRY(halo_exchange_add(s->halo, s->q, st, nl));
    PCGB_MARK(6);
  }
  k_update<<<vg, kVecBlock, 0, st>>>(s->d_ctrl, n, s->r, s->q, s->p, minv, w, xb0, s->xalt, s->partials);
  PCGB_CHECK_LAUNCH();
  *nl += 1;
  PCGB_MARK(7);
  if (!multi) {
    k_reduce<5, 2><<<1, 256, 0, st>>>(s->d_ctrl, s->partials, vg, kMaxVecGrid, s->red + 1, resvec);
    PCGB_CHECK_LAUNCH();
    *nl += 1;
  } else {
    k_reduce<5, 0><<<1, 256, 0, st>>>(s->d_ctrl, s->partials, vg, kMaxVecGrid, s->red + 1, nullptr);
    PCGB_CHECK_LAUNCH();
    PCGB_TRY(allreduce_sum(s->comm, s->red + 1, 5, st));
    k_ctrl_norms<<<1, 1, 0, st>>>(s->d_ctrl, s->red + 1, resvec);
    PCGB_CHECK_LAUNCH();
    *nl += 2;
  }
  PCGB_MARK(8);
#undef PCGB_MARK
  return PCGB_OK;
}

int fetch_ctrl(pcgb_solver_t s, cudaStream_t st) {
  PCGB_CUDA(cudaMemcpyAsync(s->h_ctrl, s->d_ctrl, sizeof(PcgCtrl), cudaMemcpyDeviceToHost, st));
  if (s->comm && s->comm->d_status) PCGB_CUDA(cudaMemcpyAsync(s->comm->h_status, s->comm->d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
  PCGB_CUDA(cudaStreamSynchronize(st));
  if (s->comm && s->comm->d_status && *s->comm->h_status != 0)
    return fail(PCGB_ERR_COMM, "a peer-memory exchange timed out waiting for another rank (rank %d of %d)", s->comm->rank, s->comm->nranks);
  return PCGB_OK;
}

}  // namespace

extern "C" {

int pcgb_apply(pcgb_solver_t s, const double *d_x, double *d_y, void *stream) {
  if (!s || !d_x || !d_y) return fail(PCGB_ERR_ARG, "pcgb_apply: null argument");
  return op_apply(s, d_x, d_y, (cudaStream_t)stream);
}

static int solve_on(pcgb_solver_t s, const double *d_b, const double *d_minv, const double *d_w, double *d_x,
                    const pcgb_options *opt, double *d_resvec, pcgb_result *res, cudaStream_t st);

int pcgb_solve(pcgb_solver_t s, const double *d_b, const double *d_minv, const double *d_w, double *d_x,
               const pcgb_options *opt, double *d_resvec, pcgb_result *res, void *stream) {
  if (!s || !d_b || !d_x || !opt || !res) return fail(PCGB_ERR_ARG, "pcgb_solve: null argument");
  if (opt->maxiter <= 0) return fail(PCGB_ERR_ARG, "pcgb_solve: maxiter must be positive");
  // order the solver's own stream after the caller's stream, run, and order the caller's stream after us
  cudaStream_t user = (cudaStream_t)stream;
  PCGB_CUDA(cudaEventRecord(s->ev_in, user));
  PCGB_CUDA(cudaStreamWaitEvent(s->own, s->ev_in, 0));
  const int rc = solve_on(s, d_b, d_minv, d_w, d_x, opt, d_resvec, res, s->own);
  if (rc == PCGB_OK) {
    PCGB_CUDA(cudaEventRecord(s->ev_out, s->own));
    PCGB_CUDA(cudaStreamWaitEvent(user, s->ev_out, 0));
  }
  return rc;
}

static int solve_on(pcgb_solver_t s, const double *d_b, const double *d_minv, const double *d_w, double *d_x,
                    const pcgb_options *opt, double *d_resvec, pcgb_result *res, cudaStream_t st) {
  const int64_t n = s->n;
  const int vg = vec_grid(n);
  memset(res, 0, sizeof(*res));
  s->launches = 0;
  int64_t matvecs = 0, graph_launch_kernels = 0;
  const int maxstag = opt->max_stag > 0 ? opt->max_stag : 3;
  const int64_t nglob = opt->n_global > 0 ? opt->n_global : n;
  // the loop works on the solver's own pair of x buffers (graph / kernel arguments never change between solves)
  double *const xw = s->xown;
  double *xbuf[2] = {xw, s->xalt};
  PCGB_CUDA(cudaEventRecord(s->ev_s0, st));

  // ---- ||b||  (pcg_solver.py:381-384) and the initial residual (:408-418) with ONE host synchronisation:
  //      sum b*b*w -> red[7], r = b - A x0 and sum r*r*w -> red[6]
  k_dot_w<<<vg, kVecBlock, 0, st>>>(n, d_b, d_b, d_w, s->partials);
  PCGB_CHECK_LAUNCH();
  s->launches += 1;
  PCGB_TRY(reduce_to_device(s, vg, 1, 7, st));
  if (opt->x0_zero) {
    // the caller states x0 = 0: A x0 = 0 exactly, hence r = b and ||r|| = ||b|| bit for bit - no operator application
    PCGB_CUDA(cudaMemsetAsync(xw, 0, (size_t)n * sizeof(double), st));
    PCGB_CUDA(cudaMemcpyAsync(s->r, d_b, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    PCGB_CUDA(cudaMemcpyAsync(s->red + 6, s->red + 7, sizeof(double), cudaMemcpyDeviceToDevice, st));
  } else {
    PCGB_CUDA(cudaMemcpyAsync(xw, d_x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    PCGB_TRY(true_residual_dev(s, d_b, d_w, xw, 6, st));
    ++matvecs;
  }
  if (d_resvec) {  // ResVec[0] = sqrt(red[6]) (:431), computed on the device
    k_sqrt_store<<<1, 1, 0, st>>>(s->red + 6, d_resvec);
    PCGB_CHECK_LAUNCH();
  }
  PCGB_TRY(fetch_red(s, st));
  const double n2b = std::sqrt(s->h_red[1]);
  const double normr = std::sqrt(s->h_red[0]);
  const double tolb = opt->tol * n2b;
  res->normb = n2b;
  if (n2b == 0.0) {  // :387-395 - returns the initial guess, flag 0, relres 0, iter 0
    res->flag = 0; res->relres = 0.0; res->iters = 0; res->launches = s->launches;
    return PCGB_OK;
  }
  if (!opt->fixed_iters && normr <= tolb) {  // :421-426
    res->flag = 0; res->relres = normr / n2b; res->iters = 0; res->matvecs = matvecs; res->launches = s->launches;
    return PCGB_OK;
  }
  // ---- loop state (:399-406)
  int64_t mm = nglob / 50;
  if (mm > 5) mm = 5;
  if (nglob - opt->maxiter < mm) mm = nglob - opt->maxiter;
  const int64_t maxmsteps = mm;
  PcgCtrl c;
  memset(&c, 0, sizeof(c));
  c.rho = 1.0; c.rho_prev = 1.0; c.alpha = 0.0; c.beta = 0.0;
  c.normr = normr; c.normr_act = normr; c.normrmin = normr;
  c.tolb = tolb; c.n2b = n2b; c.eps = 2.220446049250313e-16;
  c.state = ST_RUN; c.flag = 1; c.iter = 0; c.stag = 0; c.moresteps = 0; c.imin = 0; c.xcur = 0; c.xmin = 0;
  c.alias = 1;   // MP_XMin = MP_X (:379-380): the same array until the first improvement is recorded
  c.maxiter = opt->maxiter; c.maxstag = maxstag; c.fixed_iters = opt->fixed_iters ? 1 : 0;
  *s->h_ctrl = c;
  // (pinned source, stream-ordered: the next write to h_ctrl is the D2H copy of fetch_ctrl on the same stream)
  PCGB_CUDA(cudaMemcpyAsync(s->d_ctrl, s->h_ctrl, sizeof(PcgCtrl), cudaMemcpyHostToDevice, st));
  // head of iteration 0: rho = z.r (:446-469)
  PCGB_TRY(rz_to_device(s, d_minv, d_w, st));
  k_ctrl_head<<<1, 1, 0, st>>>(s->d_ctrl, s->red + 6, 0);
  PCGB_CHECK_LAUNCH();
  s->launches += 1;

  int batch = opt->check_every > 0 ? opt->check_every : 16;
  if (batch > opt->maxiter) batch = opt->maxiter;
  const bool want_graph = opt->use_graph != 0 && !opt->time_kernels;
  size_t ktimed = 0;           // iterations whose SpMV launches carry event brackets
  PCGB_CUDA(cudaEventRecord(s->ev_l0, st));
  int flag = 1;
  int too_small = 0;

  for (;;) {
    // ---- enqueue `batch` iterations
    if (want_graph && batch > 1) {
      const int variant = (peer_path(s) ? 1 : 0) | ((s->A && spmv_split_available(s->A->P)) ? 2 : 0) | (s->fork_halo ? 4 : 0);
      GraphKey key{d_minv, d_w, xw, d_resvec, batch, variant};
      if (!s->gexec || !(s->gkey == key)) {
        if (s->gexec) { cudaGraphExecDestroy(s->gexec); s->gexec = nullptr; }
        cudaGraph_t graph = nullptr;
        PCGB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        int rc = PCGB_OK, nl = 0;
        for (int k = 0; k < batch && rc == PCGB_OK; ++k) rc = enqueue_iteration(s, d_minv, d_w, xw, d_resvec, st, &nl);
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc != PCGB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        PCGB_CUDA(ce);
        PCGB_CUDA(cudaGraphInstantiate(&s->gexec, graph, 0));
        cudaGraphDestroy(graph);
        s->gkey = key;
        s->launches_per_iter = nl / batch;
      }
      PCGB_CUDA(cudaGraphLaunch(s->gexec, st));
      graph_launch_kernels += (int64_t)s->launches_per_iter * batch;
    } else {
      for (int k = 0; k < batch; ++k) {
        int nl = 0;
        cudaEvent_t *ev = nullptr;
        if (opt->time_kernels && ktimed < 1024) {
          while (s->ev_k.size() < (size_t)kEvPerIter * (ktimed + 1)) { cudaEvent_t e; PCGB_CUDA(cudaEventCreate(&e)); s->ev_k.push_back(e); }
          ev = &s->ev_k[(size_t)kEvPerIter * ktimed]; ++ktimed;
        }
        PCGB_TRY(enqueue_iteration(s, d_minv, d_w, xw, d_resvec, st, &nl, ev));
        s->launches += nl;
      }
    }
    PCGB_TRY(fetch_ctrl(s, st));
    c = *s->h_ctrl;
    if (c.state == ST_RUN) continue;
    if (c.state == ST_TRIGGER) {
      // ---- verification with the true residual (:527-552)
      double normr_act = 0.0;
      PCGB_TRY(true_residual(s, d_b, d_w, xbuf[c.xcur], &normr_act, st));
      ++matvecs;
      c.normr_act = normr_act;
      if (normr_act <= tolb) { flag = 0; c.flag = 0; break; }  // :540-543
      if (c.stag >= maxstag && c.moresteps == 0) c.stag = 0;   // :545
      c.moresteps += 1;                                         // :546
      if (c.moresteps >= maxmsteps) {                           // :548-552 (the reference raises here)
        too_small = 1; flag = 3; c.flag = 3; break;
      }
      if (normr_act < c.normrmin) { c.normrmin = normr_act; c.xmin = c.xcur; c.imin = c.iter; c.alias = 0; }  // :555-558
      if (c.stag >= maxstag) { flag = 3; c.flag = 3; break; }  // :560-562
      // continue with the replaced residual: head of the next iteration
      c.state = ST_RUN;
      *s->h_ctrl = c;
      PCGB_CUDA(cudaMemcpyAsync(s->d_ctrl, s->h_ctrl, sizeof(PcgCtrl), cudaMemcpyHostToDevice, st));
      PCGB_TRY(rz_to_device(s, d_minv, d_w, st));
      k_ctrl_head<<<1, 1, 0, st>>>(s->d_ctrl, s->red + 6, 1);
      PCGB_CHECK_LAUNCH();
      s->launches += 1;
      batch = 1;  // MoreSteps > 0: every further iteration is verified (:527)
      continue;
    }
    flag = c.flag;  // ST_BREAK (2,3,4) or ST_EXHAUSTED (1)
    break;
  }
  PCGB_CUDA(cudaEventRecord(s->ev_l1, st));

  // ---- finalisation (:566-584)
  const int i = c.iter;
  int iter_out;
  double relres;
  int xout;
  if (flag == 0) {
    relres = c.normr_act / n2b;
    iter_out = i;
    xout = c.xcur;
  } else {
    // XMin is the current iterate while it is still aliased to X (no improvement recorded yet): c.xmin == c.xcur then
    double normr_min = 0.0;
    PCGB_TRY(true_residual(s, d_b, d_w, xbuf[c.xmin], &normr_min, st));
    ++matvecs;
    if (normr_min < c.normr_act) { iter_out = c.imin; relres = normr_min / n2b; }
    else { iter_out = i; relres = c.normr_act / n2b; }
    xout = c.xmin;  // the reference exports XMin on this path in both cases (:569, :598)
  }
  iter_out += 1;  // :584
  PCGB_CUDA(cudaMemcpyAsync(d_x, xbuf[xout], (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  PCGB_CUDA(cudaEventRecord(s->ev_s1, st));
  PCGB_CUDA(cudaEventSynchronize(s->ev_s1));
  {
    float ms = 0.f;
    PCGB_CUDA(cudaEventElapsedTime(&ms, s->ev_l0, s->ev_l1));
    res->loop_ms = ms;
    res->loop_iters = c.iter + 1;
    PCGB_CUDA(cudaEventElapsedTime(&ms, s->ev_s0, s->ev_l0));
    res->setup_ms = ms;
    PCGB_CUDA(cudaEventElapsedTime(&ms, s->ev_l1, s->ev_s1));
    res->final_ms = ms;
    // phase sums over the bracketed iterations: p-update | SpMV | halo pack | p.q reduce + all-reduce | halo unpack |
    // fused update | norms reduce + all-reduce   (phase_ms[7] = their total)
    double tot = 0.0;
    static const int seg[7][2] = {{0, 1}, {1, 2}, {2, 3}, {4, 5}, {5, 6}, {6, 7}, {7, 8}};
    for (size_t k = 0; k < ktimed; ++k) {
      const cudaEvent_t *e = &s->ev_k[(size_t)kEvPerIter * k];
      float t = 0.f;
      for (int ph = 0; ph < 7; ++ph) {
        PCGB_CUDA(cudaEventElapsedTime(&t, e[seg[ph][0]], e[seg[ph][1]]));
        res->phase_ms[ph] += t;
      }
      PCGB_CUDA(cudaEventElapsedTime(&t, e[3], e[4]));   // interior tiles of a split SpMV (zero-length otherwise)
      res->phase_ms[1] += t;
      PCGB_CUDA(cudaEventElapsedTime(&t, e[0], e[8]));
      res->phase_ms[7] += t;
    }
    tot = res->phase_ms[1];
    res->spmv_ms = tot; res->spmv_timed = (int64_t)ktimed;
  }
  res->flag = flag; res->iters = iter_out; res->relres = relres; res->imin = c.imin; res->stag = c.stag;
  res->moresteps = c.moresteps; res->too_small_tol = too_small;
  // matvecs inside the loop = iterations started
  res->matvecs = matvecs + (int64_t)(c.iter + 1);
  res->launches = s->launches + graph_launch_kernels;
  return PCGB_OK;
}

// ------------------------------------------------------------------------------------ hex generator
int64_t pcgb_hex_nrows(const pcgb_hex_box *box) {
  if (!box) return 0;
  return 3 * hex_geom(box).nfree_nodes;
}

int pcgb_hex_count(const pcgb_hex_box *box, int64_t *d_rowcount, void *stream) {
  if (!box || !d_rowcount) return fail(PCGB_ERR_ARG, "pcgb_hex_count: null argument");
  PCGB_TRY(require_device());
  HexGeom g = hex_geom(box);
  if (g.nx <= 0 || g.ny <= 0 || g.nz <= 0 || g.nfree_nodes <= 0) return fail(PCGB_ERR_ARG, "pcgb_hex_count: empty box");
  k_hex_count<<<(unsigned)((g.nfree_nodes + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, d_rowcount);
  PCGB_CHECK_LAUNCH();
  return PCGB_OK;
}

int pcgb_hex_fill(const pcgb_hex_box *box, const double *ke_host, double ck, const int64_t *d_rowptr, int32_t *d_col,
                  double *d_val, void *stream) {
  if (!box || !ke_host || !d_rowptr || !d_col || !d_val) return fail(PCGB_ERR_ARG, "pcgb_hex_fill: null argument");
  PCGB_TRY(require_device());
  HexGeom g = hex_geom(box);
  PCGB_CUDA(cudaMemcpyToSymbolAsync(c_hex_ke, ke_host, 24 * 24 * sizeof(double), 0, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  const int64_t threads = g.nfree_nodes * 32;
  k_hex_fill<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, ck, d_rowptr, d_col, d_val);
  PCGB_CHECK_LAUNCH();
  return PCGB_OK;
}

}  // extern "C"
