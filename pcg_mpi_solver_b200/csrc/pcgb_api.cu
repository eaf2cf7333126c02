// pcgb_api.cu - C ABI of libpcgb200.so (see include/pcgb200.h) and the host side of the PCG loop.
// sm_100a only.  No CPU fallback: every compute entry point requires a CUDA device.
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "spmv.cuh"
#include "pcg_kernels.cuh"
#include "comm.cuh"
#include "hexgen.cuh"
#include "ebe.cuh"
#include "ebe_color.cuh"

using namespace pcgb;

struct pcgb_csr_s {
  CsrPlan P;
};

struct pcgb_ebe_s {
  EbePlan P;
};

struct pcgb_ebe2_s {
  EbeColorPlan C;
};

struct GraphKey {
  const void *minv, *w, *xb0, *resvec;
  int iters;
  bool operator==(const GraphKey &o) const {
    return minv == o.minv && w == o.w && xb0 == o.xb0 && resvec == o.resvec && iters == o.iters;
  }
};

struct pcgb_solver_s {
  pcgb_csr_t A = nullptr;
  pcgb_ebe_t E = nullptr;       // experimental matrix-free operator (exactly one of A / E is set)
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
  cudaEvent_t ev_in = nullptr, ev_out = nullptr, ev_l0 = nullptr, ev_l1 = nullptr;
  std::vector<cudaEvent_t> ev_k;  // SpMV brackets (time_kernels)
  cudaGraphExec_t gexec = nullptr;
  GraphKey gkey{nullptr, nullptr, nullptr, nullptr, 0};
  int launches = 0;             // launches issued outside graphs (running counter per solve)
  int launches_per_iter = 0;
};

static int require_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(PCGB_ERR_NODEVICE, "no CUDA device visible: libpcgb200 has no CPU fallback");
  }
  return PCGB_OK;
}

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

int pcgb_csr_diag(pcgb_csr_t A, double *d_diag, void *stream) {
  if (!A || !d_diag) return fail(PCGB_ERR_ARG, "pcgb_csr_diag: null argument");
  const CsrPlan &P = A->P;
  if (P.nrows == 0) return PCGB_OK;
  const unsigned grid = (unsigned)((P.nrows * 8 + 255) / 256);
  if (P.rp64) k_csr_diag<int64_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const int64_t *)P.rowptr, P.col, P.val, P.nrows, d_diag);
  else k_csr_diag<int32_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const int32_t *)P.rowptr, P.col, P.val, P.nrows, d_diag);
  PCGB_CHECK_LAUNCH();
  return PCGB_OK;
}

int64_t pcgb_spmv_bytes(pcgb_csr_t A) {
  if (!A) return 0;
  const CsrPlan &P = A->P;
  return 12 * P.nnz + (P.rp64 ? 8 : 4) * (P.nrows + 1) + 8 * P.ncols + 8 * P.nrows;
}

/* bytes the selected kernel actually streams from HBM per SpMV (staged-x: 10 B per non-zero + window tables) */
int64_t pcgb_spmv_stream_bytes(pcgb_csr_t A) {
  if (!A) return 0;
  const CsrPlan &P = A->P;
  const int64_t per = P.staged ? 10 : 12;
  return per * P.nnz + (P.rp64 ? 8 : 4) * (P.nrows + 1) + 8 * P.ncols + 8 * P.nrows + (P.staged ? 8 * P.nwin + 8 * (int64_t)P.ntiles : 0) + 12 * (int64_t)P.ntiles;
}

int pcgb_csr_plan_info(pcgb_csr_t A, int64_t info[12]) {
  if (!A || !info) return fail(PCGB_ERR_ARG, "pcgb_csr_plan_info: null argument");
  const CsrPlan &P = A->P;
  info[0] = P.ntiles; info[1] = P.tile_items; info[2] = P.lanes; info[3] = P.snap ? 1 : 0;
  info[4] = P.nfix; info[5] = P.staged ? P.smem_staged : P.smem_bytes; info[6] = P.max_row; info[7] = P.use_tma ? 1 : 0;
  info[8] = P.persist ? 2 : (P.staged ? 1 : 0); info[9] = P.nwin; info[10] = P.cap_x; info[11] = P.max_nw;
  if (P.persist) info[5] = P.smem_persist;
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
  static thread_local double *scratch = nullptr;
  if (!scratch) PCGB_CUDA(cudaMalloc(&scratch, kMaxVecGrid * sizeof(double)));
  const int grid = vec_grid(n);
  k_dot_w<<<grid, kVecBlock, 0, st>>>(n, d_a, d_b, d_w, scratch);
  PCGB_CHECK_LAUNCH();
  return reduce_rows(scratch, grid, 1, d_out, st);
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

int pcgb_comm_create(int rank, int nranks, const unsigned char id[PCGB_UNIQUE_ID_BYTES], pcgb_comm_t *out) {
  if (!out || !id || rank < 0 || rank >= nranks) return fail(PCGB_ERR_ARG, "pcgb_comm_create: bad argument");
  PCGB_TRY(require_device());
  NcclApi *api = nullptr;
  PCGB_TRY(nccl_api(&api));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  pcgb_comm_t c = new pcgb_comm_s();
  c->api = api; c->rank = rank; c->nranks = nranks;
  int e = api->CommInitRank(&c->comm, nranks, uid, rank);
  if (e != ncclSuccess) {
    delete c;
    return fail(PCGB_ERR_NCCL, "ncclCommInitRank -> %s", api->GetErrorString ? api->GetErrorString(e) : "error");
  }
  *out = c;
  return PCGB_OK;
}

int pcgb_comm_destroy(pcgb_comm_t c) {
  if (!c) return PCGB_OK;
  if (c->comm && c->api && c->api->CommDestroy) c->api->CommDestroy(c->comm);
  delete c;
  return PCGB_OK;
}

int pcgb_allreduce_sum(pcgb_comm_t c, double *d_buf, int count, void *stream) {
  if (!c || !d_buf || count < 0) return fail(PCGB_ERR_ARG, "pcgb_allreduce_sum: bad argument");
  PCGB_NCCL(c->api, c->api->AllReduce(d_buf, d_buf, (size_t)count, ncclFloat64, ncclSum, c->comm, (cudaStream_t)stream));
  return PCGB_OK;
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
  if (h->m > INT32_MAX) { delete h; return fail(PCGB_ERR_ARG, "pcgb_halo_create: too many shared dofs"); }
  if (h->m > 0) {
    std::vector<int> idx((size_t)h->m);
    for (int64_t k = 0; k < h->m; ++k) {
      if (idx_host[k] < 0 || idx_host[k] >= nlocal) { delete h; return fail(PCGB_ERR_ARG, "pcgb_halo_create: index %lld out of range", (long long)idx_host[k]); }
      idx[(size_t)k] = (int)idx_host[k];
    }
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
    auto up = [&](int **d, const std::vector<int> &v) -> cudaError_t {
      cudaError_t e = cudaMalloc(d, v.size() * sizeof(int));
      if (e != cudaSuccess) return e;
      return cudaMemcpy(*d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice);
    };
    PCGB_CUDA(up(&h->d_idx, idx));
    PCGB_CUDA(up(&h->d_dof, dof));
    PCGB_CUDA(up(&h->d_ptr, ptr));
    PCGB_CUDA(up(&h->d_pos, pos));
    PCGB_CUDA(cudaMalloc(&h->d_send, (size_t)h->m * sizeof(double)));
    PCGB_CUDA(cudaMalloc(&h->d_recv, (size_t)h->m * sizeof(double)));
  }
  *out = h;
  return PCGB_OK;
}

int pcgb_halo_destroy(pcgb_halo_t h) {
  if (!h) return PCGB_OK;
  cudaFree(h->d_idx); cudaFree(h->d_dof); cudaFree(h->d_ptr); cudaFree(h->d_pos); cudaFree(h->d_send); cudaFree(h->d_recv);
  delete h;
  return PCGB_OK;
}

int pcgb_halo_exchange_add(pcgb_halo_t h, double *d_y, void *stream) {
  if (!h || !d_y) return fail(PCGB_ERR_ARG, "pcgb_halo_exchange_add: null argument");
  return halo_exchange_add(h, d_y, (cudaStream_t)stream);
}

int64_t pcgb_halo_bytes(pcgb_halo_t h) { return h ? h->m * 8 : 0; }

// ------------------------------------------------------------------------------------ EBE operator (experimental)
int pcgb_ebe_create(int64_t n, int ngroups, const pcgb_ebe_group *groups, pcgb_ebe_t *out) {
  if (!out || n < 0 || ngroups < 0 || (ngroups > 0 && !groups)) return fail(PCGB_ERR_ARG, "pcgb_ebe_create: bad argument");
  if (n >= (1 << 30)) return fail(PCGB_ERR_ARG, "pcgb_ebe_create: more than 2^30 dofs");
  PCGB_TRY(require_device());
  pcgb_ebe_t E = new pcgb_ebe_s();
  EbePlan &P = E->P;
  P.n = n;
  int slots = 0;
  std::vector<int> blk_group;
  std::vector<int64_t> blk_e0;
  P.bytes = 16 * n;
  for (int g = 0; g < ngroups; ++g) {
    const pcgb_ebe_group &src = groups[g];
    if (src.nd <= 0 || src.nd > 96 || src.ne < 0 || !src.ke_host || (src.ne > 0 && (!src.d_idx || !src.d_ck))) {
      delete E;
      return fail(PCGB_ERR_ARG, "pcgb_ebe_create: group %d: pattern size must be 1..96 and arrays non-null", g);
    }
    EbeGroup eg;
    eg.nd = src.nd; eg.ne = src.ne; eg.idx = src.d_idx; eg.sign = src.d_sign; eg.ck = src.d_ck;
    double *dke = nullptr;
    PCGB_CUDA(cudaMalloc(&dke, (size_t)src.nd * src.nd * sizeof(double)));
    PCGB_CUDA(cudaMemcpy(dke, src.ke_host, (size_t)src.nd * src.nd * sizeof(double), cudaMemcpyHostToDevice));
    eg.ke = dke;
    if (src.nd == 24 && slots < kEbeMaxSlots) {
      PCGB_CUDA(cudaMemcpyToSymbol(c_ebe_ke24, src.ke_host, 576 * sizeof(double), (size_t)slots * 576 * sizeof(double)));
      eg.slot = slots++;
    } else {
      for (int64_t e0 = 0; e0 < src.ne; e0 += kEbeWarpsPerBlock) { blk_group.push_back((int)P.groups.size()); blk_e0.push_back(e0); }
    }
    P.bytes += src.ne * ((int64_t)src.nd * (4 + (src.d_sign ? 1 : 0)) + 8);
    P.groups.push_back(eg);
  }
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

int pcgb_ebe_destroy(pcgb_ebe_t E) {
  if (!E) return PCGB_OK;
  for (EbeGroup &g : E->P.groups) cudaFree(const_cast<double *>(g.ke));
  cudaFree(E->P.d_groups); cudaFree(E->P.d_blk_group); cudaFree(E->P.d_blk_e0);
  delete E;
  return PCGB_OK;
}

int pcgb_ebe_apply(pcgb_ebe_t E, const double *d_x, double *d_y, void *stream) {
  if (!E || !d_x || !d_y) return fail(PCGB_ERR_ARG, "pcgb_ebe_apply: null argument");
  return ebe_apply(E->P, d_x, d_y, (cudaStream_t)stream);
}

int64_t pcgb_ebe_bytes(pcgb_ebe_t E) { return E ? E->P.bytes : 0; }

// ------------------------------------------------------------------------------------ coloured EBE operator (round-2 prep)
// groups[] holds one entry per (pattern group, colour), sorted by colour; phase[g] is the colour.  Pattern matrices of
// the 24-dof groups are de-duplicated into the constant-memory slots by content of ke_host (same pointer = same slot).
int pcgb_ebe2_create(int64_t n, int ngroups, const pcgb_ebe_group *groups, const int32_t *phase, pcgb_ebe2_t *out) {
  if (!out || n < 0 || ngroups < 0 || (ngroups > 0 && (!groups || !phase))) return fail(PCGB_ERR_ARG, "pcgb_ebe2_create: bad argument");
  if (n >= (1 << 30)) return fail(PCGB_ERR_ARG, "pcgb_ebe2_create: more than 2^30 dofs");
  PCGB_TRY(require_device());
  for (int g = 1; g < ngroups; ++g)
    if (phase[g] < phase[g - 1]) return fail(PCGB_ERR_ARG, "pcgb_ebe2_create: groups must be sorted by phase (colour)");
  pcgb_ebe2_t E = new pcgb_ebe2_s();
  EbePlan &P = E->C.P;
  P.n = n;
  P.bytes = 16 * n;
  std::vector<const double *> slot_key;
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
      int slot = -1;
      for (size_t k = 0; k < slot_key.size(); ++k) if (slot_key[k] == src.ke_host) slot = (int)k;
      if (slot < 0 && (int)slot_key.size() < kEbeMaxSlots) {
        slot = (int)slot_key.size();
        PCGB_CUDA(cudaMemcpyToSymbol(c_ebe_ke24, src.ke_host, 576 * sizeof(double), (size_t)slot * 576 * sizeof(double)));
        slot_key.push_back(src.ke_host);
      }
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
  const size_t nb = (size_t)(s->n > 0 ? s->n : 1) * sizeof(double);
  PCGB_CUDA(cudaMalloc(&s->r, nb));
  PCGB_CUDA(cudaMalloc(&s->p, nb));
  PCGB_CUDA(cudaMalloc(&s->q, nb));
  PCGB_CUDA(cudaMalloc(&s->xalt, nb));
  PCGB_CUDA(cudaMalloc(&s->xown, nb));
  PCGB_CUDA(cudaMemset(s->p, 0, nb));
  PCGB_CUDA(cudaMalloc(&s->partials, 5 * kMaxVecGrid * sizeof(double)));
  s->stage_cap = ((A ? A->P.ntiles : 0) + 4095) / 4096 + 1;
  PCGB_CUDA(cudaMalloc(&s->stage, (size_t)s->stage_cap * sizeof(double)));
  PCGB_CUDA(cudaMalloc(&s->red, 8 * sizeof(double)));
  PCGB_CUDA(cudaMemset(s->red, 0, 8 * sizeof(double)));
  PCGB_CUDA(cudaMalloc(&s->d_ctrl, sizeof(PcgCtrl)));
  PCGB_CUDA(cudaMallocHost(&s->h_ctrl, sizeof(PcgCtrl)));
  PCGB_CUDA(cudaMallocHost(&s->h_red, 8 * sizeof(double)));
  PCGB_CUDA(cudaStreamCreateWithFlags(&s->own, cudaStreamNonBlocking));
  PCGB_CUDA(cudaEventCreateWithFlags(&s->ev_in, cudaEventDisableTiming));
  PCGB_CUDA(cudaEventCreateWithFlags(&s->ev_out, cudaEventDisableTiming));
  PCGB_CUDA(cudaEventCreate(&s->ev_l0));
  PCGB_CUDA(cudaEventCreate(&s->ev_l1));
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
  if (s->ev_in) cudaEventDestroy(s->ev_in);
  if (s->ev_out) cudaEventDestroy(s->ev_out);
  if (s->ev_l0) cudaEventDestroy(s->ev_l0);
  if (s->ev_l1) cudaEventDestroy(s->ev_l1);
  for (cudaEvent_t e : s->ev_k) cudaEventDestroy(e);
  cudaFree(s->r); cudaFree(s->p); cudaFree(s->q); cudaFree(s->xalt); cudaFree(s->xown); cudaFree(s->partials); cudaFree(s->stage);
  cudaFree(s->red); cudaFree(s->d_ctrl);
  cudaFreeHost(s->h_ctrl); cudaFreeHost(s->h_red);
  delete s;
  return PCGB_OK;
}

}  // extern "C"

// ---- host helpers of the solve -------------------------------------------------------------
namespace {

// y = A x + interface sum  (calcMPFint, pcg_solver.py:339-342)
int op_apply(pcgb_solver_t s, const double *x, double *y, cudaStream_t st) {
  if (s->E) PCGB_TRY(ebe_apply(s->E->P, x, y, st, &s->launches));
  else PCGB_TRY(spmv_launch(s->A->P, x, y, false, st, &s->launches));
  if (s->halo) PCGB_TRY(halo_exchange_add(s->halo, y, st, &s->launches));
  return PCGB_OK;
}

// sum `nv` rows of s->partials (count entries each) over blocks and ranks; result in s->h_red[0..nv)
int reduce_to_host(pcgb_solver_t s, int count, int nv, cudaStream_t st) {
  double *dst = s->red + 6;  // nv <= 2
  PCGB_TRY(reduce_rows(s->partials, count, nv, dst, st));
  s->launches += 1;
  if (s->comm && s->comm->nranks > 1) PCGB_TRY(pcgb_allreduce_sum(s->comm, dst, nv, st));
  PCGB_CUDA(cudaMemcpyAsync(s->h_red, dst, nv * sizeof(double), cudaMemcpyDeviceToHost, st));
  PCGB_CUDA(cudaStreamSynchronize(st));
  return PCGB_OK;
}

// r = b - A x ; returns sqrt(sum r*r*w) over all ranks
int true_residual(pcgb_solver_t s, const double *b, const double *w, const double *x, double *normr, cudaStream_t st) {
  PCGB_TRY(op_apply(s, x, s->q, st));
  const int grid = vec_grid(s->n);
  k_residual<<<grid, kVecBlock, 0, st>>>(s->n, b, s->q, w, s->r, s->partials);
  PCGB_CHECK_LAUNCH();
  s->launches += 1;
  PCGB_TRY(reduce_to_host(s, grid, 1, st));
  *normr = std::sqrt(s->h_red[0]);
  return PCGB_OK;
}

// (sum z.(r w), #inf z) of the current r into s->red[6..7] (device, all-reduced)
int rz_to_device(pcgb_solver_t s, const double *minv, const double *w, cudaStream_t st) {
  const int grid = vec_grid(s->n);
  k_rz<<<grid, kVecBlock, 0, st>>>(s->n, s->r, minv, w, s->partials);
  PCGB_CHECK_LAUNCH();
  PCGB_TRY(reduce_rows(s->partials, grid, 2, s->red + 6, st));
  s->launches += 2;
  if (s->comm && s->comm->nranks > 1) PCGB_TRY(pcgb_allreduce_sum(s->comm, s->red + 6, 2, st));
  return PCGB_OK;
}

// one PCG iteration enqueued on st (pcg_solver.py:438-562); every kernel no-ops once the state is frozen
int enqueue_iteration(pcgb_solver_t s, const double *minv, const double *w, double *xb0, double *resvec, cudaStream_t st, int *nl,
                      cudaEvent_t ka = nullptr, cudaEvent_t kb = nullptr) {
  const int64_t n = s->n;
  const int vg = vec_grid(n);
  const bool multi = s->comm && s->comm->nranks > 1;
  k_pupdate<<<vg, kVecBlock, 0, st>>>(s->d_ctrl, n, s->r, minv, s->p);
  PCGB_CHECK_LAUNCH();
  *nl += 1;
  // p.q : per-tile partials -> (stage) -> scalar.  In the multi-GPU case the partials are those of the
  // UNASSEMBLED local product, whose rank sum equals the reference's weighted dot of the assembled q
  // (p is consistent on shared dofs and K = sum of the subdomain matrices); see DESIGN.md.
  const double *pq_src;
  int pq_cnt;
  if (s->E) {  // experimental matrix-free operator: product, then a separate unweighted dot of the local product
    if (ka) PCGB_CUDA(cudaEventRecord(ka, st));
    PCGB_TRY(ebe_apply(s->E->P, s->p, s->q, st, nl, &s->d_ctrl->state));
    if (kb) PCGB_CUDA(cudaEventRecord(kb, st));
    k_dot_w<<<vg, kVecBlock, 0, st>>>(n, s->p, s->q, nullptr, s->partials);
    PCGB_CHECK_LAUNCH();
    *nl += 1;
    pq_src = s->partials; pq_cnt = vg;
  } else {
    const CsrPlan &P = s->A->P;
    if (ka) PCGB_CUDA(cudaEventRecord(ka, st));
    PCGB_TRY(spmv_launch(P, s->p, s->q, true, st, nl, &s->d_ctrl->state));
    if (kb) PCGB_CUDA(cudaEventRecord(kb, st));
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
  if (s->halo) PCGB_TRY(halo_exchange_add(s->halo, s->q, st, nl));
  if (!multi) {
    k_reduce<1, 1><<<1, 256, 0, st>>>(s->d_ctrl, pq_src, pq_cnt, 0, s->red, nullptr);
    PCGB_CHECK_LAUNCH();
    *nl += 1;
  } else {
    k_reduce<1, 0><<<1, 256, 0, st>>>(s->d_ctrl, pq_src, pq_cnt, 0, s->red, nullptr);
    PCGB_CHECK_LAUNCH();
    PCGB_TRY(pcgb_allreduce_sum(s->comm, s->red, 1, st));
    k_ctrl_alpha<<<1, 1, 0, st>>>(s->d_ctrl, s->red);
    PCGB_CHECK_LAUNCH();
    *nl += 2;
  }
  k_update<<<vg, kVecBlock, 0, st>>>(s->d_ctrl, n, s->r, s->q, s->p, minv, w, xb0, s->xalt, s->partials);
  PCGB_CHECK_LAUNCH();
  *nl += 1;
  if (!multi) {
    k_reduce<5, 2><<<1, 256, 0, st>>>(s->d_ctrl, s->partials, vg, kMaxVecGrid, s->red + 1, resvec);
    PCGB_CHECK_LAUNCH();
    *nl += 1;
  } else {
    k_reduce<5, 0><<<1, 256, 0, st>>>(s->d_ctrl, s->partials, vg, kMaxVecGrid, s->red + 1, nullptr);
    PCGB_CHECK_LAUNCH();
    PCGB_TRY(pcgb_allreduce_sum(s->comm, s->red + 1, 5, st));
    k_ctrl_norms<<<1, 1, 0, st>>>(s->d_ctrl, s->red + 1, resvec);
    PCGB_CHECK_LAUNCH();
    *nl += 2;
  }
  return PCGB_OK;
}

int fetch_ctrl(pcgb_solver_t s, cudaStream_t st) {
  PCGB_CUDA(cudaMemcpyAsync(s->h_ctrl, s->d_ctrl, sizeof(PcgCtrl), cudaMemcpyDeviceToHost, st));
  PCGB_CUDA(cudaStreamSynchronize(st));
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

  // ---- ||b||  (pcg_solver.py:381-384)
  k_dot_w<<<vg, kVecBlock, 0, st>>>(n, d_b, d_b, d_w, s->partials);
  PCGB_CHECK_LAUNCH();
  s->launches += 1;
  PCGB_TRY(reduce_to_host(s, vg, 1, st));
  const double n2b = std::sqrt(s->h_red[0]);
  const double tolb = opt->tol * n2b;
  res->normb = n2b;
  if (n2b == 0.0) {  // :387-395 - returns the initial guess, flag 0, relres 0, iter 0
    res->flag = 0; res->relres = 0.0; res->iters = 0; res->launches = s->launches;
    return PCGB_OK;
  }
  // ---- initial residual (:408-418)
  double normr = 0.0;
  PCGB_CUDA(cudaMemcpyAsync(xw, d_x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  PCGB_TRY(true_residual(s, d_b, d_w, xw, &normr, st));
  ++matvecs;
  if (d_resvec) {
    s->h_red[2] = normr;  // ResVec[0] (:431)
    PCGB_CUDA(cudaMemcpyAsync(d_resvec, &s->h_red[2], sizeof(double), cudaMemcpyHostToDevice, st));
    PCGB_CUDA(cudaStreamSynchronize(st));
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
  c.maxiter = opt->maxiter; c.maxstag = maxstag; c.fixed_iters = opt->fixed_iters ? 1 : 0;
  *s->h_ctrl = c;
  PCGB_CUDA(cudaMemcpyAsync(s->d_ctrl, s->h_ctrl, sizeof(PcgCtrl), cudaMemcpyHostToDevice, st));
  PCGB_CUDA(cudaStreamSynchronize(st));
  // head of iteration 0: rho = z.r (:446-469)
  PCGB_TRY(rz_to_device(s, d_minv, d_w, st));
  k_ctrl_head<<<1, 1, 0, st>>>(s->d_ctrl, s->red + 6, 0);
  PCGB_CHECK_LAUNCH();
  s->launches += 1;

  int batch = opt->check_every > 0 ? opt->check_every : 16;
  if (batch > opt->maxiter) batch = opt->maxiter;
  const bool want_graph = opt->use_graph != 0 && !opt->time_kernels;
  size_t kpairs = 0;
  PCGB_CUDA(cudaEventRecord(s->ev_l0, st));
  int flag = 1;
  int too_small = 0;

  for (;;) {
    // ---- enqueue `batch` iterations
    if (want_graph && batch > 1) {
      GraphKey key{d_minv, d_w, xw, d_resvec, batch};
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
        cudaEvent_t ka = nullptr, kb = nullptr;
        if (opt->time_kernels && kpairs < 1024) {
          while (s->ev_k.size() < 2 * (kpairs + 1)) { cudaEvent_t e; PCGB_CUDA(cudaEventCreate(&e)); s->ev_k.push_back(e); }
          ka = s->ev_k[2 * kpairs]; kb = s->ev_k[2 * kpairs + 1]; ++kpairs;
        }
        PCGB_TRY(enqueue_iteration(s, d_minv, d_w, xw, d_resvec, st, &nl, ka, kb));
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
      if (normr_act < c.normrmin) { c.normrmin = normr_act; c.xmin = c.xcur; c.imin = c.iter; }  // :555-558
      if (c.stag >= maxstag) { flag = 3; c.flag = 3; break; }  // :560-562
      // continue with the replaced residual: head of the next iteration
      c.state = ST_RUN;
      *s->h_ctrl = c;
      PCGB_CUDA(cudaMemcpyAsync(s->d_ctrl, s->h_ctrl, sizeof(PcgCtrl), cudaMemcpyHostToDevice, st));
      PCGB_CUDA(cudaStreamSynchronize(st));
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
  PCGB_CUDA(cudaEventSynchronize(s->ev_l1));
  {
    float ms = 0.f;
    PCGB_CUDA(cudaEventElapsedTime(&ms, s->ev_l0, s->ev_l1));
    res->loop_ms = ms;
    res->loop_iters = c.iter + 1;
    double tot = 0.0;
    for (size_t k = 0; k < kpairs; ++k) { float t = 0.f; PCGB_CUDA(cudaEventElapsedTime(&t, s->ev_k[2 * k], s->ev_k[2 * k + 1])); tot += t; }
    res->spmv_ms = tot; res->spmv_timed = (int64_t)kpairs;
  }

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
    double normr_min = 0.0;
    PCGB_TRY(true_residual(s, d_b, d_w, xbuf[c.xmin], &normr_min, st));
    ++matvecs;
    if (normr_min < c.normr_act) { iter_out = c.imin; relres = normr_min / n2b; }
    else { iter_out = i; relres = c.normr_act / n2b; }
    xout = c.xmin;  // the reference exports XMin on this path in both cases (:569, :598)
  }
  iter_out += 1;  // :584
  PCGB_CUDA(cudaMemcpyAsync(d_x, xbuf[xout], (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  PCGB_CUDA(cudaStreamSynchronize(st));
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
