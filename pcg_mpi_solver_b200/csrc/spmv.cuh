// spmv.cuh - merge-path CSR SpMV for sm_100a (fp64 values, int32 columns, int32/int64 row offsets).
//
// Replaces the operator of the reference, calcMatVecProd(...,'Strain') pcg_solver.py:242-300, on the
// assembled matrix A = K[Eff,Eff].
//
// Design (see DESIGN.md "SpMV"):
//   * plan time (once per matrix): a 2-D merge-path search over (row ends, nnz indices) cuts the
//     work into tiles of `tile_items` merge items; when no row is long, tile starts are snapped back
//     to a row boundary so that no row is split (no fix-up pass, bit-reproducible);
//   * run time, one CTA per tile: the tile's slice of `val` and `col` is brought into shared memory
//     by two TMA bulk copies (cp.async.bulk + mbarrier, L2 evict-first) - the streaming arrays never
//     pass through L1, which is left to the gather of x (ld.global.nc); the row offsets of the tile
//     are loaded meanwhile; then sub-warp groups of LANES lanes walk the rows of the tile, multiply
//     from shared memory against gathered x and finish each row with a shuffle reduction; row sums
//     are staged in shared memory and written to y coalesced, with an optional fused dot-product
//     epilogue  sum_r x[r]*y[r]  (the p.q of pcg_solver.py:487).
#pragma once
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace pcgb {

constexpr int kSpmvBlock = 256;

struct CsrPlan {
  int64_t nrows = 0, ncols = 0, nnz = 0;
  const void *rowptr = nullptr;
  bool rp64 = false;
  const int *col = nullptr;
  const double *val = nullptr;
  // plan
  int tile_items = 0, lanes = 0, ntiles = 0;
  bool snap = false, use_tma = true;
  int cap_nnz = 0, cap_rows = 0, smem_bytes = 0, max_row = 0;
  int *tile_row = nullptr;       // [ntiles+1]
  int64_t *tile_k = nullptr;     // [ntiles+1]
  double *carry = nullptr;       // [ntiles]   head-partial row sums (split mode)
  int nfix = 0;                  // rows that span more than one tile
  int *fix_row = nullptr, *fix_first = nullptr, *fix_cnt = nullptr;
  double *dot_partials = nullptr;  // [ntiles]  fused  x.y  partials
};

// ------------------------------------------------------------------ PTX wrappers (TMA bulk copy)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// global -> shared bulk copy (UBLKCP), completion signalled on the mbarrier as transaction bytes
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}

// ------------------------------------------------------------------ plan kernels
template <typename RP>
__global__ void k_max_row(const RP *__restrict__ rowptr, int64_t nrows, int *out_max) {
  int m = 0;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
    int64_t len = (int64_t)rowptr[r + 1] - (int64_t)rowptr[r];
    m = max(m, (int)min(len, (int64_t)INT32_MAX));
  }
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out_max, m);
}

// Merge-path split of diagonal d over (row_end = rowptr+1, nnz counting sequence): the largest i with
// row_end[r] <= d - r - 1 for all r < i.
template <typename RP>
__device__ __forceinline__ int64_t merge_path_rows(const RP *__restrict__ rowptr, int64_t nrows, int64_t nnz, int64_t d) {
  int64_t lo = d > nnz ? d - nnz : 0, hi = d < nrows ? d : nrows;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)rowptr[mid + 1] <= d - mid - 1) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

template <typename RP>
__global__ void k_partition(const RP *__restrict__ rowptr, int64_t nrows, int64_t nnz, int tile_items, int ntiles, int snap,
                            int *tile_row, int64_t *tile_k) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > ntiles) return;
  int64_t total = nrows + nnz;
  int64_t d = (int64_t)b * tile_items;
  if (d > total || b == ntiles) d = total;
  int64_t i = merge_path_rows(rowptr, nrows, nnz, d);
  int64_t j = d - i;
  if (snap) j = (int64_t)rowptr[i];
  tile_row[b] = (int)i;
  tile_k[b] = j;
}

template <typename RP>
__global__ void k_tile_stats(const RP *__restrict__ rowptr, const int *__restrict__ tile_row, const int64_t *__restrict__ tile_k,
                             int ntiles, int *max_cnt, int *max_rows, unsigned char *head_flag) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= ntiles) return;
  int64_t k0 = tile_k[b], k1 = tile_k[b + 1];
  int r0 = tile_row[b], r1 = tile_row[b + 1];
  int cnt = (int)(k1 - (k0 & ~(int64_t)3));
  atomicMax(max_cnt, cnt);
  atomicMax(max_rows, r1 - r0 + 1);
  head_flag[b] = (int64_t)rowptr[r0] < k0 ? 1 : 0;
}

// ------------------------------------------------------------------ the SpMV kernel
template <int LANES, bool TMA, bool DOT, typename RP>
__global__ void __launch_bounds__(kSpmvBlock)
k_spmv_merge(const RP *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
             const double *__restrict__ x, double *__restrict__ y, const int *__restrict__ tile_row,
             const int64_t *__restrict__ tile_k, int64_t nrows, int64_t nnz, int cap_nnz, int cap_rows,
             double *__restrict__ carry, double *__restrict__ dot_partials, const int *__restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;  // PCG state frozen (pcg_kernels.cuh): nothing to do
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *sval = reinterpret_cast<double *>(smem_raw);
  double *srow = sval + cap_nnz;                          // cap_rows doubles
  int *scol = reinterpret_cast<int *>(srow + cap_rows);   // cap_nnz ints  (cap_nnz % 4 == 0, cap_rows % 2 == 0)
  int *soff = scol + cap_nnz;                             // cap_rows + 1 ints
  __shared__ uint64_t bar;
  __shared__ double red[32];

  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const int r0 = tile_row[b], r1 = tile_row[b + 1];
  const int64_t k0 = tile_k[b], k1 = tile_k[b + 1];
  const int64_t ka = k0 & ~(int64_t)3;  // 16-byte aligned start of the staged window
  const int k0l = (int)(k0 - ka), k1l = (int)(k1 - ka);
  const bool tail = (r1 < nrows) && (k1 > (int64_t)rowptr[r1]);
  const int R = r1 - r0 + (tail ? 1 : 0);  // rows that own at least a (possibly empty) segment here
  const bool head = (R > 0) && ((int64_t)rowptr[r0] < k0);

  // ---- stage val/col slices into shared memory
  if (TMA) {
    if (tid == 0) {
      mbar_init(&bar, 1);
      mbar_fence_init();
      int64_t kend = (k1 + 3) & ~(int64_t)3;
      const int64_t lim = nnz & ~(int64_t)3;
      if (kend > lim) kend = lim;
      const int nb = (int)(kend - ka);
      if (nb > 0) {
        const uint64_t pol = l2_evict_first_policy();
        mbar_expect_tx(&bar, (uint32_t)nb * 12u);
        bulk_g2s(sval, val + ka, (uint32_t)nb * 8u, &bar, pol);
        bulk_g2s(scol, col + ka, (uint32_t)nb * 4u, &bar, pol);
      } else {
        mbar_arrive(&bar);
      }
    }
    // the (at most 3) elements beyond the last full 16-byte quad of the arrays
    {
      const int64_t lim = nnz & ~(int64_t)3;
      int64_t kt = (lim > ka ? lim : ka) + tid;
      if (tid < 4 && kt >= lim && kt < k1) {
        sval[kt - ka] = val[kt];
        scol[kt - ka] = col[kt];
      }
    }
  } else {
#pragma unroll 4
    for (int j = k0l + tid; j < k1l; j += kSpmvBlock) {
      sval[j] = ld_stream(val + ka + j);
      scol[j] = ld_stream(col + ka + j);
    }
  }
  // ---- row offsets of the tile, relative to ka and clamped to the tile's nnz window
  for (int i = tid; i <= R; i += kSpmvBlock) {
    int64_t v = (int64_t)rowptr[r0 + i] - ka;
    v = v < k0l ? k0l : (v > k1l ? k1l : v);
    soff[i] = (int)v;
  }
  __syncthreads();
  if (TMA) mbar_wait(&bar, 0);

  // ---- segmented reduction: one group of LANES lanes per row, shuffle tree at the end
  constexpr int G = kSpmvBlock / LANES;
  const int gid = tid / LANES, gl = tid % LANES;
  // NOTE: the trip count is uniform over the CTA (i0, not i, is tested) because the shuffle
  // reduction below names all 32 lanes; lanes of a group past the last row just idle.
  for (int i0 = 0; i0 < R; i0 += G) {
    const int i = i0 + gid;
    const bool live = i < R;
    const int a = live ? soff[i] : 0, e = live ? soff[i + 1] : 0;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int j = a + gl;
    for (; j + 3 * LANES < e; j += 4 * LANES) {
      const int c0 = scol[j], c1 = scol[j + LANES], c2 = scol[j + 2 * LANES], c3 = scol[j + 3 * LANES];
      const double v0 = sval[j], v1 = sval[j + LANES], v2 = sval[j + 2 * LANES], v3 = sval[j + 3 * LANES];
      const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
      acc0 = fma(v0, x0, acc0);
      acc1 = fma(v1, x1, acc1);
      acc2 = fma(v2, x2, acc2);
      acc3 = fma(v3, x3, acc3);
    }
    for (; j < e; j += LANES) acc0 = fma(sval[j], __ldg(x + scol[j]), acc0);
    double acc = group_sum<LANES>((acc0 + acc1) + (acc2 + acc3));
    if (live && gl == 0) srow[i] = acc;
  }
  __syncthreads();

  // ---- coalesced write-back (+ fused dot epilogue)
  double d = 0.0;
  for (int i = tid; i < R; i += kSpmvBlock) {
    const double s = srow[i];
    const int r = r0 + i;
    if (i == 0 && head) carry[b] = s;
    else y[r] = s;
    if (DOT) d = fma(s, __ldg(x + r), d);
  }
  if (DOT) {
    double v[1] = {d};
    block_sum<1, kSpmvBlock>(v, red);
    if (tid == 0) dot_partials[b] = v[0];
  }
}

// rows that span several tiles: y[row] (written by the tile where the row starts) += carries, in tile order
__global__ void k_spmv_fixup(double *__restrict__ y, const double *__restrict__ carry, const int *__restrict__ fix_row,
                             const int *__restrict__ fix_first, const int *__restrict__ fix_cnt, int nfix) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nfix) return;
  const int row = fix_row[f], first = fix_first[f], cnt = fix_cnt[f];
  double s = y[row];
  for (int t = 0; t < cnt; ++t) s += carry[first + t];
  y[row] = s;
}

template <typename RP>
__global__ void k_csr_diag(const RP *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
                           int64_t nrows, double *__restrict__ diag) {
  // one 8-lane group per row
  const int64_t gidx = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 3;
  const int gl = threadIdx.x & 7;
  if (gidx >= nrows) return;
  double d = 0.0;
  for (int64_t k = (int64_t)rowptr[gidx] + gl; k < (int64_t)rowptr[gidx + 1]; k += 8)
    if (col[k] == gidx) d += val[k];
  d = group_sum<8>(d);
  if (gl == 0) diag[gidx] = d;
}

// ------------------------------------------------------------------ host side
inline int spmv_configure(const CsrPlan &P);
inline int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

template <typename RP>
inline int build_plan_t(CsrPlan &P, cudaStream_t st) {
  const RP *rp = static_cast<const RP *>(P.rowptr);
  int *d_stats = nullptr;
  PCGB_CUDA(cudaMalloc(&d_stats, 4 * sizeof(int)));
  PCGB_CUDA(cudaMemsetAsync(d_stats, 0, 4 * sizeof(int), st));
  if (P.nrows > 0) {
    int grid = (int)std::min<int64_t>((P.nrows + 255) / 256, 148 * 8);
    k_max_row<RP><<<grid, 256, 0, st>>>(rp, P.nrows, d_stats);
    PCGB_CHECK_LAUNCH();
  }
  int h_stats[4] = {0, 0, 0, 0};
  PCGB_CUDA(cudaMemcpyAsync(h_stats, d_stats, sizeof(int), cudaMemcpyDeviceToHost, st));
  PCGB_CUDA(cudaStreamSynchronize(st));
  P.max_row = h_stats[0];

  P.tile_items = env_int("PCGB_SPMV_TILE", 2048);  // 8 CTAs/SM: sweep in profiles/spmv_sweep_r1.txt
  if (P.tile_items < 256) P.tile_items = 256;
  const double avg = P.nrows ? (double)P.nnz / (double)P.nrows : 0.0;
  int lanes = avg <= 12.0 ? 4 : avg <= 100.0 ? 8 : avg <= 200.0 ? 16 : 32;
  lanes = env_int("PCGB_SPMV_LANES", lanes);
  if (lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32) lanes = 16;
  P.lanes = lanes;
  P.snap = P.max_row <= P.tile_items / 4 && env_int("PCGB_SPMV_SNAP", 1) != 0;
  P.use_tma = env_int("PCGB_SPMV_TMA", 1) != 0;
  // TMA bulk copies need 16-byte aligned global sources
  if ((reinterpret_cast<uintptr_t>(P.val) & 15) || (reinterpret_cast<uintptr_t>(P.col) & 15)) P.use_tma = false;

  const int64_t total = P.nrows + P.nnz;
  int64_t nt = (total + P.tile_items - 1) / P.tile_items;
  if (nt < 1) nt = 1;
  if (nt > INT32_MAX - 2) return fail(PCGB_ERR_ARG, "matrix too large for the tile index (ntiles=%lld)", (long long)nt);
  P.ntiles = (int)nt;
  PCGB_CUDA(cudaMalloc(&P.tile_row, (size_t)(P.ntiles + 1) * sizeof(int)));
  PCGB_CUDA(cudaMalloc(&P.tile_k, (size_t)(P.ntiles + 1) * sizeof(int64_t)));
  PCGB_CUDA(cudaMalloc(&P.carry, (size_t)P.ntiles * sizeof(double)));
  PCGB_CUDA(cudaMalloc(&P.dot_partials, (size_t)P.ntiles * sizeof(double)));
  PCGB_CUDA(cudaMemsetAsync(P.carry, 0, (size_t)P.ntiles * sizeof(double), st));
  unsigned char *d_head = nullptr;
  PCGB_CUDA(cudaMalloc(&d_head, (size_t)P.ntiles));
  k_partition<RP><<<(P.ntiles + 1 + 255) / 256, 256, 0, st>>>(rp, P.nrows, P.nnz, P.tile_items, P.ntiles, P.snap ? 1 : 0,
                                                               P.tile_row, P.tile_k);
  PCGB_CHECK_LAUNCH();
  k_tile_stats<RP><<<(P.ntiles + 255) / 256, 256, 0, st>>>(rp, P.tile_row, P.tile_k, P.ntiles, d_stats + 1, d_stats + 2, d_head);
  PCGB_CHECK_LAUNCH();
  PCGB_CUDA(cudaMemcpyAsync(h_stats, d_stats, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  PCGB_CUDA(cudaStreamSynchronize(st));
  P.cap_nnz = (h_stats[1] + 8 + 3) & ~3;
  P.cap_rows = (h_stats[2] + 2 + 1) & ~1;
  P.smem_bytes = P.cap_nnz * 12 + P.cap_rows * 8 + (P.cap_rows + 2) * 4;
  P.smem_bytes = (P.smem_bytes + 127) & ~127;
  if (P.smem_bytes > 200 * 1024)
    return fail(PCGB_ERR_ARG, "SpMV tile needs %d bytes of shared memory (tile_items=%d, max_row=%d)", P.smem_bytes,
                P.tile_items, P.max_row);

  // fix-up list for rows that span tiles (split mode only)
  P.nfix = 0;
  if (!P.snap) {
    std::vector<unsigned char> head(P.ntiles);
    std::vector<int> trow(P.ntiles + 1);
    PCGB_CUDA(cudaMemcpy(head.data(), d_head, (size_t)P.ntiles, cudaMemcpyDeviceToHost));
    PCGB_CUDA(cudaMemcpy(trow.data(), P.tile_row, (size_t)(P.ntiles + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int> frow, ffirst, fcnt;
    for (int b = 0; b < P.ntiles;) {
      if (!head[b]) { ++b; continue; }
      int e = b;
      while (e + 1 < P.ntiles && head[e + 1] && trow[e + 1] == trow[b]) ++e;
      frow.push_back(trow[b]); ffirst.push_back(b); fcnt.push_back(e - b + 1);
      b = e + 1;
    }
    P.nfix = (int)frow.size();
    if (P.nfix) {
      PCGB_CUDA(cudaMalloc(&P.fix_row, P.nfix * sizeof(int)));
      PCGB_CUDA(cudaMalloc(&P.fix_first, P.nfix * sizeof(int)));
      PCGB_CUDA(cudaMalloc(&P.fix_cnt, P.nfix * sizeof(int)));
      PCGB_CUDA(cudaMemcpy(P.fix_row, frow.data(), P.nfix * sizeof(int), cudaMemcpyHostToDevice));
      PCGB_CUDA(cudaMemcpy(P.fix_first, ffirst.data(), P.nfix * sizeof(int), cudaMemcpyHostToDevice));
      PCGB_CUDA(cudaMemcpy(P.fix_cnt, fcnt.data(), P.nfix * sizeof(int), cudaMemcpyHostToDevice));
    }
  }
  cudaFree(d_head);
  cudaFree(d_stats);
  return spmv_configure(P);
}

template <int LANES, bool TMA, bool DOT, typename RP>
inline int launch_spmv_inst(const CsrPlan &P, const double *x, double *y, cudaStream_t st, const int *skip) {
  auto kern = k_spmv_merge<LANES, TMA, DOT, RP>;
  if (P.ntiles == 0) return PCGB_OK;
  if (skip == reinterpret_cast<const int *>(1)) {  // configuration request from build_plan
    PCGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return PCGB_OK;
  }
  kern<<<P.ntiles, kSpmvBlock, P.smem_bytes, st>>>(static_cast<const RP *>(P.rowptr), P.col, P.val, x, y, P.tile_row, P.tile_k,
                                                   P.nrows, P.nnz, P.cap_nnz, P.cap_rows, P.carry, P.dot_partials, skip);
  PCGB_CHECK_LAUNCH();
  return PCGB_OK;
}

template <bool TMA, bool DOT, typename RP>
inline int launch_spmv_lanes(const CsrPlan &P, const double *x, double *y, cudaStream_t st, const int *skip) {
  switch (P.lanes) {
    case 4: return launch_spmv_inst<4, TMA, DOT, RP>(P, x, y, st, skip);
    case 8: return launch_spmv_inst<8, TMA, DOT, RP>(P, x, y, st, skip);
    case 32: return launch_spmv_inst<32, TMA, DOT, RP>(P, x, y, st, skip);
    default: return launch_spmv_inst<16, TMA, DOT, RP>(P, x, y, st, skip);
  }
}

// y = A x ; with_dot additionally leaves per-tile partials of x.y in P.dot_partials.
// Returns the number of kernel launches through *launches.
inline int spmv_launch(const CsrPlan &P, const double *x, double *y, bool with_dot, cudaStream_t st, int *launches = nullptr,
                       const int *skip = nullptr) {
  int rc;
  if (P.rp64) {
    if (P.use_tma) rc = with_dot ? launch_spmv_lanes<true, true, int64_t>(P, x, y, st, skip) : launch_spmv_lanes<true, false, int64_t>(P, x, y, st, skip);
    else rc = with_dot ? launch_spmv_lanes<false, true, int64_t>(P, x, y, st, skip) : launch_spmv_lanes<false, false, int64_t>(P, x, y, st, skip);
  } else {
    if (P.use_tma) rc = with_dot ? launch_spmv_lanes<true, true, int32_t>(P, x, y, st, skip) : launch_spmv_lanes<true, false, int32_t>(P, x, y, st, skip);
    else rc = with_dot ? launch_spmv_lanes<false, true, int32_t>(P, x, y, st, skip) : launch_spmv_lanes<false, false, int32_t>(P, x, y, st, skip);
  }
  PCGB_TRY(rc);
  if (skip == reinterpret_cast<const int *>(1)) return PCGB_OK;
  int n = 1;
  if (P.nfix > 0) {
    k_spmv_fixup<<<(P.nfix + 127) / 128, 128, 0, st>>>(y, P.carry, P.fix_row, P.fix_first, P.fix_cnt, P.nfix);
    PCGB_CHECK_LAUNCH();
    ++n;
  }
  if (launches) *launches += n;
  return PCGB_OK;
}

// raise the dynamic shared memory limit of the instantiations this plan will launch (done once at
// plan time so that nothing but launches happens inside CUDA-graph capture)
inline int spmv_configure(const CsrPlan &P) {
  const int *cfg = reinterpret_cast<const int *>(1);
  PCGB_TRY(spmv_launch(P, nullptr, nullptr, false, 0, nullptr, cfg));
  PCGB_TRY(spmv_launch(P, nullptr, nullptr, true, 0, nullptr, cfg));
  return PCGB_OK;
}

inline void free_plan(CsrPlan &P) {
  cudaFree(P.tile_row); cudaFree(P.tile_k); cudaFree(P.carry); cudaFree(P.dot_partials);
  cudaFree(P.fix_row); cudaFree(P.fix_first); cudaFree(P.fix_cnt);
}

}  // namespace pcgb
