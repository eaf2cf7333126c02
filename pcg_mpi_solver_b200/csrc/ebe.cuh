// ebe.cuh - opt-in matrix-free operator (parity-green on B200; ncu: profiles/ncu_ebe_t24_r2.txt; 0.27 ms per application at 128^3):
// the reference's OWN operator on the GPU - the pattern-grouped, matrix-free element-by-element product
//     y = sum_groups scatter( S . Ke . (Ck o (S . gather(x))) )            calcMatVecProd, pcg_solver.py:263-300
// instead of the assembled CSR form.  SURVEY.md 8(f1): ~108 B per element (24 int32 dof ids + Ck + signs)
// against ~2.4 KB per element of CSR for the hex mesh, i.e. ~20x fewer HBM bytes per matvec.
//
// Layout per pattern group (same as the reference's type groups, partition_mesh.py:470-491, but int32 and on
// the free-dof numbering): idx[nd][ne] (-1 = clamped dof), sign[nd][ne] (uint8, may be NULL), ck[ne], ke[nd][nd].
// Kernels:
//   k_ebe_t24    one THREAD per element for the 24-dof patterns (cube; 85 % of concrete, 100 % of the hex mesh):
//                u[24] in registers, Ke from constant memory (operand of the DFMA), 576 DFMA per element;
//   k_ebe_warp   one WARP per element for every other pattern size (24 < nd <= 96), Ke read column-wise
//                (symmetric) through L1.
// Scatter-add uses fp64 atomicAdd (RED.E.ADD.F64): summation order across elements is not fixed, results are
// reproducible to rounding only; ebe_color.cuh is the colouring-based deterministic variant (bit-reproducible, 3x slower).
#pragma once
#include <vector>

#include "common.cuh"

namespace pcgb {

constexpr int kEbeMaxSlots = 8;  // 24-dof pattern matrices resident in constant memory
__constant__ double c_ebe_ke24[kEbeMaxSlots][24 * 24];

struct EbeGroup {
  int nd = 0;
  int64_t ne = 0;
  const int *idx = nullptr;             // [nd][ne]
  const unsigned char *sign = nullptr;  // [nd][ne] or null
  const double *ck = nullptr;           // [ne]
  const double *ke = nullptr;           // [nd][nd] device copy
  int slot = -1;                        // constant-memory slot for nd == 24
};

struct EbePlan {
  int64_t n = 0;
  std::vector<EbeGroup> groups;
  EbeGroup *d_groups = nullptr;  // device copy for k_ebe_warp
  int *d_blk_group = nullptr;    // k_ebe_warp: block -> group
  int64_t *d_blk_e0 = nullptr;   //             block -> first element
  int nblk_warp = 0;
  int64_t bytes = 0;             // algorithmic bytes of one application
};

constexpr int kEbeWarpsPerBlock = 8;

template <bool SIGN>
__global__ void __launch_bounds__(128)
k_ebe_t24(const int *__restrict__ idx, const unsigned char *__restrict__ sign, const double *__restrict__ ck, int slot, int64_t ne,
          const double *__restrict__ x, double *__restrict__ y, const int *__restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  const int64_t e = blockIdx.x * (int64_t)128 + threadIdx.x;
  if (e >= ne) return;
  int id[24];
  double u[24];
  unsigned sbits = 0;
#pragma unroll
  for (int j = 0; j < 24; ++j) {
    id[j] = idx[(int64_t)j * ne + e];
    double v = id[j] >= 0 ? __ldg(x + id[j]) : 0.0;
    if (SIGN && sign[(int64_t)j * ne + e]) { v = -v; sbits |= 1u << j; }   // pcg_solver.py:278
    u[j] = v;
  }
  const double c = ck[e];                                                   // :279  Ke @ (Ck * U)
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 24; ++j) acc = fma(c_ebe_ke24[slot][i * 24 + j], u[j], acc);
    acc *= c;
    if (SIGN && ((sbits >> i) & 1u)) acc = -acc;                            // :280
    if (id[i] >= 0) atomicAdd(y + id[i], acc);                              // np.bincount scatter-add, :300
  }
}

__global__ void __launch_bounds__(kEbeWarpsPerBlock * 32)
k_ebe_warp(const EbeGroup *__restrict__ groups, const int *__restrict__ blk_group, const int64_t *__restrict__ blk_e0,
           const double *__restrict__ x, double *__restrict__ y, const int *__restrict__ skip) {
  if (skip != nullptr && *skip != 0) return;
  __shared__ double su[kEbeWarpsPerBlock][96];
  __shared__ int sid[kEbeWarpsPerBlock][96];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const EbeGroup g = groups[blk_group[blockIdx.x]];
  const int64_t e = blk_e0[blockIdx.x] + warp;
  if (e >= g.ne) return;
  const int nd = g.nd;
  for (int j = lane; j < nd; j += 32) {
    const int id = g.idx[(int64_t)j * g.ne + e];
    double v = id >= 0 ? __ldg(x + id) : 0.0;
    const bool s = g.sign != nullptr && g.sign[(int64_t)j * g.ne + e] != 0;
    su[warp][j] = s ? -v : v;
    sid[warp][j] = s ? (id | (int)0x40000000) : id;  // bit 30 carries the sign flag (ids < 2^30)
  }
  __syncwarp();
  const double c = g.ck[e];
  for (int i = lane; i < nd; i += 32) {
    double acc = 0.0;
    for (int j = 0; j < nd; ++j) acc = fma(__ldg(g.ke + (int64_t)j * nd + i), su[warp][j], acc);  // Ke symmetric: column i = row i
    acc *= c;
    int id = sid[warp][i];
    if (id >= 0) {
      if (id & 0x40000000) { acc = -acc; id &= 0x3fffffff; }
      atomicAdd(y + id, acc);
    }
  }
}

inline int ebe_apply(const EbePlan &P, const double *x, double *y, cudaStream_t st, int *launches = nullptr, const int *skip = nullptr) {
  PCGB_CUDA(cudaMemsetAsync(y, 0, (size_t)P.n * sizeof(double), st));
  int nl = 0;
  for (const EbeGroup &g : P.groups) {
    if (g.nd != 24 || g.slot < 0 || g.ne == 0) continue;
    const unsigned grid = (unsigned)((g.ne + 127) / 128);
    if (g.sign) k_ebe_t24<true><<<grid, 128, 0, st>>>(g.idx, g.sign, g.ck, g.slot, g.ne, x, y, skip);
    else k_ebe_t24<false><<<grid, 128, 0, st>>>(g.idx, nullptr, g.ck, g.slot, g.ne, x, y, skip);
    PCGB_CHECK_LAUNCH();
    ++nl;
  }
  if (P.nblk_warp > 0) {
    k_ebe_warp<<<P.nblk_warp, kEbeWarpsPerBlock * 32, 0, st>>>(P.d_groups, P.d_blk_group, P.d_blk_e0, x, y, skip);
    PCGB_CHECK_LAUNCH();
    ++nl;
  }
  if (launches) *launches += nl;
  return PCGB_OK;
}

}  // namespace pcgb
