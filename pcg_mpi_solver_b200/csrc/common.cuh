// common.cuh - error handling, small device helpers shared by all kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>

#include "pcgb200.h"

namespace pcgb {

inline std::string &last_error() {
  static thread_local std::string s;
  return s;
}

inline int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

#define PCGB_CUDA(call)                                                                          \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return ::pcgb::fail(PCGB_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

#define PCGB_CHECK_LAUNCH() PCGB_CUDA(cudaGetLastError())

#define PCGB_TRY(call)      \
  do {                      \
    int r_ = (call);        \
    if (r_ != PCGB_OK) return r_; \
  } while (0)

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

inline int num_sms() {
  static int n = -1;
  if (n < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = kSMs;
  }
  return n;
}

// ---- warp / block reductions (cub-style: shuffle tree inside the warp, one smem hop across warps)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

template <int LANES>
__device__ __forceinline__ double group_sum(double v) {  // reduce over aligned groups of LANES lanes
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o, LANES);
  return v;
}

// Block-wide sum of NV values per thread; result valid in thread 0.  `red` needs NV*32 doubles.
template <int NV, int BLOCK>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int NW = BLOCK / 32;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = warp_sum(v[k]);
    if (lane == 0) red[k * 32 + wid] = s;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = lane < NW ? red[k * 32 + lane] : 0.0;
      v[k] = warp_sum(s);
    }
  }
}

// streaming (evict-first) loads for data that is read exactly once per SpMV
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int *p) { return __ldcs(p); }

}  // namespace pcgb
