// peer.cuh - collectives of the PCG hot path written directly over NVLink / NVSwitch peer memory (CUDA IPC).
//
// The reference exchanges through mpi4py: MPI_SUM (pcg_solver.py:622-628) three times per iteration and the
// neighbour Isend/Recv of the interface values (pcg_solver.py:303-334).  Here every rank maps the other ranks'
// "window" (one cudaMalloc'ed buffer per rank, opened with cudaIpcOpenMemHandle) and the kernels of the iteration
// store straight into the peers' memory:
//
//   all-reduce   fused into the reduction kernel that produces the local sums (k_reduce_ar): the single CTA writes its
//                NV partial sums into slot [parity][my rank] of EVERY rank's window, releases a flag (st.release.sys),
//                spins (ld.acquire.sys) until all nranks slots of its own window carry the epoch, adds them in rank
//                order (bit-identical on every rank) and runs the scalar PCG logic - one launch, no NCCL kernel;
//   halo         k_halo_pack_peer gathers the interface values and stores them into the neighbours' receive buffers
//                (the last CTA to finish releases one flag per neighbour); k_halo_unpack_peer acquires the flags and
//                adds in neighbour order (the reference's summation order, pcg_solver.py:332-334; deterministic).  In
//                the PCG loop the unpack runs on a forked stream beside the p.q all-reduce (both only wait for remote
//                data); an optional interface-first split of the SpMV can put the interior tiles behind the pack.
//
// Epochs come from device memory (graph replays need no new kernel arguments); buffers are double-buffered by the
// epoch parity, which is enough because a rank can run at most one exchange ahead of its slowest peer (every exchange
// needs the peer's contribution to the previous one).  Every spin is bounded: on time-out the kernel raises the sticky
// status word of the communicator and the host reports PCGB_ERR_COMM instead of hanging the GPU.
#pragma once
#include "common.cuh"
#include "pcg_kernels.cuh"

namespace pcgb {

constexpr int kMaxPeers = 16;                 // ranks of one NVLink domain
constexpr int kArSlotWords = 16;              // 128 B per (parity, rank): 8 doubles of data, flag in the second 64-B line
constexpr int kArMaxVals = 8;
constexpr size_t kWinArBytes = (size_t)2 * kMaxPeers * kArSlotWords * 8;   // 4 KB
constexpr size_t kWinBytes = kWinArBytes + 256;                            // + epoch counter, scratch
constexpr long long kSpinLimit = 40000000000ll;                           // ~20 s of SM clocks

// passed by value to the kernels
struct PeerWin {
  int rank = 0, nranks = 1;
  unsigned long long *base[kMaxPeers];   // window of every rank (own entry = local pointer), 8-byte words
  unsigned long long *epoch = nullptr;   // local: number of all-reduces completed
  int *status = nullptr;                 // local, sticky: != 0 after a time-out
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// wait until *flag >= want; false (and status raised) on time-out or when the status is already raised
__device__ __forceinline__ bool spin_until(const unsigned long long *flag, unsigned long long want, int *status) {
  if (ld_acquire_sys(flag) >= want) return true;
  const long long t0 = clock64();
  for (;;) {
    if (ld_acquire_sys(flag) >= want) return true;
    if (*(volatile int *)status != 0) return false;
    if (clock64() - t0 > kSpinLimit) { atomicExch(status, 1); return false; }
    __nanosleep(64);
  }
}

__device__ __forceinline__ unsigned long long *ar_slot(unsigned long long *base, int parity, int rank) {
  return base + ((size_t)parity * kMaxPeers + rank) * kArSlotWords;
}

// The all-reduce proper, executed by ONE CTA (>= nranks threads): in sv[0..nv) the local values (shared memory),
// out in sv[0..nv) the rank-ordered global sums.  All threads of the CTA must call it.
template <int NV>
__device__ __forceinline__ void cta_allreduce(const PeerWin &w, double *sv /* shared, >= NV */, double *sall /* shared, kMaxPeers*NV */) {
  __shared__ unsigned long long s_ep;
  const int tid = threadIdx.x;
  if (tid == 0) s_ep = *w.epoch + 1;
  __syncthreads();
  const unsigned long long ep = s_ep;
  const int par = (int)(ep & 1ull);
  if (tid < w.nranks) {
    // my contribution into slot [par][my rank] of rank `tid` (NVLink store for tid != rank)
    unsigned long long *dst = ar_slot(w.base[tid], par, w.rank);
    double *dd = reinterpret_cast<double *>(dst);
#pragma unroll
    for (int k = 0; k < NV; ++k) dd[k] = sv[k];
    st_release_sys(dst + 8, ep);   // release: the values above are visible before the flag
    // contribution of rank `tid` in my own window
    const unsigned long long *src = ar_slot(w.base[w.rank], par, tid);
    const bool ok = spin_until(src + 8, ep, w.status);
    const double *sd = reinterpret_cast<const double *>(src);
#pragma unroll
    for (int k = 0; k < NV; ++k) sall[tid * NV + k] = ok ? ld_relaxed_sys_f64(sd + k) : 0.0;
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = 0.0;
      for (int r = 0; r < w.nranks; ++r) s += sall[r * NV + k];   // rank order: identical bits on every rank
      sv[k] = s;
    }
    *w.epoch = ep;
  }
  __syncthreads();
}

// Reduction of the per-block partials + all-reduce + the scalar logic of the iteration, one CTA.
// MODE 0: reduce + all-reduce only.  MODE 1: + alpha logic.  MODE 2: + norms logic.
// The exchange always runs (all ranks stay in step even when the PCG state is frozen); the logic only in ST_RUN.
template <int NV, int MODE>
__global__ void __launch_bounds__(256)
k_reduce_ar(PeerWin w, PcgCtrl *ctrl, const double *__restrict__ partials, int count, int row_stride, double *__restrict__ out,
            double *resvec) {
  __shared__ double red[NV * 32];
  __shared__ double sv[kArMaxVals];
  __shared__ double sall[kMaxPeers * NV];
  double v[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = 0.0;
    for (int i = threadIdx.x; i < count; i += 256) s += partials[(size_t)k * row_stride + i];
    v[k] = s;
  }
  block_sum<NV, 256>(v, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) sv[k] = v[k];
  }
  __syncthreads();
  cta_allreduce<NV>(w, sv, sall);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) out[k] = sv[k];
    if (MODE != 0 && ctrl->state == ST_RUN) {
      if (MODE == 1) ctrl_alpha(ctrl, sv[0]);
      if (MODE == 2) ctrl_norms(ctrl, sv[0], sv[1 % NV], sv[2 % NV], sv[3 % NV], sv[4 % NV], resvec);
    }
  }
}

// in-place all-reduce of buf[0..count), count <= kArMaxVals (pcgb_allreduce_sum on the peer transport)
__global__ void __launch_bounds__(32) k_allreduce_peer(PeerWin w, double *buf, int count) {
  __shared__ double sv[kArMaxVals];
  __shared__ double sall[kMaxPeers * kArMaxVals];
  if (threadIdx.x < kArMaxVals) sv[threadIdx.x] = (int)threadIdx.x < count ? buf[threadIdx.x] : 0.0;
  __syncthreads();
  cta_allreduce<kArMaxVals>(w, sv, sall);
  if ((int)threadIdx.x < count) buf[threadIdx.x] = sv[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------- halo over peer memory
struct PeerHalo {
  int n_nbr = 0;
  int64_t m = 0;                       // my shared entries (with multiplicity over neighbours)
  const int *nbr_ptr = nullptr;        // device [n_nbr+1]
  const int *ent_nbr = nullptr;        // device [m]    neighbour slot of entry i
  double *const *remote = nullptr;     // device [n_nbr] start of MY segment inside neighbour j's receive buffer (parity 0)
  const int64_t *remote_m = nullptr;   // device [n_nbr] neighbour j's m (stride between its parity halves)
  unsigned long long *const *remote_flag = nullptr;  // device [n_nbr] neighbour j's flag pair for me: [2]
  double *recv = nullptr;              // local receive buffer [2][m] (peers store into it)
  unsigned long long *flags = nullptr; // local [n_nbr][2]
  unsigned long long *epoch = nullptr; // local: exchanges started (written by the pack kernel)
  unsigned int *done = nullptr;        // local: CTA completion counter of the pack kernel
  int *status = nullptr;
};

// send[i] = y[idx[i]] stored directly into the neighbour's receive buffer; the last CTA releases the flags.
// kPackPerThread entries per thread; ONE system fence per CTA (bar.sync makes the CTA's stores visible to thread 0, whose
// fence.sys is cumulative - the grid-sync idiom) instead of one per thread: the fence waits for the NVLink write acks.
constexpr int kPackPerThread = 4;
__global__ void __launch_bounds__(256)
k_halo_pack_peer(PeerHalo h, const int *__restrict__ idx, const double *__restrict__ y) {
  __shared__ bool s_last;
  const unsigned long long ep = *h.epoch + 1;   // only the last CTA advances it, after every CTA has read it
  const int par = (int)(ep & 1ull);
  const int64_t base = blockIdx.x * (int64_t)(256 * kPackPerThread) + threadIdx.x;
#pragma unroll
  for (int u = 0; u < kPackPerThread; ++u) {
    const int64_t i = base + u * 256;
    if (i < h.m) {
      const int j = h.ent_nbr[i];
      h.remote[j][(size_t)par * (size_t)h.remote_m[j] + (size_t)(i - h.nbr_ptr[j])] = y[idx[i]];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = atomicAdd(h.done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();   // belt and braces: the flag-writing threads learnt of the other CTAs' completion through thread 0
  if ((int)threadIdx.x < h.n_nbr) st_release_sys(h.remote_flag[threadIdx.x] + par, ep);
  if (threadIdx.x == 0) { *h.done = 0; *h.epoch = ep; }
}

// one thread per distinct interface dof: y[dof] += received copies in neighbour order (pcg_solver.py:332-334)
__global__ void __launch_bounds__(256)
k_halo_unpack_peer(PeerHalo h, int64_t ndof, const int *__restrict__ dof, const int *__restrict__ ptr, const int *__restrict__ pos,
                   double *__restrict__ y) {
  __shared__ int s_ok;
  const unsigned long long ep = *h.epoch;       // the exchange the preceding pack kernel started
  const int par = (int)(ep & 1ull);
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if ((int)threadIdx.x < h.n_nbr) {
    if (!spin_until(h.flags + 2 * threadIdx.x + par, ep, h.status)) s_ok = 0;
  }
  __syncthreads();
  if (!s_ok) return;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= ndof) return;
  const double *rv = h.recv + (size_t)par * (size_t)h.m;
  double s = y[dof[t]];
  for (int k = ptr[t]; k < ptr[t + 1]; ++k) s += ld_relaxed_sys_f64(rv + pos[k]);
  y[dof[t]] = s;
}

}  // namespace pcgb
