/* pcgb200.h - C ABI of libpcgb200.so: the B200 (sm_100a) implementation of the PCG inner
 * iteration of ankitskr/PCG-MPI-solver (reference: src/solver/pcg_solver.py).
 *
 * The reference is pure Python (numpy + mpi4py) and has no FFI of its own; the functions
 * below are what a ctypes binding of its hot path binds.  Each entry point cites the
 * reference code it replaces (file:line in /root/reference).
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer owned by the caller (torch tensors in the
 *     Python host code); the library never copies between host and device behind the
 *     caller's back, except the few control scalars pcgb_solve reads back to steer the loop;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - return value: 0 = ok, negative = error (see pcgb_last_error()); the MATLAB-style
 *     solver status (0..4, pcg_solver.py:399,449,468,477,493,497,541,561) is returned
 *     separately through pcgb_result;
 *   - handles are opaque; a handle may be used by one host thread at a time.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef PCGB200_H
#define PCGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCGB_VERSION 200 /* 0.2.0 */

typedef struct pcgb_csr_s *pcgb_csr_t;     /* device CSR matrix + merge-path SpMV plan          */
typedef struct pcgb_comm_s *pcgb_comm_t;   /* communicator: peer-memory windows (+ NCCL), one rank = one GPU */
typedef struct pcgb_halo_s *pcgb_halo_t;   /* interface ("halo") exchange-add plan              */
typedef struct pcgb_solver_s *pcgb_solver_t; /* PCG workspace bound to one operator             */
typedef struct pcgb_ebe_s *pcgb_ebe_t;     /* opt-in matrix-free element-by-element operator (f1) */
typedef struct pcgb_asm_s *pcgb_asm_t;     /* device assembly of K_i[Eff,Eff] from the pattern groups (symbolic phase state) */
typedef struct pcgb_ebe2_s *pcgb_ebe2_t;   /* coloured (atomics-free, bit-reproducible) variant of the EBE operator */

/* error codes */
#define PCGB_OK 0
#define PCGB_ERR_ARG (-1)
#define PCGB_ERR_CUDA (-2)
#define PCGB_ERR_NCCL (-3)
#define PCGB_ERR_NODEVICE (-4)
#define PCGB_ERR_COMM (-5) /* a peer-memory exchange timed out waiting for another rank */

int pcgb_version(void);
/* sizeof(pcgb_options), sizeof(pcgb_result), sizeof(pcgb_hex_box), sizeof(pcgb_ebe_group): lets a binding verify its struct layouts */
void pcgb_abi_sizes(int32_t out[4]);
const char *pcgb_last_error(void); /* thread-local message of the last failing call */
int pcgb_device_count(void);       /* number of visible CUDA devices (0 on a CPU box) */

/* ---------------------------------------------------------------- CSR operator (a2, a7)
 * Replaces calcMatVecProd(..., 'Strain') (pcg_solver.py:242-300) on the free ("Eff") dofs:
 * A = K[Eff,Eff] assembled to CSR (SURVEY Fact 1), so the P_unq[LocDofEff] scatter /
 * Q_unq[LocDofEff] gather around the matvec (pcg_solver.py:482-484) disappear.
 * fp64 values, int32 column indices, row offsets int32 or int64 (int64 is mandatory once
 * nnz >= 2^31, e.g. the 256^3 block).  Arrays stay owned by the caller and must outlive the
 * handle.  pcgb_csr_create builds the merge-path tile plan (device-side searches).
 */
int pcgb_csr_create(int64_t nrows, int64_t ncols, int64_t nnz, const void *d_rowptr, int rowptr_is_64,
                    const int32_t *d_col, const double *d_val, void *stream, pcgb_csr_t *out);
int pcgb_csr_destroy(pcgb_csr_t A);
/* y = A x  (one launch of the merge-path kernel, plus a fix-up launch only if rows are split) */
int pcgb_spmv(pcgb_csr_t A, const double *d_x, double *d_y, void *stream);
/* d[i] = A[i,i]  (Jacobi diagonal, replaces calcMatVecProd(...,'Preconditioner'), pcg_solver.py:282-287) */
int pcgb_csr_diag(pcgb_csr_t A, double *d_diag, void *stream);
/* The persistent staged-x SpMV kernel reads its own 16-bit index stream, never d_col.  pcgb_csr_release_col caches the
 * diagonal and drops the library's reference to d_col; the caller may then free it (2 GB at 128^3, 16 GB at 256^3).
 * PCGB_ERR_ARG when the selected kernel still needs the column array. */
int pcgb_csr_release_col(pcgb_csr_t A, void *stream);
/* Interface-first split of the SpMV for the multi-GPU overlap ("order boundary rows first, launch exchange,
 * compute interior rows"): d_rows = distinct local rows taking part in the interface exchange (pcg_solver.py:304-312).
 * Tiles owning such a row run in a first launch, all others in a second one.  pcgb_solver_create registers the rows
 * of its halo plan itself; these two entry points make the split testable on one GPU.  d_out_dot (may be NULL) = x.y */
int pcgb_csr_set_boundary_rows(pcgb_csr_t A, const int32_t *d_rows, int64_t count, void *stream);
int pcgb_spmv_split(pcgb_csr_t A, const double *d_x, double *d_y, double *d_out_dot, void *stream);
/* Algorithmic bytes of one SpMV: 12*nnz + R*(nrows+1) + 8*ncols + 8*nrows. */
int64_t pcgb_spmv_bytes(pcgb_csr_t A);
/* Bytes the selected kernel really streams from HBM (the staged-x kernel reads 16-bit local column
 * indices: 10 B per non-zero instead of 12; with the column-triple index 8 + 2/3 B). */
int64_t pcgb_spmv_stream_bytes(pcgb_csr_t A);
/* plan introspection for tests / DESIGN.md: tiles, tile items, lanes, snap, split rows, smem bytes,
 * max row, tma, staged, x windows, staged doubles per tile (cap), max windows per tile, column-triple index (0/1),
 * interface tiles (-1 = no split registered), resident CTAs of the persistent kernel, column array released (0/1) */
int pcgb_csr_plan_info(pcgb_csr_t A, int64_t info[16]);

/* ---------------------------------------------------------------- vector kernels (a4-a10)
 * out[0] = sum_i a[i]*b[i]*w[i]   (w may be NULL = all ones).  np.dot(a, b*w) of
 * pcg_solver.py:381,415,462,487,504-506,532,573.  Deterministic two-level reduction.  */
int pcgb_dot_w(int64_t n, const double *d_a, const double *d_b, const double *d_w, double *d_out, void *stream);
/* y[i] = a*x[i] + b*y[i] */
int pcgb_axpby(int64_t n, double a, const double *d_x, double b, double *d_y, void *stream);
/* z[i] = x[i]*y[i] */
int pcgb_mul(int64_t n, const double *d_x, const double *d_y, double *d_z, void *stream);
/* out[i] = 1/d[i]  (updatePreconditioner, pcg_solver.py:351) */
int pcgb_reciprocal(int64_t n, const double *d_d, double *d_out, void *stream);

/* ---------------------------------------------------------------- communicator (a13)
 * Replaces mpi4py COMM_WORLD (pcg_solver.py:968-970).  Two transports for the data path:
 *   PEER (default)  the library's own kernels over CUDA-IPC mapped peer memory (NVLink / NVSwitch): the all-reduce is
 *                   fused into the reduction kernel of the iteration, the halo values are stored straight into the
 *                   neighbours' receive buffers (csrc/peer.cuh).  Needs the window exchange below.
 *   NCCL            ncclAllReduce / ncclSend / ncclRecv (baseline and fallback; PCGB_COMM=nccl selects it).
 * NCCL is resolved at run time with dlopen("libnccl.so.2") so the library loads on a box without NCCL.  Rank 0 obtains a
 * unique id and the host code broadcasts it (torch.distributed / mpi4py / any side channel).  id == NULL creates a
 * peer-only communicator (no NCCL at all; at most 16 ranks).                                                     */
#define PCGB_UNIQUE_ID_BYTES 128
#define PCGB_IPC_BLOB_BYTES 64
#define PCGB_TRANSPORT_NCCL 0
#define PCGB_TRANSPORT_PEER 1
int pcgb_comm_unique_id(unsigned char id[PCGB_UNIQUE_ID_BYTES]);
int pcgb_comm_create(int rank, int nranks, const unsigned char *id /* [PCGB_UNIQUE_ID_BYTES] or NULL */, pcgb_comm_t *out);
int pcgb_comm_destroy(pcgb_comm_t c);
/* window exchange: export -> all-gather the blobs over the host side channel (rank order) -> import */
int pcgb_comm_window_export(pcgb_comm_t c, unsigned char blob[PCGB_IPC_BLOB_BYTES]);
int pcgb_comm_window_import(pcgb_comm_t c, const unsigned char *blobs /* nranks * PCGB_IPC_BLOB_BYTES */);
int pcgb_comm_transport(pcgb_comm_t c);                  /* transport the data path uses right now */
int pcgb_comm_set_transport(pcgb_comm_t c, int transport);
int pcgb_comm_status(pcgb_comm_t c, void *stream);       /* 0 healthy; != 0 after a peer exchange timed out (sticky) */
/* in-place sum of `count` doubles over all ranks (MPI_SUM, pcg_solver.py:622-628); rank-ordered, bit-identical on every rank */
int pcgb_allreduce_sum(pcgb_comm_t c, double *d_buf, int count, void *stream);

/* ---------------------------------------------------------------- halo exchange-add (a14)
 * Replaces the pack / Isend / Recv / Waitall / += of pcg_solver.py:303-334.
 *   nbr_rank[j]        neighbour j's rank              (NbrMPIdVector)
 *   nbr_ptr[j..j+1]    range of neighbour j in idx     (host array, n_nbr+1 entries)
 *   idx[k]             local row index (in the Eff numbering of A) of the k-th shared dof,
 *                      ordered by ascending GLOBAL node id then direction on both sides
 *                      (OvrlpLocalDofVecList, partition_mesh.py:822-827)
 * The unpack adds the received values in neighbour order j = 0..n_nbr-1, i.e. the same
 * summation order as the reference loop at pcg_solver.py:333-334 (deterministic).        */
int pcgb_halo_create(pcgb_comm_t c, int n_nbr, const int32_t *nbr_rank, const int64_t *nbr_ptr,
                     const int64_t *idx_host, int64_t nlocal, pcgb_halo_t *out);
int pcgb_halo_destroy(pcgb_halo_t h);
/* peer transport: every rank exports its receive block, the host all-gathers the blobs (pcgb_halo_blob_bytes each, rank
 * order) and imports them; the import checks that both sides of every interface list the same number of shared dofs */
int64_t pcgb_halo_blob_bytes(pcgb_halo_t h);
int pcgb_halo_export(pcgb_halo_t h, unsigned char *blob);
int pcgb_halo_import(pcgb_halo_t h, const unsigned char *blobs);
/* y[idx] += (values of the same dofs on the neighbours) */
int pcgb_halo_exchange_add(pcgb_halo_t h, double *d_y, void *stream);
int64_t pcgb_halo_bytes(pcgb_halo_t h); /* bytes sent (= received) per exchange by this rank */

/* ---------------------------------------------------------------- PCG (a1, a4-a13)
 * Replaces PCG(RefMeshPart) (pcg_solver.py:356-598): Jacobi-PCG with MATLAB `pcg`
 * semantics on the free dofs of one subdomain.                                         */
typedef struct pcgb_options {
  double tol;           /* GlobData['Tol']                                                 */
  int32_t maxiter;      /* GlobData['MaxIter']                                             */
  int64_t n_global;     /* GlobData['GlobNDofEff'] (only used for MaxMSteps, :404)          */
  int32_t max_stag;     /* MaxStagSteps = 3 (:403); <=0 selects 3                           */
  int32_t check_every;  /* iterations enqueued between host polls of the device state; <=0 = 16 */
  int32_t use_graph;    /* capture the batch of iterations in a CUDA graph (1) or launch directly (0) */
  int32_t fixed_iters;  /* benchmark mode: ignore convergence, run exactly maxiter iterations */
  int32_t record_resvec; /* keep ||r|| per iteration (reference has this commented out, :428-434) */
  int32_t time_kernels; /* bracket every SpMV launch of the loop with CUDA events (forces direct launches) */
  int32_t x0_zero;      /* the caller guarantees d_x is all zeros on entry: r0 = b exactly, the initial operator application is skipped */
} pcgb_options;

typedef struct pcgb_result {
  int32_t flag;       /* 0 converged, 1 maxiter, 2 precond inf, 3 stagnation, 4 breakdown       */
  int32_t iters;      /* `Iter` after the +1 of pcg_solver.py:584                                */
  double relres;      /* RelRes                                                                  */
  double normb;       /* NormRefLoadVector (n2b)                                                 */
  int32_t imin;       /* iMin                                                                    */
  int32_t stag;       /* Stag at exit                                                            */
  int32_t moresteps;  /* MoreSteps at exit                                                       */
  int32_t too_small_tol; /* 1 if the reference would have raised Warning('PCG : TooSmallTolerance') (:549) */
  int64_t matvecs;    /* operator applications performed                                         */
  int64_t launches;   /* CUDA kernel launches issued by this solve (bench 'gpu_launches')        */
  double loop_ms;     /* device time of the iteration loop only (CUDA events on the solver stream)  */
  double spmv_ms;     /* sum of the SpMV launch durations inside the loop (time_kernels only)        */
  int64_t spmv_timed; /* number of SpMV launches in spmv_ms                                         */
  int64_t loop_iters; /* iterations executed inside the timed loop                                  */
  double setup_ms;    /* device time from the entry of the solve to the start of the loop (||b||, r0, rho0) */
  double final_ms;    /* device time from the end of the loop to the copy-out of x (finalisation matvec)    */
  double phase_ms[8]; /* time_kernels: sums over the bracketed iterations of p-update | SpMV | halo pack | p.q reduce+all-reduce |
                         halo unpack-add | fused update | norms reduce+all-reduce | whole iteration                  */
} pcgb_result;

int pcgb_solver_create(pcgb_csr_t A, pcgb_halo_t halo /* may be NULL */, pcgb_comm_t comm /* may be NULL */,
                       pcgb_solver_t *out);
int pcgb_solver_destroy(pcgb_solver_t s);
/* d_b: right-hand side Fext[LocDofEff]; d_minv: inverse diagonal (InvDiagPreCondVector0) or
 * NULL = identity (ExistDP0 False); d_w: DofWeightVector_Eff or NULL = ones;
 * d_x: in = initial guess X0, out = solution (X, or XMin on the non-converged path, :569-582).
 * d_resvec: optional device buffer of maxiter+2 doubles when record_resvec is set.          */
int pcgb_solve(pcgb_solver_t s, const double *d_b, const double *d_minv, const double *d_w, double *d_x,
               const pcgb_options *opt, double *d_resvec, pcgb_result *res, void *stream);
/* y = A x followed by the interface sum: calcMPFint (pcg_solver.py:339-342) on Eff dofs */
int pcgb_apply(pcgb_solver_t s, const double *d_x, double *d_y, void *stream);

/* ---------------------------------------------------------------- opt-in: matrix-free EBE operator (f1)
 * The reference's own operator form: calcMatVecProd(...,'Strain'), pcg_solver.py:263-300, on the GPU, one
 * pattern group per reference type group (partition_mesh.py:470-491).  Parity-tested on B200 (tests/test_gpu_ebe.py), profiled
 * (profiles/ncu_ebe_t24_r2.txt); opt-in - the assembled CSR path is the default.  d_idx is [nd][ne] int32 in the FREE-dof numbering
 * (-1 = clamped dof), d_sign [nd][ne] uint8 or NULL, d_ck [ne], ke_host the nd x nd pattern matrix (host).  */
typedef struct pcgb_ebe_group {
  int32_t nd;
  int64_t ne;
  const int32_t *d_idx;
  const unsigned char *d_sign;
  const double *d_ck;
  const double *ke_host;
} pcgb_ebe_group;
int pcgb_ebe_create(int64_t n, int ngroups, const pcgb_ebe_group *groups, pcgb_ebe_t *out);
int pcgb_ebe_destroy(pcgb_ebe_t E);
int pcgb_ebe_apply(pcgb_ebe_t E, const double *d_x, double *d_y, void *stream); /* y = K_i x (no interface sum) */
int64_t pcgb_ebe_bytes(pcgb_ebe_t E);                                            /* algorithmic bytes per application */
int pcgb_solver_create_ebe(pcgb_ebe_t E, pcgb_halo_t halo /* may be NULL */, pcgb_comm_t comm /* may be NULL */,
                           pcgb_solver_t *out);

/* ---------------------------------------------------------------- device assembly of the CSR operator (f2)
 * A_i = K_i[Eff,Eff] = sum_e P_e^T (Ck_e S_e Ke S_e) P_e  from the reference's pattern type groups (partition_mesh.py:443-491;
 * the operator calcMatVecProd applies element by element, pcg_solver.py:263-300), assembled on the device with hand-written
 * kernels (csrc/assemble.cuh): incidence lists per dof -> per-row sorted distinct columns (symbolic) -> per-row
 * accumulation in a fixed element order (numeric).  No atomics in the arithmetic: bit-reproducible.  `groups` uses the same
 * layout as the EBE operator (d_idx in the free-dof numbering, -1 = clamped).  Two phases because the caller owns the
 * output arrays: symbolic fills d_rowptr[0..n] (int64) and reports nnz; numeric fills d_col / d_val (ascending columns). */
int pcgb_assemble_symbolic(int64_t n, int ngroups, const pcgb_ebe_group *groups, int64_t *d_rowptr, int64_t *nnz_out, void *stream,
                           pcgb_asm_t *out);
int pcgb_assemble_numeric(pcgb_asm_t a, const int64_t *d_rowptr, int32_t *d_col, double *d_val, void *stream);
int pcgb_assemble_destroy(pcgb_asm_t a);

/* Operator-level (not reachable from pcgb_solve; green on B200, tests/test_gpu_ebe_colored.py): the same operator
 * with an atomics-free deterministic scatter.  groups[] = one entry per (pattern group, colour) slice, sorted by
 * colour; phase[g] = colour of entry g; no two elements of one colour may share a dof (coloring.py).            */
int pcgb_ebe2_create(int64_t n, int ngroups, const pcgb_ebe_group *groups, const int32_t *phase, pcgb_ebe2_t *out);
int pcgb_ebe2_destroy(pcgb_ebe2_t E);
int pcgb_ebe2_apply(pcgb_ebe2_t E, const double *d_x, double *d_y, void *stream);
int pcgb_ebe2_launches(pcgb_ebe2_t E); /* kernel launches per application */

/* ---------------------------------------------------------------- structured hex generator
 * On-device generator of the sub-assembled stiffness matrix of one box of trilinear hex
 * elements (benchmark configs C2/C3/C5).  The box holds elements
 * [e0[a], e0[a]+ne[a]) of a global ng[0] x ng[1] x ng[2] mesh; nodes with global x index 0
 * are clamped (removed from the Eff numbering).  ke is the 24x24 element matrix (host
 * pointer, row-major, local node l at offsets (l&1, l>>1&1, l>>2&1), 3 dofs per node) and ck
 * the per-element scale (ElemList_Ck = E*h).  Free dof numbering: x fastest, then y, z, then
 * direction.  pcgb_hex_count writes rowcount[i+1] (int64, rowcount[0] = 0) so that an
 * inclusive scan by the caller yields the row offsets.                                     */
typedef struct pcgb_hex_box {
  int32_t ng[3];
  int32_t e0[3];
  int32_t ne[3];
} pcgb_hex_box;
int64_t pcgb_hex_nrows(const pcgb_hex_box *box);
int pcgb_hex_count(const pcgb_hex_box *box, int64_t *d_rowcount, void *stream);
int pcgb_hex_fill(const pcgb_hex_box *box, const double *ke_host, double ck, const int64_t *d_rowptr,
                  int32_t *d_col, double *d_val, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PCGB200_H */
