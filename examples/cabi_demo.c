/* cabi_demo.c - the drop-in boundary used from plain C: no Python, no torch, only include/pcgb200.h + the CUDA runtime.
 *
 * Builds the 27-point operator of config C1 (diagonal 26, off-diagonals -1 on an n^3 grid), b = A x* with a
 * deterministic x*, and solves it with the Jacobi-PCG of libpcgb200.so through the same entry points a ctypes / cgo / JNI
 * binding would use (pcgb_csr_create -> pcgb_csr_diag -> pcgb_reciprocal -> pcgb_solver_create -> pcgb_solve), i.e. what
 * PCG(RefMeshPart) (pcg_solver.py:356-598) + updatePreconditioner (:346-352) do for one rank.
 *
 *   gcc -O2 -Iinclude -I/usr/local/cuda/include examples/cabi_demo.c -o examples/cabi_demo \
 *       -Lpcg_mpi_solver_b200/csrc -lpcgb200 -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/pcg_mpi_solver_b200/csrc
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "pcgb200.h"

#define CHECK_CUDA(c) do { cudaError_t e_ = (c); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA: %s (%s:%d)\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)
#define CHECK_PCGB(c) do { int r_ = (c); if (r_ != PCGB_OK) { fprintf(stderr, "pcgb: %s (rc %d, %s:%d)\n", pcgb_last_error(), r_, __FILE__, __LINE__); return 3; } } while (0)

int main(int argc, char **argv) {
  const int g = argc > 1 ? atoi(argv[1]) : 24;
  const int64_t n = (int64_t)g * g * g;
  int32_t sizes[4];
  pcgb_abi_sizes(sizes);
  if (sizes[0] != (int32_t)sizeof(pcgb_options) || sizes[1] != (int32_t)sizeof(pcgb_result)) { fprintf(stderr, "ABI struct mismatch\n"); return 1; }
  if (pcgb_device_count() == 0) { fprintf(stderr, "no CUDA device: libpcgb200 has no CPU fallback\n"); return 4; }

  /* host CSR of the 27-point operator (sorted columns) */
  int32_t *rowptr = malloc((n + 1) * sizeof(int32_t));
  int32_t *col = malloc(n * 27 * sizeof(int32_t));
  double *val = malloc(n * 27 * sizeof(double));
  int64_t nnz = 0;
  rowptr[0] = 0;
  for (int z = 0; z < g; ++z)
    for (int y = 0; y < g; ++y)
      for (int x = 0; x < g; ++x) {
        for (int dz = -1; dz <= 1; ++dz)
          for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
              const int xx = x + dx, yy = y + dy, zz = z + dz;
              if (xx < 0 || yy < 0 || zz < 0 || xx >= g || yy >= g || zz >= g) continue;
              col[nnz] = (int32_t)(((int64_t)zz * g + yy) * g + xx);
              val[nnz] = (dx == 0 && dy == 0 && dz == 0) ? 26.0 : -1.0;
              ++nnz;
            }
        rowptr[((int64_t)z * g + y) * g + x + 1] = (int32_t)nnz;
      }
  double *xs = malloc(n * sizeof(double)), *b = calloc(n, sizeof(double)), *x = malloc(n * sizeof(double));
  for (int64_t i = 0; i < n; ++i) xs[i] = sin(0.37 * (double)i) + 0.01 * (double)(i % 7);
  for (int64_t i = 0; i < n; ++i)
    for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) b[i] += val[k] * xs[col[k]];

  /* device buffers: owned by the caller, as the C ABI prescribes */
  void *d_rowptr, *d_col, *d_val, *d_b, *d_x, *d_diag, *d_minv;
  CHECK_CUDA(cudaMalloc(&d_rowptr, (n + 1) * sizeof(int32_t)));
  CHECK_CUDA(cudaMalloc(&d_col, nnz * sizeof(int32_t)));
  CHECK_CUDA(cudaMalloc(&d_val, nnz * sizeof(double)));
  CHECK_CUDA(cudaMalloc(&d_b, n * sizeof(double)));
  CHECK_CUDA(cudaMalloc(&d_x, n * sizeof(double)));
  CHECK_CUDA(cudaMalloc(&d_diag, n * sizeof(double)));
  CHECK_CUDA(cudaMalloc(&d_minv, n * sizeof(double)));
  CHECK_CUDA(cudaMemcpy(d_rowptr, rowptr, (n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_col, col, nnz * sizeof(int32_t), cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_val, val, nnz * sizeof(double), cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_b, b, n * sizeof(double), cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemset(d_x, 0, n * sizeof(double)));

  pcgb_csr_t A = NULL;
  pcgb_solver_t S = NULL;
  CHECK_PCGB(pcgb_csr_create(n, n, nnz, d_rowptr, 0, (const int32_t *)d_col, (const double *)d_val, NULL, &A));
  CHECK_PCGB(pcgb_csr_diag(A, (double *)d_diag, NULL));                               /* calcMatVecProd(...,'Preconditioner') */
  CHECK_PCGB(pcgb_reciprocal(n, (const double *)d_diag, (double *)d_minv, NULL));     /* updatePreconditioner :351 */
  CHECK_PCGB(pcgb_solver_create(A, NULL, NULL, &S));
  pcgb_options opt = {0};
  opt.tol = 1e-10; opt.maxiter = 1000; opt.n_global = n; opt.check_every = 16; opt.use_graph = 1; opt.x0_zero = 1;
  pcgb_result res;
  CHECK_PCGB(pcgb_solve(S, (const double *)d_b, (const double *)d_minv, NULL, (double *)d_x, &opt, NULL, &res, NULL));
  CHECK_CUDA(cudaDeviceSynchronize());
  CHECK_CUDA(cudaMemcpy(x, d_x, n * sizeof(double), cudaMemcpyDeviceToHost));

  /* independent check on the host: true residual and error against x* */
  double rr = 0.0, bb = 0.0, ee = 0.0, ss = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    double ax = 0.0;
    for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) ax += val[k] * x[col[k]];
    rr += (b[i] - ax) * (b[i] - ax); bb += b[i] * b[i];
    ee += (x[i] - xs[i]) * (x[i] - xs[i]); ss += xs[i] * xs[i];
  }
  const double relres = sqrt(rr / bb), relerr = sqrt(ee / ss);
  printf("n=%lld nnz=%lld flag=%d iters=%d relres(solver)=%.3e relres(host)=%.3e relerr=%.3e launches=%lld\n", (long long)n, (long long)nnz,
         res.flag, res.iters, res.relres, relres, relerr, (long long)res.launches);
  const int ok = res.flag == 0 && relres <= 1.001e-10 && relerr <= 1e-8 && fabs(res.relres - relres) <= 1e-3 * relres + 1e-16;
  pcgb_solver_destroy(S);
  pcgb_csr_destroy(A);
  cudaFree(d_rowptr); cudaFree(d_col); cudaFree(d_val); cudaFree(d_b); cudaFree(d_x); cudaFree(d_diag); cudaFree(d_minv);
  free(rowptr); free(col); free(val); free(xs); free(b); free(x);
  puts(ok ? "CABI_DEMO_OK" : "CABI_DEMO_FAILED");
  return ok ? 0 : 5;
}
