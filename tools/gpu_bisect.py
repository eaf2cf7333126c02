"""Stage-by-stage GPU bring-up check; each stage is meant to be run under `timeout` (see tools/gpu_bisect.sh)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
stage = sys.argv[1]
t0 = time.time()
def log(*a):
    print(f"[{stage} +{time.time()-t0:6.1f}s]", *a, flush=True)
import torch
log("torch imported", torch.cuda.is_available())
from oracle import ref_pcg as R
from pcg_mpi_solver_b200 import _lib
from pcg_mpi_solver_b200.csr import CsrMatrix
dev = torch.device("cuda:0")
torch.zeros(1, device=dev); torch.cuda.synchronize(); log("cuda ctx up")
A = R.poisson27(16)
x = np.random.default_rng(0).standard_normal(A.shape[0])
if stage in ("spmv", "spmv_tma0"):
    M = CsrMatrix.from_scipy(A, device=dev); torch.cuda.synchronize(); log("plan", M.plan_info())
    y = M.spmv(torch.from_numpy(x).to(dev)); torch.cuda.synchronize()
    log("spmv err", float(np.abs(y.cpu().numpy() - A @ x).max()))
elif stage.startswith("solve"):
    from pcg_mpi_solver_b200 import solve
    b = A @ x
    kw = dict(use_graph=stage.endswith("graph"), check_every=8)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out = solve(A, b, 1.0 / A.diagonal(), 1e-10, 1000, **kw)
    log("solve", out[1:], float(np.linalg.norm(out[0] - x) / np.linalg.norm(x)))
elif stage == "solve_default_stream_graph":
    from pcg_mpi_solver_b200 import solve
    b = A @ x
    out = solve(A, b, 1.0 / A.diagonal(), 1e-10, 1000, use_graph=True)
    log("solve", out[1:])
log("done")
