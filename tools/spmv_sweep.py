"""SpMV-only timing sweep over plan knobs (env PCGB_SPMV_*), 128^3 hex box.  Prints GB/s per config."""
import itertools, json, os, sys, time
sys.path.insert(0, ".")
import torch
from pcg_mpi_solver_b200.hexmesh import HexBlock, generate_matrix
from pcg_mpi_solver_b200.csr import CsrMatrix
blk_n = int(os.environ.get("SWEEP_BLOCK", "128"))
dev = torch.device("cuda:0")
blk = HexBlock((blk_n,) * 3, (0, 0, 0), (blk_n,) * 3, h=1.0 / blk_n)
base = generate_matrix(blk, device=dev)
x = torch.randn(base.shape[1], dtype=torch.float64, device=dev)
y = torch.empty_like(x)
configs = json.loads(os.environ.get("SWEEP_CONFIGS", "[]")) or [
    {"PCGB_SPMV_LANES": l, "PCGB_SPMV_TILE": t, "PCGB_SPMV_STAGES": sg, "PCGB_SPMV_CTAS": c}
    for c, sg, t, l in itertools.product([2], [4], [1792, 2048, 2176, 2304, 2432], [8, 16])] + [
    {"PCGB_SPMV_LANES": 8, "PCGB_SPMV_TILE": 2048, "PCGB_SPMV_STAGES": 4, "PCGB_SPMV_CTAS": 1},
    {"PCGB_SPMV_LANES": 8, "PCGB_SPMV_TILE": 4096, "PCGB_SPMV_STAGES": 4, "PCGB_SPMV_CTAS": 1},
    {"PCGB_SPMV_LANES": 8, "PCGB_SPMV_TILE": 1280, "PCGB_SPMV_STAGES": 4, "PCGB_SPMV_CTAS": 3},
    {"PCGB_SPMV_LANES": 8, "PCGB_SPMV_TILE": 1024, "PCGB_SPMV_STAGES": 8, "PCGB_SPMV_CTAS": 2}]
for cfg in configs:
    for k, v in cfg.items():
        os.environ[k] = str(v)
    M = CsrMatrix(base.rowptr, base.col, base.val, base.shape)
    for _ in range(3):
        M.spmv(x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        M.spmv(x, out=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({**cfg, "ms": round(ms, 4), "GBps": round(M.spmv_bytes() / ms / 1e6, 1), "smem": M.plan_info()["smem_bytes"], "staged": M.plan_info()["staged"], "xcap": M.plan_info()["x_cap"], "maxw": M.plan_info()["max_windows_per_tile"]}), flush=True)
    del M
