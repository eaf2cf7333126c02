#!/usr/bin/env python
"""ORACLE tool (test infrastructure): residual histories of the bench workload, for bench.py's `parity` block.

For every GPU count N in {1, 2, 4, 8} the benchmark's GLOBAL mesh (128^3 elements per GPU stacked 1x1x1 / 2x1x1 / 2x2x1 /
2x2x2, clamped x = 0, traction on x = max) is solved for ITERS Jacobi-PCG iterations by the restated reference
(oracle/ref_pcg.py <- pcg_solver.py:356-598, element-by-element operator <- :242-336) run SPMD on 8 host processes
(oracle/spmd.py), and ||r_k||, k = 0..ITERS, plus ||b|| are written to tests/golden/hex{B}_N{N}_resvec.json.

The residual history of the global problem does not depend on how the mesh is cut into parts beyond fp64 round-off
(the reference itself: 6e-11 on the solution across partitionings), so the 8-box CPU run is the golden for the
N-GPU run of the same global mesh.  bench.py compares its first min(K, ITERS)+1 recorded residual norms with these.

    python oracle/make_golden_resvec.py [--block 128] [--iters 40] [--gpus 1 2 4 8]
"""
import argparse
import json
import os
import sys
import time

for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
    os.environ.setdefault(_v, "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _rank(rank, size, comm, ng, pgrid, iters):
    from oracle import ref_pcg as R
    from oracle.hex_parts import hex_box_part_spmd
    from pcg_mpi_solver_b200.hexmesh import partition_blocks
    blocks = partition_blocks(ng, pgrid)
    part = R.EbePart(hex_box_part_spmd(blocks, rank, h=1.0 / ng[0]))
    comm.setup_halo(part, comm._halo_box)
    R.update_bc([part], comm=comm)
    minv = R.Operator([part], comm).jacobi()
    nglob = 3 * ng[0] * (ng[1] + 1) * (ng[2] + 1)
    resvec = []
    out = R.ref_pcg([part], minv, 1e-300, iters, nglob=nglob, comm=comm, resvec=resvec)
    return {"resvec": [float(v) for v in resvec], "normb": out["normb"], "flag": out["Flag"], "iter": out["Iter"], "n": part.n}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--block", type=int, default=128)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--gpus", type=int, nargs="*", default=[1, 2, 4, 8])
    ap.add_argument("--procs", type=int, default=8)
    args = ap.parse_args()
    from oracle.spmd import run_spmd
    from pcg_mpi_solver_b200.hexmesh import block_grid
    for n in args.gpus:
        pg = block_grid(n)
        ng = tuple(args.block * pg[a] for a in range(3))
        p = args.procs
        while p > 1 and any(g > m for g, m in zip(block_grid(p), ng)):
            p //= 2
        t0 = time.time()
        res = run_spmd(p, _rank, (ng, block_grid(p), args.iters))
        r0 = res[0]
        assert all(r["resvec"] == r0["resvec"] for r in res), "ranks disagree on the residual history"
        out = {"workload": f"hex{args.block}^3 elements per GPU, global {ng[0]}x{ng[1]}x{ng[2]}", "n_gpus": n, "ng": list(ng),
               "n_global": 3 * ng[0] * (ng[1] + 1) * (ng[2] + 1), "normb": r0["normb"], "resvec": r0["resvec"],
               "oracle": f"oracle/ref_pcg.py (EBE operator) SPMD on {p} host processes, boxes {list(block_grid(p))}",
               "generator": "oracle/make_golden_resvec.py", "seconds": time.time() - t0}
        path = os.path.join(ROOT, "tests", "golden", f"hex{args.block}_N{n}_resvec.json")
        with open(path, "w") as f:
            json.dump(out, f, indent=1)
        print(f"N={n} ng={ng} procs={p} {time.time() - t0:.1f}s  normb={r0['normb']:.6e} r0={r0['resvec'][0]:.6e} r_last={r0['resvec'][-1]:.6e}", flush=True)


if __name__ == "__main__":
    main()
