#!/usr/bin/env python
"""bench.py - PCG iterations/s and SpMV GB/s (fp64) on B200, next to the CPU reference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--block 128]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1] at N=1, the C5 stacking rule at N>1, weak scaling): every GPU owns
one box of `block`^3 trilinear hex elements (default 128^3: n = 6 390 144 free dofs, nnz = 509 597 550,
6.1 GB of CSR per GPU - far larger than the 126 MB L2) of a global mesh stacked 1x1x1 / 2x1x1 / 2x2x1 /
2x2x2, clamped at x = 0, traction on x = max; the matrix is generated on the device.  A "step" is one
PCG iteration (CSR SpMV + 2 reductions + Jacobi + AXPYs [+ halo exchange + allreduces]).

    value    K iterations / device time of the iteration loop (CUDA events on the solver stream, max over ranks)
    e2e      the same K iterations through the public solve() with HOST buffers: b from pinned host memory,
             x back to the host, setup / verification matvecs and all host polling inside the timed region
    roofline the merge-path SpMV kernel: algorithmic bytes (12 nnz + 4|8 (n+1) + 16 n) / mean launch duration
             (event pairs around every SpMV launch of a separate K-iteration pass) against MEASURED_PEAKS.json
    cpu_baseline / --impl reference: the oracle port of the reference's numpy element-by-element PCG
             (oracle/ref_pcg.py <- pcg_solver.py:242-598) on the same mesh, one process per part like
             `mpiexec -np P`, bounded to a few iterations.
"""
from __future__ import annotations

import argparse
import json
import os

# the reference pins every BLAS to one thread per rank BEFORE numpy is imported (pcg_solver.py:10-15);
# the CPU arm forks one process per mesh part, so the same must hold here
for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
    os.environ.setdefault(_v, "1")
# stdout carries exactly one JSON line: NCCL's own banner / debug output goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "PCG iterations/sec (fp64 Jacobi-PCG, CSR SpMV) on 3-D elastostatic hex mesh"
UNIT = "iterations/s"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one SpMV launch of the default workload, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json); None when no capture has been recorded."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)["spmv_dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# --------------------------------------------------------------------------------------- CPU reference arm
def _cpu_rank(rank, size, comm, ng, pgrid, iters):
    """One 'MPI rank' of the oracle port (oracle/spmd.py): its own box of the mesh as an element-by-element part,
    reductions and the interface exchange through shared memory + barriers - the reference's communication pattern
    (3 allreduces + 1 neighbour exchange per iteration, pcg_solver.py:303-334, 622-628)."""
    from oracle import ref_pcg as R
    from oracle.hex_parts import hex_box_part_spmd
    from pcg_mpi_solver_b200.hexmesh import partition_blocks
    blocks = partition_blocks(ng, pgrid)
    part = R.EbePart(hex_box_part_spmd(blocks, rank, h=1.0 / ng[0]))
    comm.setup_halo(part, comm._halo_box)
    R.update_bc([part], comm=comm)
    minv = R.Operator([part], comm).jacobi()
    nglob = 3 * (ng[0]) * (ng[1] + 1) * (ng[2] + 1)
    kw = dict(nglob=nglob, comm=comm)
    R.ref_pcg([part], minv, 1e-300, 2, **kw)                 # untimed warm-up (page faults, BLAS init)
    t0 = time.perf_counter()
    R.ref_pcg([part], minv, 1e-300, 1, **kw)
    t1 = time.perf_counter()
    out = R.ref_pcg([part], minv, 1e-300, 1 + iters, **kw)
    t2 = time.perf_counter()
    # difference of two runs = `iters` loop iterations only (set-up and the two residual matvecs cancel)
    return ((t2 - t1) - (t1 - t0), out["Iter"], part.n)


def cpu_reference(ng, iters, max_procs=None):
    """iterations/s of the oracle port on this host: the mesh is cut into P boxes (P = cores, power of two, <= 64),
    one process per part and one BLAS thread per process exactly like the reference's `mpiexec -np P` with
    OMP_NUM_THREADS=1 (pcg_solver.py:10-15); the parts are coupled like the reference's ranks (shared-memory
    allreduce and interface exchange, oracle/spmd.py)."""
    from oracle.spmd import run_spmd
    from pcg_mpi_solver_b200.hexmesh import block_grid
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    p = 1
    while p * 2 <= min(cores, max_procs or 64):
        p *= 2
    while p > 1 and any(g > n for g, n in zip(block_grid(p), ng)):
        p //= 2
    res = run_spmd(p, _cpu_rank, (ng, block_grid(p), iters))
    dt = max(r[0] for r in res)
    return {"value": iters / dt, "unit": UNIT, "cores": p, "kind": "port",
            "sample": f"{iters} PCG loop iterations (difference of a {iters}+1 and a 1 iteration run) of the numpy element-by-element reference path "
                      f"(oracle/ref_pcg.py <- pcg_solver.py:242-598) on the same {ng[0]}x{ng[1]}x{ng[2]} hex mesh cut into {p} boxes, 1 process/box, "
                      f"1 BLAS thread each, shared-memory allreduce + interface exchange every iteration",
            "seconds": dt}


# --------------------------------------------------------------------------------------- main
def main():
    import faulthandler
    faulthandler.enable()
    faulthandler.dump_traceback_later(int(os.environ.get("PCGB_BENCH_WATCHDOG", "240")), exit=False)  # stacks on stderr if a phase hangs
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--block", type=int, default=int(os.environ.get("PCGB_BENCH_BLOCK", "128")), help="hex elements per axis per GPU")
    ap.add_argument("--cpu-iters", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="hex", choices=["hex", "hex_metis", "concrete"],
                    help="hex = block-partitioned device-generated mesh (default, C2/C5); hex_metis = same global mesh through "
                         "partition_mesh()/METIS (C3); concrete = data/concrete.zip through partition_mesh()/METIS (C4)")
    ap.add_argument("--operator", default="csr", choices=["csr", "ebe"],
                    help="csr = assembled merge-path SpMV (the north-star path, default); ebe = opt-in matrix-free operator (f1)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = args.steps, max(args.warmup, 3)
    CE = 50 if world == 1 else 25   # iterations per CUDA graph / host poll

    from pcg_mpi_solver_b200.hexmesh import block_grid
    # the CPU arm works on the SAME global mesh as the N-GPU arm even when it is started without torchrun
    mesh_world = max(world, args.gpus if args.impl == "reference" else 1, 1)
    pgrid = block_grid(mesh_world)
    ng = tuple(args.block * pgrid[a] for a in range(3))
    config = {"workload": f"hex{args.block}^3 elements per GPU, global {ng[0]}x{ng[1]}x{ng[2]} trilinear hex elastostatics "
                          f"(E=1, nu=0.3, h=1/{ng[0]}), clamped x=0, traction on x=max, Jacobi-PCG fixed {K} iterations",
              "per_gpu_block": args.block, "process_grid": list(pgrid), "parallelism": f"dd{mesh_world}",
              "l2_policy": "inputs larger than L2 (CSR 6.1 GB per GPU vs 126 MB L2), no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        iters = max(3, min(args.steps, args.cpu_iters))
        base = cpu_reference(ng, iters)
        line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": iters,
                "warmup": 0, "ms_per_step": 1e3 / base["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    t_start = time.time()

    def log(msg):
        if rank == 0:
            print(f"[bench +{time.time() - t_start:6.1f}s] {msg}", file=sys.stderr, flush=True)

    import torch
    import torch.distributed as dist
    from pcg_mpi_solver_b200 import solve
    from pcg_mpi_solver_b200.hexmesh import generate_matrix, interface_lists, load_vector, partition_blocks
    from pcg_mpi_solver_b200.solver import Communicator, SubdomainOperator

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the b200 arm has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = Communicator.from_torch_distributed(dev)

    log(f"process group / communicator up (world {world})")
    if args.workload == "hex":
        blocks = partition_blocks(ng, pgrid)
        blk = blocks[rank]
        blk.h = 1.0 / ng[0]
        for b_ in blocks:
            b_.h = blk.h
        if args.operator == "ebe":
            from pcg_mpi_solver_b200.hexmesh import generate_ebe
            A = generate_ebe(blk, device=dev)
        else:
            A = generate_matrix(blk, device=dev)
        nbr, lists, w = interface_lists(blocks, rank) if world > 1 else ([], [], None)
        n_global = 3 * ng[0] * (ng[1] + 1) * (ng[2] + 1)
        op = SubdomainOperator(A, comm, nbr, lists, w, n_global=n_global)
        b = load_vector(blk, device=dev)
    else:
        # the general pipeline: model -> METIS (run_metis.py) -> subdomain builder (partition_mesh.py) -> device assembly
        from pcg_mpi_solver_b200.partition import partition_mesh
        if args.workload == "concrete":
            zp = os.path.join(ROOT, "oracle", "_ref", "concrete.zip")
            if not os.path.exists(zp):
                zp = "/root/reference/data/concrete.zip"
            ep = None
            gold = os.path.join(ROOT, "tests", "golden", f"concrete_elepart_{world}.npy")
            if os.path.exists(gold):
                ep = np.load(gold).astype(np.int64)       # the partition the reference's 8-rank golden run used
            subs = partition_mesh(zp, world, elepart=ep, assemble=False)
            config["workload"] = f"data/concrete.zip (124 693 octree SBFEM elements, 616 413 free dofs), METIS {world}-way, Jacobi-PCG fixed {K} iterations"
        else:
            from pcg_mpi_solver_b200.hexmesh import hex_mdf_model
            subs = partition_mesh(hex_mdf_model(ng), world, assemble=False)
            config["workload"] = config["workload"].replace("trilinear hex elastostatics", f"trilinear hex elastostatics, METIS {world}-way element partition")
        sub = subs[rank]
        n_global = sub.n_global_eff
        op = sub.to_operator(comm, device=dev, kind=args.operator)
        A = op.A
        b = torch.from_numpy(sub.b).to(dev)
        del subs
    minv = op.jacobi()
    n = A.shape[0]
    is_csr = args.operator == "csr"
    log(f"operator ready ({args.operator}): n={n} halo={op.halo_bytes()} B")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up (also builds the CUDA graph of the iteration batch)
    op.solve(b, minv, 0.0, W, fixed_iters=True, check_every=min(W, CE))
    op.solve(b, minv, 0.0, K, fixed_iters=True, check_every=CE)
    barrier()
    log("warm-up done")

    # ---- timed region: exactly K iterations, device-timed loop
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    x, info = op.solve(b, minv, 0.0, K, fixed_iters=True, check_every=CE)
    barrier()
    loop_ms = max_over_ranks(info.loop_ms)
    log(f"timed loop: {loop_ms / K:.4f} ms/iter")
    assert info.loop_iters == K, (info.loop_iters, K)

    # ---- e2e: public API, host buffers (pinned b in, x out), everything inside the timed region
    # pinned host buffers exist before the timed region (a real caller reuses them across time steps)
    b_pin = b.cpu().pin_memory()
    x_host = torch.empty(n, dtype=torch.float64).pin_memory()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    b_dev = b_pin.to(dev, non_blocking=True)                                                   # H2D: this solve's right-hand side
    x_e2e, info_e = op.solve(b_dev, minv, 0.0, K, fixed_iters=True, check_every=CE)            # public operator API
    x_host.copy_(x_e2e, non_blocking=True)                                                     # D2H: the solution
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    log(f"e2e: {e2e_ms:.1f} ms")
    barrier()

    # ---- roofline pass: same K iterations with an event pair around every SpMV launch
    _, info_k = op.solve(b, minv, 0.0, K, fixed_iters=True, check_every=CE, time_kernels=True)
    spmv_ms = info_k.spmv_ms / max(info_k.spmv_timed, 1)
    spmv_ms = max_over_ranks(spmv_ms)
    spmv_share = info_k.spmv_ms / info_k.loop_ms if info_k.loop_ms > 0 else None
    peak, peak_src = measured_peaks()
    bytes_spmv = A.spmv_bytes()
    achieved = bytes_spmv / (spmv_ms * 1e-3) / 1e9
    iter_bytes = bytes_spmv + 96 * n
    barrier()
    log(f"roofline pass: spmv {spmv_ms:.4f} ms")

    full_solve = None
    if args.workload == "concrete":   # config C4: full solve to tol 1e-8 (the reference's own run: 1085 iterations at 1e-7, 12.6 s on 8 cores)
        xs, fi = op.solve(b, minv, 1e-8, 10000, check_every=16)
        full_solve = {"tol": 1e-8, "flag": fi.flag, "iterations": fi.iters, "relres": fi.relres, "loop_ms": max_over_ranks(fi.loop_ms),
                      "iterations_per_s": fi.iters / (max_over_ranks(fi.loop_ms) * 1e-3)}
        barrier()

    if rank == 0:
        value = K / (loop_ms * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": loop_ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(config, operator=args.operator, n_per_gpu=n, nnz_per_gpu=A.nnz if is_csr else A.nnz_equivalent, n_global=n_global,
                               plan=A.plan_info() if is_csr else {"kernel": "k_ebe_t24", "pattern_groups": 1},
                               halo_bytes_per_exchange=op.halo_bytes()),
                "dof_iterations_per_s": value * n_global,
                "e2e": {"value": K / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * n / K, "d2h_bytes_per_step": 8 * n / K,
                        "note": "one solve() of K iterations: pinned-host b -> device, K iterations + 2 residual matvecs + host polling, x -> pinned host"},
                "gpu_launches": int(info.launches),
                "clocks": clocks,
                "roofline": {"kernel": ({0: "k_spmv_merge", 1: "k_spmv_staged", 2: "k_spmv_persist"}[A.plan_info()["staged"]] + " (merge-path CSR SpMV, fp64)") if is_csr else "k_ebe_t24 (matrix-free EBE operator, fp64; bytes = its own 108 B/element, not the CSR figure)", "bound": "hbm", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic() if is_csr else None, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": bytes_spmv, "streamed_bytes_per_launch": A.stream_bytes() if is_csr else bytes_spmv,
                             "streamed_GBps": (A.stream_bytes() if is_csr else bytes_spmv) / (spmv_ms * 1e-3) / 1e9, "mean_launch_ms": spmv_ms, "launches_timed": int(info_k.spmv_timed),
                             "spmv_share_of_step": spmv_share,
                             "iteration": {"algorithmic_bytes": iter_bytes, "achieved_GBps": iter_bytes / (loop_ms / K * 1e-3) / 1e9,
                                           "frac": iter_bytes / (loop_ms / K * 1e-3) / 1e9 / peak}},
                "solve_check": {"flag": info.flag, "relres_after_K": info.relres}, "full_solve": full_solve}
        if world == 1 and not args.no_cpu and args.workload == "hex":
            try:
                line["cpu_baseline"] = cpu_reference(ng, max(1, args.cpu_iters))
            except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
