#!/usr/bin/env python
"""Run the UNMODIFIED upstream reference (read-only at /root/reference) under the fake-MPI
shim and harvest golden vectors.  TEST INFRASTRUCTURE ONLY - never imported by the product.

Pipeline reproduced (examples/run_basic_script.bash:13-52 of the reference):
    read_input_model.py -> run_metis.py -> partition_mesh.py -> GlobSettings.zpkl -> pcg_solver.py

Because mpi4py / mgmetis / matplotlib are absent in this image:
  * `oracle/fake_mpi` supplies `mpi4py` and an empty `matplotlib` (PYTHONPATH shim);
  * for N > 1 parts the element partition (`MeshPart_N.npy`, run_metis.py:92) is written by
    the caller (METIS_PartMeshDual from the CUDA toolkit, see pcg_mpi_solver_b200/metis.py) -
    the reference's own mgmetis call cannot run here;
  * the builder runs on ONE shimmed rank (its MPGSize path handles all parts,
    partition_mesh.py:113-116); the solver runs with one OS process per part (fork +
    multiprocessing queues), each executing pcg_solver.py via runpy as `__main__`.

CLI:
    python oracle/run_reference.py ranks N <script.py> [args...]   # internal: N-rank launch
Library use: see oracle/make_golden.py.
"""
from __future__ import annotations

import os
import pickle
import runpy
import subprocess
import sys
import zlib

import numpy as np

REF = os.environ.get("PCGB_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "fake_mpi")


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([SHIM, REF, env.get("PYTHONPATH", "")])
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        env[k] = "1"  # pcg_solver.py:10-15
    return env


def _run(script, args, cwd, nranks=1, quiet=True):
    path = os.path.join(REF, script)
    if nranks == 1:
        cmd = [sys.executable, path] + [str(a) for a in args]
    else:
        cmd = [sys.executable, os.path.abspath(__file__), "ranks", str(nranks), path] + [str(a) for a in args]
    res = subprocess.run(cmd, cwd=cwd, env=_env(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0 or not quiet:
        sys.stdout.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError(f"reference stage {script} failed ({res.returncode})")
    return res.stdout


def ingest(workdir, model_name, model_zip):
    """read_input_model.py WorkDir ModelName ScratchPath Zip (read_input_model.py:17-20)."""
    os.makedirs(workdir, exist_ok=True)
    return _run("src/data/read_input_model.py", [workdir, model_name, os.path.join(workdir, "data"), model_zip], workdir)


def mdf_path(workdir):
    return os.path.join(workdir, "data", "ModelData", "MDF") + "/"


def metis_stage(workdir, nparts, elepart=None):
    """run_metis.py N.  N=1 runs the reference; N>1 needs `elepart` (see module docstring)."""
    if not os.path.exists(mdf_path(workdir) + "MeshData_Glob.zpkl"):
        _run("src/solver/run_metis.py", [1], workdir)  # also writes MeshData_Glob.zpkl (run_metis.py:64)
    if nparts > 1:
        assert elepart is not None
        np.save(mdf_path(workdir) + f"MeshPart_{nparts}.npy", np.asarray(elepart, dtype=np.int64))


def partition_stage(workdir, nparts):
    """partition_mesh.py N 0 on one shimmed rank -> data/ModelData/MPI/N_<id>.mpidat."""
    return _run("src/solver/partition_mesh.py", [nparts, 0], workdir)


def write_settings(workdir, tol, maxiter, time_step_delta=(0, 1)):
    """GlobSettings.zpkl exactly as examples/run_basic_script.bash:30-49."""
    settings = {"TimeHistoryParam": {"ExportFlag": True, "ExportFrmRate": 1, "ExportFrms": [], "PlotFlag": False,
                                     "TimeStepDelta": list(time_step_delta), "ExportVars": "U"},
                "SolverParam": {"Tol": tol, "MaxIter": maxiter}}
    os.makedirs(os.path.join(workdir, "__pycache__"), exist_ok=True)
    with open(os.path.join(workdir, "__pycache__", "GlobSettings.zpkl"), "wb") as f:
        f.write(zlib.compress(pickle.dumps(settings, pickle.HIGHEST_PROTOCOL)))


def solve_stage(workdir, nparts, run_id, speed_test=0):
    """pcg_solver.py RunId SpeedTestFlag with one process per part."""
    return _run("src/solver/pcg_solver.py", [run_id, speed_test], workdir, nranks=nparts, quiet=False)


def load_mesh_part(workdir, nparts, part_id):
    """Decode one reference fixture N_<id>.mpidat (pcg_solver.py:100-106)."""
    base = os.path.join(workdir, "data", "ModelData", "MPI", str(nparts))
    meta = np.load(base + "_metadat.npy", allow_pickle=True).item()
    raw = np.fromfile(f"{base}_{part_id}.mpidat", dtype=meta["DTypeData"][part_id], count=meta["NfData"][part_id])
    return pickle.loads(zlib.decompress(raw.tobytes()))


def read_results(workdir, model_name, nparts, run_id, glob_ndof, frame=1):
    """Flag/Iter/RelRes of the single load step + the global displacement vector
    (file_operations.py:517-531, export_vtk.py:157-159)."""
    res = os.path.join(workdir, "data", f"Results_Run{run_id}")
    td = np.load(os.path.join(res, "PlotData", f"{model_name}_MP{nparts}_TimeData.npz"), allow_pickle=True)["TimeData"].item()

    def rd(name, dtype):
        return np.fromfile(os.path.join(res, "ResVecData", name + ".mpidat"), dtype=dtype)

    meta = np.load(os.path.join(res, "ResVecData", "Dof_metadat.npy"), allow_pickle=True).item()
    dof = rd("Dof", meta["DTypeData"][0])
    u = np.zeros(glob_ndof)
    u[dof] = rd(f"U_{frame}", np.float64)
    return {"Flag": int(td["Flag"][1]), "Iter": int(td["Iter"][1]), "RelRes": float(td["RelRes"][1]),
            "CalcTime": float(td["Mean_CalcTime"]), "CommWaitTime": float(td["Mean_CommWaitTime"]),
            "TotalTime": float(td["TotalTime"])}, u


# ---------------------------------------------------------------------------------------
def _rank_main(rank, size, inboxes, barrier, script, argv):
    try:
        os.sched_setaffinity(0, {rank % os.cpu_count()})  # one rank per core, like --map-by
    except Exception:
        pass
    from mpi4py import MPI  # the shim (PYTHONPATH)
    MPI._attach(rank, size, inboxes, barrier)
    sys.argv = [script] + argv
    runpy.run_path(script, run_name="__main__")


def _launch_ranks(n, script, argv):
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    inboxes = [ctx.Queue() for _ in range(n)]
    barrier = ctx.Barrier(n)
    procs = [ctx.Process(target=_rank_main, args=(r, n, inboxes, barrier, script, argv)) for r in range(n)]
    for p in procs:
        p.start()
    rc = 0
    for p in procs:
        p.join()
        rc |= p.exitcode or 0
    sys.exit(rc)


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[1] == "ranks":
        _launch_ranks(int(sys.argv[2]), sys.argv[3], sys.argv[4:])
    else:
        print(__doc__)
