#!/usr/bin/env python
"""Golden run of the UNMODIFIED reference on a displacement-controlled hex model (prescribed Ud != 0 on the x = max
face, no load) with a TWO-STEP ramp TimeStepDelta = [0, 0.5, 1.0]: pins updateBC (Fext = F*delta - K (Ud*delta),
pcg_solver.py:226-238), Un = X + Udi (:598) and the time-step shell (:1002-1008).  1 and 2 parts.  Build container only."""
import json
import os
import shutil
import sys
import zipfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import run_reference as rr  # noqa: E402
from oracle.hex_mdf import write_hex_mdf  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NG, PULL, TOL, MAXITER, DELTAS = (8, 5, 4), 0.01, 1e-10, 5000, (0, 0.5, 1.0)
work = sys.argv[1] if len(sys.argv) > 1 else "/tmp/pcgb_ref_hex_pull"
shutil.rmtree(work, ignore_errors=True)
os.makedirs(work)
src = os.path.join(work, "mdf_src")
info = write_hex_mdf(src, NG, pull=PULL)
zpath = os.path.join(work, "hexpull.zip")
with zipfile.ZipFile(zpath, "w") as z:
    for f in os.listdir(src):
        z.write(os.path.join(src, f), f)
rr.ingest(work, "hexpull", zpath)
rr.metis_stage(work, 1)
nx, ny, nz = NG
ez, ey, ex = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
ep2 = (ex.ravel() >= nx // 2).astype(np.int64)
out = {"ng": NG, "pull": PULL, "tol": TOL, "maxiter": MAXITER, "deltas": DELTAS, "runs": {}}
arrays = {"elepart_2": ep2.astype(np.int8)}
for run_id, (name, nparts, ep) in enumerate([("p1", 1, None), ("p2", 2, ep2)], start=1):
    rr.metis_stage(work, nparts, ep)
    rr.partition_stage(work, nparts)
    rr.write_settings(work, TOL, MAXITER, DELTAS)
    rr.solve_stage(work, nparts, run_id=run_id)
    res1, u1 = rr.read_results(work, "hexpull", nparts, run_id, info["ndof"], frame=1)
    res2, u2 = rr.read_results(work, "hexpull", nparts, run_id, info["ndof"], frame=2)
    import numpy as _np
    td = _np.load(os.path.join(work, "data", f"Results_Run{run_id}", "PlotData", f"hexpull_MP{nparts}_TimeData.npz"), allow_pickle=True)["TimeData"].item()
    out["runs"][name] = {"nparts": nparts, "Flag": [int(v) for v in td["Flag"]], "Iter": [int(v) for v in td["Iter"]],
                         "RelRes": [float(v) for v in td["RelRes"]]}
    arrays[f"U1_{name}"], arrays[f"U2_{name}"] = u1, u2
    print(name, out["runs"][name], np.linalg.norm(u2 - 2 * u1) / np.linalg.norm(u2), flush=True)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hex_pull_ref.npz"), Ud=info["Ud"], eff=info["eff"], **arrays)
with open(os.path.join(ROOT, "tests", "golden", "hex_pull_ref.json"), "w") as f:
    json.dump(out, f, indent=1)
