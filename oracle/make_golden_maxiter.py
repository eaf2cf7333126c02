#!/usr/bin/env python
"""Golden vectors for the NON-CONVERGED exit of PCG (pcg_solver.py:566-598) from the UNMODIFIED reference.

The reference binds `MP_XMin = MP_X` (pcg_solver.py:379-380) and updates MP_X in place (`MP_X += Alpha*MP_P`, :516), so
XMin FOLLOWS the current iterate until the first `MP_XMin = np.array(MP_X)` copy (:557).  On a MaxIter exit before
any improvement of the residual the exported solution is therefore the LATEST iterate, not the initial guess.  Two runs on
the structured hex model of make_golden_hex.py (its residual grows during the first iterations):
    maxiter3   3 iterations: no improvement recorded -> XMin is still aliased
    maxiter30  30 iterations: improvements recorded  -> XMin is a frozen copy
Writes tests/golden/hex_maxiter_ref.{json,npz}.  Build container only (needs /root/reference)."""
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
NG = (10, 8, 6)
TOL = 1e-10
work = sys.argv[1] if len(sys.argv) > 1 else "/tmp/pcgb_ref_maxiter"
shutil.rmtree(work, ignore_errors=True)
os.makedirs(work)
src = os.path.join(work, "mdf_src")
info = write_hex_mdf(src, NG)
zpath = os.path.join(work, "hexmodel.zip")
with zipfile.ZipFile(zpath, "w") as z:
    for f in os.listdir(src):
        z.write(os.path.join(src, f), f)
rr.ingest(work, "hexmodel", zpath)
rr.metis_stage(work, 1)
rr.partition_stage(work, 1)
out = {"ng": NG, "tol": TOL, "runs": {}}
arrays = {}
for run_id, maxiter in enumerate((3, 30), start=1):
    rr.write_settings(work, TOL, maxiter)
    rr.solve_stage(work, 1, run_id=run_id)
    res, u = rr.read_results(work, "hexmodel", 1, run_id, info["ndof"])
    out["runs"][f"maxiter{maxiter}"] = {"maxiter": maxiter, **{k: res[k] for k in ("Flag", "Iter", "RelRes")}}
    arrays[f"U_maxiter{maxiter}"] = u
    print(maxiter, res["Flag"], res["Iter"], res["RelRes"], np.linalg.norm(u), flush=True)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hex_maxiter_ref.npz"), F=info["F"], eff=info["eff"], **arrays)
with open(os.path.join(ROOT, "tests", "golden", "hex_maxiter_ref.json"), "w") as f:
    json.dump(out, f, indent=1)
