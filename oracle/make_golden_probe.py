#!/usr/bin/env python
"""Golden input/output vectors of the reference's OPERATOR on data/concrete.zip, produced by importing the
unmodified reference module (src/solver/pcg_solver.py) under the fake-MPI shim and calling its own
functions: calcMPFint (pcg_solver.py:339-342), updatePreconditioner (:346-352), updateBC (:226-238).
Needs the 1-part fixture written by oracle/make_golden_concrete.py.  Build container only."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import run_reference as rr  # noqa: E402

work = sys.argv[1] if len(sys.argv) > 1 else "/tmp/pcgb_ref_concrete"
sys.path.insert(0, rr.REF)
sys.path.insert(0, rr.SHIM)
os.chdir(work)
import src.solver.pcg_solver as ps  # noqa: E402  (the reference module, imported not copied)
from mpi4py import MPI  # noqa: E402  (the shim)

ps.Comm, ps.Rank, ps.N_Workers = MPI.COMM_WORLD, 0, 1
mp = rr.load_mesh_part(work, 1, 0)
glob = {"FintCalcMode": "outbin", "TimeStepCount": 1, "TimeStepDelta": [0, 1],
        "MP_TimeRecData": {"dT_FileRead": 0.0, "dT_Calc": 0.0, "dT_CommWait": 0.0, "t0": 0.0}}
glob.update(mp["GlobData"])
mp["GlobData"] = glob
ndof = mp["NDOF"]
dofv = np.asarray(mp["DofVector"])
v = np.sin(0.001 * dofv) + 0.25 * np.cos(0.37 * dofv)   # deterministic probe, function of the GLOBAL dof id
v[np.asarray(mp["LocFixedDof"])] = 0.0
y = ps.calcMPFint(v, mp)
ps.updatePreconditioner(mp)
ps.updateBC(mp)
minv = mp["InvDiagPreCondVector0"]
fext = mp["Fext"]
idx = np.arange(0, ndof, 101)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "concrete_probe.npz"), idx=idx, y=y[idx], minv_idx=np.arange(0, len(minv), 101),
                    minv=minv[::101], fext=fext[idx])
out = {"norm_v": float(np.linalg.norm(v)), "norm_y": float(np.linalg.norm(y)), "sum_y": float(y.sum()), "norm_minv": float(np.linalg.norm(minv)),
       "norm_fext": float(np.linalg.norm(fext)), "vTy": float(v @ y), "ndof": int(ndof), "neff": int(len(mp["LocDofEff"]))}
with open(os.path.join(ROOT, "tests", "golden", "concrete_probe.json"), "w") as f:
    json.dump(out, f, indent=1)
print(out)
