#!/usr/bin/env python
"""Generate tests/golden/concrete_ref.json (+ sampled solution) by running the UNMODIFIED
reference on data/concrete.zip under the fake-MPI shim.  Test infrastructure; run in the
build container only (needs /root/reference).  ~3 min on 8 vCPUs.

    python oracle/make_golden_concrete.py [workdir]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import run_reference as rr  # noqa: E402
from pcg_mpi_solver_b200.metis import run_metis  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
work = sys.argv[1] if len(sys.argv) > 1 else "/tmp/pcgb_ref_concrete"
zip_path = os.path.join(rr.REF, "data", "concrete.zip")

rr.ingest(work, "concrete", zip_path)
mdf = rr.mdf_path(work)
rr.metis_stage(work, 1)
import pickle, zlib
glob = pickle.loads(zlib.decompress(open(mdf + "MeshData_Glob.zpkl", "rb").read()))
ne, ndof = glob["GlobNElem"], glob["GlobNDof"]
flat = np.fromfile(mdf + "NodeGlbFlat.bin", dtype=np.int32)
off = np.fromfile(mdf + "NodeGlbOffset.bin", dtype=np.int64).reshape((ne, 2), order="F")

out = {"model": "concrete", "GlobNElem": ne, "GlobNDof": ndof, "GlobNDofEff": glob["GlobNDofEff"],
       "Tol": 1e-7, "MaxIter": 10000, "runs": {}}
sample_idx = np.arange(0, ndof, 97)
samples = {}
for nparts in (1, 8):
    t0 = time.time()
    elepart = run_metis(flat, off, nparts) if nparts > 1 else None
    rr.metis_stage(work, nparts, elepart)
    rr.partition_stage(work, nparts)
    t_build = time.time() - t0
    rr.write_settings(work, 1e-7, 10000)
    rr.solve_stage(work, nparts, run_id=nparts)
    info, u = rr.read_results(work, "concrete", nparts, nparts, ndof)
    rng = np.random.default_rng(1234)
    probe = rng.standard_normal(ndof)
    info.update(norm2_U=float(np.linalg.norm(u)), max_abs_U=float(np.abs(u).max()),
                probe_dot=float(probe @ u), build_seconds=t_build)
    if nparts > 1:
        info["elements_per_part"] = np.bincount(elepart).tolist()
        nb = {}
        for p in range(nparts):
            mp_ = rr.load_mesh_part(work, nparts, p)
            nb[str(p)] = {"nbrs": [int(v) for v in mp_["NbrMPIdVector"]],
                          "shared_dofs": [int(len(v)) for v in mp_["OvrlpLocalDofVecList"]],
                          "NDOF": int(mp_["NDOF"]), "NDofEff": int(len(mp_["LocDofEff"])),
                          "weight_sum": float(mp_["DofWeightVector"].sum()),
                          "weight_sum_eff": float(mp_["DofWeightVector"][mp_["LocDofEff"]].sum())}
        info["parts"] = nb
        np.save(os.path.join(ROOT, "tests", "golden", f"concrete_elepart_{nparts}.npy"), elepart.astype(np.int8))
    out["runs"][str(nparts)] = info
    samples[f"U{nparts}"] = u[sample_idx]
    np.save(os.path.join(work, f"U_ref_{nparts}.npy"), u)
    print(nparts, info["Flag"], info["Iter"], info["RelRes"], info["norm2_U"], flush=True)

u1 = np.load(os.path.join(work, "U_ref_1.npy")); u8 = np.load(os.path.join(work, "U_ref_8.npy"))
out["rel_diff_U8_U1"] = float(np.linalg.norm(u8 - u1) / np.linalg.norm(u1))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "concrete_ref_samples.npz"), idx=sample_idx, **samples)
with open(os.path.join(ROOT, "tests", "golden", "concrete_ref.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "runs"}))
