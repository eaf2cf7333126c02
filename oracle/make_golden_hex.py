#!/usr/bin/env python
"""Generate tests/golden/hex_ref_*.npz by running the UNMODIFIED reference on a structured hex model
written in its own MDF format (oracle/hex_mdf.py), for 1 / 2 / 4 / 8 parts (box partitions and a METIS
partition).  Full solution vectors are stored (the model is small).  Build container only."""
import json
import os
import shutil
import sys
import zipfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import run_reference as rr  # noqa: E402
from oracle.hex_mdf import write_hex_mdf  # noqa: E402
from pcg_mpi_solver_b200.hexmesh import block_grid, partition_blocks  # noqa: E402
from pcg_mpi_solver_b200.metis import run_metis  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NG = (10, 8, 6)
TOL, MAXITER = 1e-10, 5000
work = sys.argv[1] if len(sys.argv) > 1 else "/tmp/pcgb_ref_hex"
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
ne = info["ne"]
nx, ny, nz = NG
out = {"ng": NG, "tol": TOL, "maxiter": MAXITER, "runs": {}}
arrays = {}


def box_elepart(world):
    ep = np.zeros(ne, dtype=np.int64)
    for r, b in enumerate(partition_blocks(NG, block_grid(world))):
        ez, ey, ex = np.meshgrid(np.arange(b.e0[2], b.e0[2] + b.ne[2]), np.arange(b.e0[1], b.e0[1] + b.ne[1]),
                                 np.arange(b.e0[0], b.e0[0] + b.ne[0]), indexing="ij")
        ep[((ez * ny + ey) * nx + ex).ravel()] = r
    return ep


flat = np.fromfile(rr.mdf_path(work) + "NodeGlbFlat.bin", dtype=np.int32)
off = np.fromfile(rr.mdf_path(work) + "NodeGlbOffset.bin", dtype=np.int64).reshape((ne, 2), order="F")
cases = [("box1", 1, None), ("box2", 2, box_elepart(2)), ("box4", 4, box_elepart(4)), ("box8", 8, box_elepart(8)),
         ("metis4", 4, run_metis(flat, off, 4))]
run_id = 0
for name, nparts, ep in cases:
    run_id += 1
    rr.metis_stage(work, nparts, ep)
    rr.partition_stage(work, nparts)
    rr.write_settings(work, TOL, MAXITER)
    rr.solve_stage(work, nparts, run_id=run_id)
    res, u = rr.read_results(work, "hexmodel", nparts, run_id, info["ndof"])
    out["runs"][name] = {k: res[k] for k in ("Flag", "Iter", "RelRes")}
    out["runs"][name]["nparts"] = nparts
    arrays[f"U_{name}"] = u
    if ep is not None:
        arrays[f"elepart_{name}"] = ep.astype(np.int8)
    # interface tables of the reference builder, for the builder parity test
    tabs = []
    for p in range(nparts):
        mp_ = rr.load_mesh_part(work, nparts, p)
        tabs.append({"nbr": [int(v) for v in mp_["NbrMPIdVector"]], "n_ovrlp": [int(len(v)) for v in mp_["OvrlpLocalDofVecList"]],
                     "wsum": float(mp_["DofWeightVector"].sum()), "ndof": int(mp_["NDOF"])})
    out["runs"][name]["parts"] = tabs
    print(name, out["runs"][name]["Flag"], out["runs"][name]["Iter"], out["runs"][name]["RelRes"], flush=True)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hex_ref.npz"), F=info["F"], eff=info["eff"], **arrays)
with open(os.path.join(ROOT, "tests", "golden", "hex_ref.json"), "w") as f:
    json.dump(out, f, indent=1)
u1 = arrays["U_box1"]
for k in arrays:
    if k.startswith("U_"):
        print(k, np.linalg.norm(arrays[k] - u1) / np.linalg.norm(u1))
