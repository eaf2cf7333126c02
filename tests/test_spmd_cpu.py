"""The SPMD (one process per part, shared-memory 'MPI') mode of the oracle used by bench.py's CPU reference arm:
same answer as the lock-step emulation and as the UNMODIFIED reference (golden hex run, 4 box parts)."""
import json
import os

import numpy as np

from oracle import ref_pcg as R
from oracle.hex_parts import hex_box_part, hex_box_part_spmd, link_parts
from oracle.spmd import run_spmd

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rank_fn(rank, size, comm, ng, tol, maxiter):
    from pcg_mpi_solver_b200.hexmesh import block_grid, partition_blocks
    blocks = partition_blocks(ng, block_grid(size))
    part = R.EbePart(hex_box_part_spmd(blocks, rank, h=1.0 / ng[0]))
    comm.setup_halo(part, comm._halo_box)
    R.update_bc([part], comm=comm)
    minv = R.Operator([part], comm).jacobi()
    out = R.ref_pcg([part], minv, tol, maxiter, nglob=3 * ng[0] * (ng[1] + 1) * (ng[2] + 1), comm=comm)
    return {"Flag": out["Flag"], "Iter": out["Iter"], "RelRes": out["RelRes"], "x": out["X"][0], "gdof": part.mp["DofVector"][part.eff]}


def test_spmd_oracle_matches_reference_golden():
    with open(os.path.join(GOLD, "hex_ref.json")) as f:
        meta = json.load(f)
    gold = np.load(os.path.join(GOLD, "hex_ref.npz"))
    ng = tuple(meta["ng"])
    res = run_spmd(4, _rank_fn, (ng, meta["tol"], meta["maxiter"]))
    run = meta["runs"]["box4"]
    assert all(r["Flag"] == 0 and r["Iter"] == run["Iter"] for r in res)
    u = np.zeros(gold["U_box4"].size)
    for r in res:
        u[r["gdof"]] = r["x"]
    assert np.linalg.norm(u - gold["U_box4"]) <= 1e-12 * np.linalg.norm(gold["U_box4"])


def test_spmd_part_tables_equal_link_parts():
    from pcg_mpi_solver_b200.hexmesh import block_grid, partition_blocks
    ng = (6, 5, 4)
    blocks = partition_blocks(ng, block_grid(8))
    linked = link_parts([hex_box_part(b.ng, b.e0, b.ne, i) for i, b in enumerate(blocks)])
    for r in range(8):
        p = hex_box_part_spmd(blocks, r, h=1.0)
        assert p["NbrMPIdVector"] == linked[r]["NbrMPIdVector"]
        assert all(np.array_equal(a, b) for a, b in zip(p["OvrlpLocalDofVecList"], linked[r]["OvrlpLocalDofVecList"]))
        assert np.array_equal(p["DofWeightVector"], linked[r]["DofWeightVector"])
