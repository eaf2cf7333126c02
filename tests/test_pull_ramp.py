"""Displacement-controlled hex model with a two-step ramp, against the UNMODIFIED reference
(tests/golden/hex_pull_ref.*, oracle/make_golden_hex_pull.py): pins updateBC with Ud != 0
(Fext = F*delta - K (Ud*delta), pcg_solver.py:226-238), Un = X + Udi (:598), the warm start of the next load step
(:358, :378) and the time-step shell (:1002-1008) - for the oracle and for the product's file-compatible stage."""
import json
import os

import numpy as np
import pytest

from oracle import ref_pcg as R
from oracle import run_reference as rr
from oracle.hex_mdf import write_hex_mdf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLD, "hex_pull_ref.json")) as f:
        meta = json.load(f)
    return meta, np.load(os.path.join(GOLD, "hex_pull_ref.npz"))


def _model(tmp_path, meta):
    from pcg_mpi_solver_b200.model import load_mdf
    mdf = os.path.join(str(tmp_path), "data", "ModelData", "MDF") + "/"
    info = write_hex_mdf(mdf, tuple(meta["ng"]), pull=meta["pull"])
    return load_mdf(mdf, "hexpull"), info


@pytest.mark.parametrize("case", ["p1", "p2"])
def test_oracle_reproduces_reference_ramp(gold, tmp_path, case):
    from pcg_mpi_solver_b200.partition import build_subdomains
    meta, arr = gold
    model, info = _model(tmp_path, meta)
    run = meta["runs"][case]
    ep = arr["elepart_2"].astype(np.int64) if run["nparts"] == 2 else np.zeros(model.n_elem, dtype=np.int64)
    subs = build_subdomains(model, ep, run["nparts"], assemble=False)
    parts = [R.EbePart(s.to_refmeshpart()) for s in subs]
    op = R.Operator(parts)
    minv = op.jacobi()
    un = [np.zeros(p.ndof) for p in parts]
    for step, delta in enumerate(meta["deltas"][1:], start=1):
        udis = R.update_bc(parts, delta)                              # updateBC (:226-238)
        for p, u in zip(parts, un):
            p.x0 = u[p.eff]                                           # warm start from the previous Un (:358, :378)
        out = R.ref_pcg(parts, minv, meta["tol"], meta["maxiter"], nglob=model.n_dof_eff)
        assert out["Flag"] == run["Flag"][step] == 0 and out["Iter"] == run["Iter"][step]
        ug = np.zeros(model.n_dof)
        for k, (p, s) in enumerate(zip(parts, subs)):
            full = np.zeros(p.ndof)
            full[p.eff] = out["X"][k]
            un[k] = full + udis[k]                                    # Un = X_unq + Udi (:598)
            ug[s.dof_vector] = un[k]
        ref = arr[f"U{step}_{case}"]
        assert np.linalg.norm(ug - ref) <= 1e-12 * np.linalg.norm(ref)


def _workdir(tmp_path, meta):
    from pcg_mpi_solver_b200.pcg_solver import exportz
    work = str(tmp_path)
    os.makedirs(os.path.join(work, "__pycache__"), exist_ok=True)
    os.makedirs(os.path.join(work, "data", "ModelData", "MPI"), exist_ok=True)
    exportz(os.path.join(work, "__pycache__", "ModelDataPaths.zpkl"),
            {"ScratchPath": os.path.join(work, "data"), "MDF_Path": os.path.join(work, "data", "ModelData", "MDF") + "/",
             "PyDataPath_Part": os.path.join(work, "data", "ModelData", "MPI") + "/", "ModelName": "hexpull"})
    rr.write_settings(work, meta["tol"], meta["maxiter"], tuple(meta["deltas"]))
    return work


def test_builder_rhs_and_cli_reproduce_reference_ramp_cpu(gold, tmp_path):
    """partition.build_subdomains' Fext for Ud != 0 and the file-compatible stage (CPU checker as the solve backend)."""
    from pcg_mpi_solver_b200.partition import partition_mesh
    from pcg_mpi_solver_b200.pcg_solver import export_mesh_parts, run
    from tests.test_cli import _oracle_backend
    meta, arr = gold
    model, info = _model(tmp_path, meta)
    sub = partition_mesh(model, 1, assemble=True)[0]
    # b = (F - K Ud)[Eff] computed by the product builder == the assembled full matrix applied to Ud
    part = R.EbePart(sub.to_refmeshpart())
    R.update_bc([part], 1.0)
    np.testing.assert_allclose(sub.b, part.b, rtol=1e-13, atol=1e-16)
    assert np.linalg.norm(sub.b) > 0
    work = _workdir(tmp_path, meta)
    export_mesh_parts(os.path.join(work, "data", "ModelData", "MPI") + "/", [sub])
    out = run(1, 0, workdir=work, backend=_oracle_backend, quiet=True)
    run1 = meta["runs"]["p1"]
    assert [int(v) for v in out["Iter"]] == run1["Iter"]
    for frame in (1, 2):
        _, u = rr.read_results(work, "hexpull", 1, 1, info["ndof"], frame=frame)
        ref = arr[f"U{frame}_p1"]
        assert np.linalg.norm(u - ref) <= 1e-12 * np.linalg.norm(ref)
        assert np.allclose(u[3 * np.nonzero(info["Ud"][0::3])[0]], meta["pull"] * meta["deltas"][frame])   # prescribed dofs carry Ud*delta


# ---------------------------------------------------------------------------------------------------------------
# world_size-2 gloo test of the MULTI-RANK host logic of the file-compatible stage: per-rank fixtures, the interface
# sum of K (Ud delta) across ranks, result files written by two ranks at gathered offsets (file_operations.py:348-375)
class _GlooComm:
    """Reductions / interface exchange of the CPU checker over torch.distributed (gloo); sums in rank order."""

    def __init__(self, dist):
        self.dist, self.rank, self.size = dist, dist.get_rank(), dist.get_world_size()

    def allreduce(self, v):
        box = [None] * self.size
        self.dist.all_gather_object(box, np.atleast_1d(np.asarray(v, dtype=float)))
        tot = box[0]
        for b in box[1:]:
            tot = tot + b
        return float(tot[0]) if np.ndim(v) == 0 else tot

    def exchange_add_full(self, part, y):
        box = [None] * self.size
        self.dist.all_gather_object(box, {nb: y[idx].copy() for nb, idx in zip(part.nbr, part.ovrlp_full)})
        for nb, idx in zip(part.nbr, part.ovrlp_full):
            y[idx] += box[nb][self.rank]
        return y


def _gloo_backend(mp, ranks):
    from pcg_mpi_solver_b200.partition import SubdomainData, TypeGroup
    comm = _GlooComm(ranks.dist)
    part = R.EbePart(mp)
    groups = [TypeGroup(int(g["ElemTypeId"]), g["ElemList_LocDofVector"], g["ElemList_SignVector"], g["ElemList_Ck"], g["ElemStiffMat"], None)
              for g in mp["SubDomainData"]["StrucDataList"]]
    sub = SubdomainData(int(mp["Id"]), ranks.size, np.asarray(mp["DofVector"]), np.asarray(mp["NodeIdVector"]), part.eff, groups,
                        part.nbr, part.ovrlp_full, [], np.asarray(mp["DofWeightVector"], dtype=float),
                        np.asarray(mp["RefLoadVector"], dtype=float), np.asarray(mp["Ud"], dtype=float),
                        int(mp["GlobData"]["GlobNDofEff"]), int(mp["GlobData"]["GlobNDof"]))
    minv = R.Operator([part], comm).jacobi()

    def solve_step(b, x0, tol, maxiter):
        part.b, part.x0 = b, x0
        out = R.ref_pcg([part], minv, tol, maxiter, nglob=sub.n_global_eff, comm=comm)
        return out["X"][0], out["Flag"], out["RelRes"], out["Iter"]

    return sub, solve_step


def _cli_rank(rank, world, port, work, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PCGB_DIST_BACKEND="gloo")
    from pcg_mpi_solver_b200.pcg_solver import run
    out = run(2, 0, workdir=work, backend=_gloo_backend, quiet=True)
    q.put((rank, [int(v) for v in out["Iter"]], [int(v) for v in out["Flag"]]))
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


def test_cli_two_ranks_gloo_reproduce_reference_two_part_run(gold, tmp_path):
    import torch.multiprocessing as tmp_mp
    from pcg_mpi_solver_b200.partition import build_subdomains
    from pcg_mpi_solver_b200.pcg_solver import export_mesh_parts
    meta, arr = gold
    model, info = _model(tmp_path, meta)
    subs = build_subdomains(model, arr["elepart_2"].astype(np.int64), 2, assemble=False)
    work = _workdir(tmp_path, meta)
    export_mesh_parts(os.path.join(work, "data", "ModelData", "MPI") + "/", subs)
    ctx = tmp_mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cli_rank, args=(r, 2, 29433, work, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    run2 = meta["runs"]["p2"]
    assert res[0][1] == run2["Iter"] and res[0][2] == [0, 0, 0]      # rank 0 keeps the per-step records (:593-596)
    for frame in (1, 2):
        _, u = rr.read_results(work, "hexpull", 2, 2, info["ndof"], frame=frame)   # one file written by both ranks
        ref = arr[f"U{frame}_p2"]
        assert np.linalg.norm(u - ref) <= 1e-11 * np.linalg.norm(ref)
