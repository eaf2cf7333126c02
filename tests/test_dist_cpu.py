"""world_size-2 gloo tests (CPU) of the host-side multi-rank logic: box partition, interface lists,
ownership weights - the invariants the NCCL halo exchange relies on (partition_mesh.py:805-887)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _global_dofs(blk, local_dofs):
    """Global dof ids of free-dof indices of a box."""
    nxf = blk.ne[0] + 1 - blk.x_lo
    node, d = np.divmod(np.asarray(local_dofs), 3)
    lx = node % nxf + blk.x_lo
    ly = (node // nxf) % (blk.ne[1] + 1)
    lz = node // (nxf * (blk.ne[1] + 1))
    gx, gy, gz = lx + blk.e0[0], ly + blk.e0[1], lz + blk.e0[2]
    return 3 * ((gz * (blk.ng[1] + 1) + gy) * (blk.ng[0] + 1) + gx) + d


def _worker(rank, world, port, ng, pgrid, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pcg_mpi_solver_b200.hexmesh import interface_lists, partition_blocks
    blocks = partition_blocks(ng, pgrid)
    blk = blocks[rank]
    nbr, lists, w = interface_lists(blocks, rank)
    # 1) both sides of every interface list the same global dofs in the same order
    mine = {nb: _global_dofs(blk, l) for nb, l in zip(nbr, lists)}
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    ok = True
    for nb in nbr:
        ok &= rank in everyone[nb] and np.array_equal(everyone[nb][rank], mine[nb])
        ok &= bool(np.all(np.diff(mine[nb][::3]) > 0))  # ascending global node id
    # 2) ownership weights partition the free dofs: sum over ranks = number of global free dofs
    t = torch.tensor([float(w.sum())], dtype=torch.float64)
    dist.all_reduce(t)
    n_global = 3 * ng[0] * (ng[1] + 1) * (ng[2] + 1)
    ok &= t.item() == n_global
    # 3) a shared dof is owned (w = 1) by exactly the lowest rank holding it
    owned = {int(g) for g in _global_dofs(blk, np.nonzero(w == 1)[0])}
    all_owned = [None] * world
    dist.all_gather_object(all_owned, owned)
    if rank == 0:
        union = set().union(*all_owned)
        ok &= len(union) == sum(len(s) for s in all_owned) == n_global
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,ng", [(2, (6, 5, 4)), (4, (8, 6, 3))])
def test_interface_lists_gloo(world, ng):
    from pcg_mpi_solver_b200.hexmesh import block_grid
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, ng, block_grid(world), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_hex_type_group_is_the_same_operator_as_the_csr_box():
    """hexmesh.hex_type_group (input of the opt-in EBE operator) == the assembled box, through the CPU oracle."""
    sys.path.insert(0, ROOT)
    from oracle import ref_pcg as R
    from pcg_mpi_solver_b200.hexmesh import HexBlock, hex_type_group
    for e0 in ((0, 0, 0), (3, 1, 0)):
        blk = HexBlock((9, 6, 5), e0, (4, 3, 2), h=0.25)
        grp, eff, ndof = hex_type_group(blk)
        assert eff.size == blk.n
        mp_ = {"Id": 0, "NDOF": ndof, "LocDofEff": eff, "NbrMPIdVector": [], "OvrlpLocalDofVecList": [], "DofWeightVector": np.ones(ndof),
               "SubDomainData": {"StrucDataList": [{"ElemList_LocDofVector": grp.loc_dof, "ElemList_SignVector": grp.sign, "ElemList_Ck": grp.ck,
                                                    "ElemStiffMat": grp.ke, "ElemDiagStiffMat": np.diag(grp.ke).copy()}]},
               "Flat_ElemLocDof": grp.loc_dof.flatten(), "NCountDof": grp.loc_dof.size}
        part = R.EbePart(mp_)
        A = R.hex_box_csr(blk.ng, blk.e0, blk.ne, h=blk.h)
        x = np.random.default_rng(1).standard_normal(blk.n)
        y = R.Operator([part]).apply([x])[0]
        assert np.abs(y - A @ x).max() <= 1e-13 * np.abs(A @ x).max()


def test_hex_mdf_model_equals_the_mdf_files_the_reference_consumed(tmp_path):
    """hexmesh.hex_mdf_model (in-memory) == oracle/hex_mdf.py's files read back with load_mdf - the files the
    unmodified reference ran on for tests/golden/hex_ref.*; and METIS + builder + assembly run on it."""
    sys.path.insert(0, ROOT)
    from oracle import ref_pcg as R
    from oracle.hex_mdf import write_hex_mdf
    from pcg_mpi_solver_b200.hexmesh import hex_mdf_model
    from pcg_mpi_solver_b200.model import load_mdf
    from pcg_mpi_solver_b200.partition import partition_mesh
    ng = (10, 8, 6)
    write_hex_mdf(str(tmp_path), ng)
    a, b = load_mdf(str(tmp_path)), hex_mdf_model(ng)
    for f in ("node_flat", "node_offset", "dof_flat", "dof_offset", "sign_flat", "sign_offset", "etype", "ck", "F", "Ud", "dof_eff", "fixed_dof"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert (a.n_elem, a.n_dof, a.n_dof_eff) == (b.n_elem, b.n_dof, b.n_dof_eff)
    assert np.abs(a.ke[0] - b.ke[0]).max() <= 1e-15        # oracle's and product's Q1 matrices (independent codes)
    subs = partition_mesh(b, 4)                               # METIS 4-way, host assembly
    assert sum(s.weights.sum() for s in subs) == b.n_dof_eff
    parts = [R.CsrPart(s.A, s.b, s.nbr, s.ovrlp, s.weights, part_id=s.id) for s in subs]
    out = R.ref_pcg(parts, R.Operator(parts).jacobi(), 1e-10, 5000, nglob=b.n_dof_eff)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "hex_ref.npz"))
    u = np.zeros(b.n_dof)
    for s, x in zip(subs, out["X"]):
        u[s.dof_eff_global] = x
    assert out["Flag"] == 0 and np.linalg.norm(u - gold["U_box1"]) <= 1e-9 * np.linalg.norm(gold["U_box1"])
