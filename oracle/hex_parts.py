"""ORACLE helper (test infrastructure): structured hex boxes in the reference's own per-part data
layout (the dict keys calcMatVecProd / PCG read, partition_mesh.py:1310-1317), so that the restated
element-by-element operator (oracle/ref_pcg.EbePart <- pcg_solver.py:242-336) can run the benchmark
mesh exactly the way the reference would: one pattern type group, Ke (24x24) of the hexahedron,
ElemList_Ck = E*h, no sign flips.  Used by the CPU baseline of bench.py and by the CPU tests."""
from __future__ import annotations

import numpy as np

from .ref_pcg import hex_ke


def hex_box_part(ng, e0, ne, part_id=0, h=1.0, E=1.0, nu=0.3, traction=1.0):
    """RefMeshPart-like dict of one box (all local dofs, fixed ones included, like the reference)."""
    nx, ny, nz = ne
    nnx, nny, nnz_ = nx + 1, ny + 1, nz + 1
    ez, ey, ex = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ex, ey, ez = ex.ravel(), ey.ravel(), ez.ravel()
    dofs = np.empty((24, ex.size), dtype=np.int64)
    for l in range(8):
        node = ((ez + ((l >> 2) & 1)) * nny + (ey + ((l >> 1) & 1))) * nnx + (ex + (l & 1))
        for d in range(3):
            dofs[3 * l + d] = 3 * node + d
    ke = hex_ke(1.0, nu)
    group = {"ElemTypeId": 0, "ElemList_LocDofVector": dofs, "ElemList_LocDofVector_Flat": dofs.flatten(),
             "ElemList_SignVector": np.zeros(dofs.shape, dtype=bool), "ElemList_Ck": np.full(ex.size, E * h),
             "ElemStiffMat": ke, "ElemDiagStiffMat": np.diag(ke).copy(), "ElemList_LocNodeIdVector": None}
    nnode = nnx * nny * nnz_
    ndof = 3 * nnode
    lz, ly, lx = np.meshgrid(np.arange(nnz_), np.arange(nny), np.arange(nnx), indexing="ij")
    gx, gy, gz = lx.ravel() + e0[0], ly.ravel() + e0[1], lz.ravel() + e0[2]
    free_node = gx >= 1
    eff = (3 * np.nonzero(free_node)[0][:, None] + np.arange(3)[None, :]).ravel()
    F = np.zeros(ndof)
    on_face = gx == ng[0]
    cy = np.where((gy == 0) | (gy == ng[1]), 0.5, 1.0)
    cz = np.where((gz == 0) | (gz == ng[2]), 0.5, 1.0)
    F[3 * np.nonzero(on_face)[0] + 2] = (-traction * h * h * cy * cz)[on_face]
    gnode = (gz * (ng[1] + 1) + gy) * (ng[0] + 1) + gx
    return {"Id": part_id, "NDOF": ndof, "NNode": nnode, "LocDofEff": eff, "Ud": np.zeros(ndof), "RefLoadVector": F,
            "SubDomainData": {"StrucDataList": [group]}, "Flat_ElemLocDof": dofs.flatten(), "NCountDof": dofs.size,
            "NbrMPIdVector": [], "OvrlpLocalDofVecList": [], "DofWeightVector": np.ones(ndof),
            "NodeIdVector": gnode, "DofVector": (3 * gnode[:, None] + np.arange(3)[None, :]).ravel()}


def link_parts(parts):
    """Neighbour lists, overlap dof lists (ascending global node id, 3 dofs per node) and ownership
    weights exactly as config_Neighbours does (partition_mesh.py:805-887)."""
    for p in parts:
        p["NbrMPIdVector"], p["OvrlpLocalDofVecList"] = [], []
        p["DofWeightVector"] = np.ones(p["NDOF"])
    for p in parts:
        for q in parts:
            if p is q:
                continue
            common = np.intersect1d(p["NodeIdVector"], q["NodeIdVector"], assume_unique=True)  # :822
            if common.size == 0:
                continue
            order = np.argsort(p["NodeIdVector"])
            loc = order[np.searchsorted(p["NodeIdVector"][order], common)]                    # getIndices
            dofs = (3 * loc + np.array([[0], [1], [2]])).T.ravel()                             # :826
            p["NbrMPIdVector"].append(q["Id"])
            p["OvrlpLocalDofVecList"].append(dofs)
            if p["Id"] > q["Id"]:                                                              # :885-887
                p["DofWeightVector"][dofs] = 0
    return parts


def hex_box_part_spmd(blocks, rank, h, traction=1.0):
    """One rank's part INCLUDING its neighbour tables, built without the other parts in memory: the neighbours' node
    ids follow from their boxes (same rule as link_parts / partition_mesh.py:805-887)."""
    b = blocks[rank]
    p = hex_box_part(b.ng, b.e0, b.ne, rank, h=h, traction=traction)
    ng = b.ng
    for q_id, q in enumerate(blocks):
        if q_id == rank:
            continue
        lo = [max(b.e0[a], q.e0[a]) for a in range(3)]
        hi = [min(b.e0[a] + b.ne[a], q.e0[a] + q.ne[a]) for a in range(3)]
        if any(lo[a] > hi[a] for a in range(3)):
            continue
        gz, gy, gx = np.meshgrid(np.arange(lo[2], hi[2] + 1), np.arange(lo[1], hi[1] + 1), np.arange(lo[0], hi[0] + 1), indexing="ij")
        common = ((gz * (ng[1] + 1) + gy) * (ng[0] + 1) + gx).ravel()          # ascending global node id
        loc = np.searchsorted(p["NodeIdVector"], common)                         # NodeIdVector is ascending for a box
        dofs = (3 * loc + np.array([[0], [1], [2]])).T.ravel()
        p["NbrMPIdVector"].append(q_id)
        p["OvrlpLocalDofVecList"].append(dofs)
        if rank > q_id:
            p["DofWeightVector"][dofs] = 0
    return p
