"""ORACLE helper (test infrastructure): write a structured hex elastostatic model in the reference's own
MDF format so that the UNMODIFIED reference pipeline can consume it and produce
golden vectors (oracle/make_golden_hex.py).  File list and dtypes follow the readers at
run_metis.py:21-38,70-71 and partition_mesh.py:172-175, 223-225, 324-330, 503-547."""
from __future__ import annotations

import os

import numpy as np
import scipy.io

from .ref_pcg import hex_ke


def write_hex_mdf(path, ng, E=1.0, nu=0.3, h=None, traction=1.0, pull=None):
    """Hex mesh of ng elements, clamped at global x index 0, -z traction on x = max.  Returns a dict of
    the arrays written (global numbering: node = (gz*(ny+1)+gy)*(nx+1)+gx, dof = 3*node+dir).
    pull=d: displacement-controlled variant - the x = max face is ALSO fixed, with prescribed displacement
    Ud = (d, 0, 0), and there is no traction (exercises Fext = F*delta - K (Ud*delta), pcg_solver.py:226-238)."""
    os.makedirs(path, exist_ok=True)
    nx, ny, nz = ng
    h = 1.0 / nx if h is None else h
    ne = nx * ny * nz
    nnode = (nx + 1) * (ny + 1) * (nz + 1)
    ndof = 3 * nnode
    ez, ey, ex = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ex, ey, ez = ex.ravel(), ey.ravel(), ez.ravel()
    nodes = np.empty((ne, 8), dtype=np.int64)
    for l in range(8):
        nodes[:, l] = ((ez + ((l >> 2) & 1)) * (ny + 1) + (ey + ((l >> 1) & 1))) * (nx + 1) + (ex + (l & 1))
    dofs = (3 * nodes[:, :, None] + np.arange(3)[None, None, :]).reshape(ne, 24)

    def wbin(name, arr, dtype, fortran=False):
        a = np.asarray(arr, dtype=dtype)
        (a.ravel(order="F") if fortran else a.ravel()).tofile(os.path.join(path, name + ".bin"))

    def offsets(per_elem):
        start = np.arange(ne, dtype=np.int64) * per_elem
        return np.stack([start, start + per_elem - 1], axis=1)  # INCLUSIVE ends

    wbin("NodeGlbFlat", nodes, np.int32)
    wbin("NodeGlbOffset", offsets(8), np.int64, fortran=True)
    wbin("DofGlbFlat", dofs, np.int32)
    wbin("DofGlbOffset", offsets(24), np.int64, fortran=True)
    wbin("SignFlat", np.zeros(ne * 24), np.int8)
    wbin("SignOffset", offsets(24), np.int64, fortran=True)
    wbin("Type", np.zeros(ne), np.int32)
    wbin("Level", np.full(ne, h), np.float64)
    wbin("Ck", np.full(ne, E * h), np.float64)
    wbin("Cm", np.ones(ne), np.float64)
    wbin("Ce", np.ones(ne), np.float64)
    wbin("PolyMat", np.zeros(ne), np.int32)
    wbin("sctrs", np.stack([(ex + 0.5) * h, (ey + 0.5) * h, (ez + 0.5) * h], axis=1), np.float64, fortran=True)
    wbin("StrsGlb", np.tile(np.arange(6), (ne, 1)), np.int8, fortran=True)
    wbin("StrsSign", np.zeros((ne, 6)), np.int8, fortran=True)
    gz, gy, gx = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    gx, gy, gz = gx.ravel(), gy.ravel(), gz.ravel()
    coords = np.stack([gx * h, gy * h, gz * h], axis=1).ravel()
    F = np.zeros(ndof)
    face = gx == nx
    cy = np.where((gy == 0) | (gy == ny), 0.5, 1.0)
    cz = np.where((gz == 0) | (gz == nz), 0.5, 1.0)
    F[3 * np.nonzero(face)[0] + 2] = (-traction * h * h * cy * cz)[face]
    Ud = np.zeros(ndof)
    fixed_nodes = np.nonzero(gx == 0)[0]
    if pull is not None:
        F[:] = 0.0
        pulled = np.nonzero(gx == nx)[0]
        Ud[3 * pulled] = pull
        fixed_nodes = np.concatenate([fixed_nodes, pulled])
    fixed = np.sort((3 * fixed_nodes[:, None] + np.arange(3)[None, :]).ravel())
    eff = np.setdiff1d(np.arange(ndof), fixed)
    wbin("DiagM", np.ones(ndof), np.float64)
    wbin("F", F, np.float64)
    wbin("Ud", Ud, np.float64)
    wbin("Vd", np.zeros(0), np.float64)
    wbin("NodeCoordVec", coords, np.float64)
    wbin("DofEff", eff, np.int32)
    wbin("FixedDof", fixed, np.int32)
    glob_n = np.array([[ne, ndof, ne * 24, ne * 8, eff.size, 0, 0, 0, fixed.size]], dtype=float)
    scipy.io.savemat(os.path.join(path, "GlobN.mat"), {"Data": glob_n})
    scipy.io.savemat(os.path.join(path, "dt.mat"), {"Data": np.array([[0.0]])})
    cell = np.empty((1, 1), dtype=object)
    cell[0, 0] = hex_ke(1.0, nu)
    scipy.io.savemat(os.path.join(path, "Ke.mat"), {"Data": cell})
    mcell = np.empty((1, 1), dtype=object)
    mcell[0, 0] = np.eye(24)
    scipy.io.savemat(os.path.join(path, "Me.mat"), {"Data": mcell})
    mat = np.zeros((1,), dtype=[("E", "O"), ("Pos", "O"), ("Rho", "O")])
    mat[0] = (np.array([[E]]), np.array([[nu]]), np.array([[1.0]]))
    scipy.io.savemat(os.path.join(path, "MatProp.mat"), {"Data": mat})
    return {"F": F, "Ud": Ud, "eff": eff, "fixed": fixed, "nodes": nodes, "ndof": ndof, "ne": ne, "h": h}
