"""Reader of the reference's model-definition files (MDF) - the on-disk input of the pipeline.

Reference: src/data/read_input_model.py:38 unpacks `<model>.zip` into `<scratch>/ModelData/MDF/`; the
files are consumed at run_metis.py:21-38,70-71 and partition_mesh.py:172-175, 223-225, 324-330, 503-547.
Schema: raw little-endian `.bin` arrays - 2-D ones stored in FORTRAN order, the
`*Offset` arrays hold INCLUSIVE [start, end] ranges into the matching `*Flat` array, all ids 0-based -
plus a few MATLAB `.mat` files (GlobN, dt, Ke, Me, MatProp).
"""
from __future__ import annotations

import io
import os
import zipfile
from dataclasses import dataclass, field

import numpy as np
import scipy.io


@dataclass
class MdfModel:
    name: str
    n_elem: int
    n_dof: int
    n_dof_eff: int
    node_flat: np.ndarray       # NodeGlbFlat  int32
    node_offset: np.ndarray     # NodeGlbOffset (NE,2) int64, inclusive
    dof_flat: np.ndarray        # DofGlbFlat   int32
    dof_offset: np.ndarray      # DofGlbOffset (NE,2)
    sign_flat: np.ndarray       # SignFlat     bool
    sign_offset: np.ndarray     # SignOffset   (NE,2)
    etype: np.ndarray           # Type         int32 (pattern id)
    ck: np.ndarray              # Ck           float64 (= E*h)
    F: np.ndarray               # nodal loads  float64 [n_dof]
    Ud: np.ndarray              # prescribed displacements [n_dof]
    dof_eff: np.ndarray         # DofEff int32 sorted
    fixed_dof: np.ndarray       # FixedDof int32
    ke: list                    # pattern stiffness matrices (Ke.mat 'Data')
    dt: float = 0.0
    extra: dict = field(default_factory=dict)

    @property
    def n_node(self) -> int:
        return self.n_dof // 3


class _Source:
    def __init__(self, path):
        self.zip = zipfile.ZipFile(path) if os.path.isfile(path) else None
        self.dir = None if self.zip else path

    def has(self, name):
        return name in self.zip.namelist() if self.zip else os.path.exists(os.path.join(self.dir, name))

    def raw(self, name) -> bytes:
        if self.zip:
            return self.zip.read(name)
        with open(os.path.join(self.dir, name), "rb") as f:
            return f.read()

    def bin(self, name, dtype, shape=None):
        a = np.frombuffer(self.raw(name + ".bin"), dtype=dtype)
        if shape is not None and len(shape) == 2:
            a = a.reshape(shape, order="F")  # file_operations.py:332-333
        return a

    def mat(self, name):
        return scipy.io.loadmat(io.BytesIO(self.raw(name + ".mat")))


def load_mdf(path: str, name: str | None = None) -> MdfModel:
    """Load a model from `<model>.zip` or from an unpacked MDF directory."""
    src = _Source(path)
    glob_n = src.mat("GlobN")["Data"][0]  # run_metis.py:21-34
    ne, ndof, n_dof_flat, n_node_flat, ndof_eff = (int(glob_n[k]) for k in range(5))
    n_fixed = int(glob_n[8])
    node_flat = src.bin("NodeGlbFlat", np.int32)
    dof_flat = src.bin("DofGlbFlat", np.int32)
    assert node_flat.size == n_node_flat and dof_flat.size == n_dof_flat
    ke = [np.array(k, dtype=float) for k in src.mat("Ke")["Data"][0]]  # partition_mesh.py:546
    dt = float(src.mat("dt")["Data"][0][0]) if src.has("dt.mat") else 0.0
    m = MdfModel(
        name=name or os.path.splitext(os.path.basename(path.rstrip("/")))[0], n_elem=ne, n_dof=ndof, n_dof_eff=ndof_eff,
        node_flat=node_flat, node_offset=src.bin("NodeGlbOffset", np.int64, (ne, 2)),
        dof_flat=dof_flat, dof_offset=src.bin("DofGlbOffset", np.int64, (ne, 2)),
        sign_flat=src.bin("SignFlat", np.int8).astype(bool), sign_offset=src.bin("SignOffset", np.int64, (ne, 2)),
        etype=src.bin("Type", np.int32), ck=src.bin("Ck", np.float64),
        F=src.bin("F", np.float64), Ud=src.bin("Ud", np.float64),
        dof_eff=src.bin("DofEff", np.int32).astype(np.int64), fixed_dof=src.bin("FixedDof", np.int32).astype(np.int64),
        ke=ke, dt=dt)
    assert m.dof_eff.size == ndof_eff and m.fixed_dof.size == n_fixed and m.F.size == ndof
    return m
