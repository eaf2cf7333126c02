"""METIS front-end: the `run_metis.py` stage of the reference, as a function.

Reference: src/solver/run_metis.py:70-92 builds the per-element node lists from
`NodeGlbFlat.bin` / `NodeGlbOffset.bin` (offsets are INCLUSIVE [start, end]) and calls
`mgmetis.metis.part_mesh_dual(nparts, cells, vwgt=ones)`; the result is the element ->
part vector stored as `MeshPart_<N>.npy` (0-based ids, int64).  For one part it
short-circuits to zeros (run_metis.py:84-85).

mgmetis (PyPI, version unpinned in the reference, README.md:23,36) is a thin binding of
METIS_PartMeshDual.  It is not installed here; the same METIS routine ships inside the
CUDA toolkit as `libmetis_static.a` (64-bit idx_t, 32-bit real_t), which we wrap in a
shared object built at `build()` time (csrc/Makefile target `metis`) and call through
ctypes.  Uniform element weights are METIS' default, so `vwgt` is passed as NULL.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "csrc", "libpcgb200_metis.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "csrc"), "metis"])
        _lib = ctypes.CDLL(_LIB_PATH)
        i64p = ctypes.POINTER(ctypes.c_int64)
        _lib.METIS_PartMeshDual.restype = ctypes.c_int
        _lib.METIS_PartMeshDual.argtypes = [i64p, i64p, i64p, i64p, ctypes.c_void_p, ctypes.c_void_p,
                                            i64p, i64p, ctypes.c_void_p, ctypes.c_void_p,
                                            i64p, i64p, i64p]
    return _lib


def part_mesh_dual(nparts: int, eptr: np.ndarray, eind: np.ndarray, ncommon: int = 1):
    """Partition a mesh through its dual graph (elements are graph vertices).

    eptr/eind: CSR-style element -> node lists (eptr has ne+1 entries, EXCLUSIVE ends).
    Returns (objval, epart[ne], npart[nn]) like mgmetis.metis.part_mesh_dual.
    """
    eptr = np.ascontiguousarray(eptr, dtype=np.int64)
    eind = np.ascontiguousarray(eind, dtype=np.int64)
    ne = eptr.size - 1
    nn = int(eind.max()) + 1 if eind.size else 0
    if nparts == 1:  # run_metis.py:84-85
        return 0, np.zeros(ne, dtype=np.int64), np.zeros(nn, dtype=np.int64)
    lib = _load()
    c = lambda v: ctypes.byref(ctypes.c_int64(v))
    objval = ctypes.c_int64(0)
    epart = np.zeros(ne, dtype=np.int64)
    npart = np.zeros(nn, dtype=np.int64)
    i64p = ctypes.POINTER(ctypes.c_int64)
    rc = lib.METIS_PartMeshDual(c(ne), c(nn), eptr.ctypes.data_as(i64p), eind.ctypes.data_as(i64p),
                                None, None, c(ncommon), c(nparts), None, None,
                                ctypes.byref(objval), epart.ctypes.data_as(i64p), npart.ctypes.data_as(i64p))
    if rc != 1:  # METIS_OK
        raise RuntimeError(f"METIS_PartMeshDual failed with status {rc}")
    return int(objval.value), epart, npart


def run_metis(node_glb_flat: np.ndarray, node_glb_offset: np.ndarray, nparts: int, ncommon: int = 1) -> np.ndarray:
    """`run_metis.py N` as a function: element -> part ids (int64, 0-based).

    node_glb_offset is the reference's (NE, 2) array of INCLUSIVE [start, end] ranges
    into node_glb_flat (run_metis.py:71,77).
    """
    off = np.asarray(node_glb_offset, dtype=np.int64)
    ne = off.shape[0]
    if nparts == 1:
        return np.zeros(ne, dtype=np.int64)
    starts, ends = off[:, 0], off[:, 1] + 1
    if np.array_equal(starts[1:], ends[:-1]) and starts[0] == 0:
        eptr = np.concatenate([starts, ends[-1:]])
        eind = np.asarray(node_glb_flat, dtype=np.int64)[: ends[-1]]
    else:  # general (non-contiguous) ranges
        lens = ends - starts
        eptr = np.concatenate([[0], np.cumsum(lens)])
        idx = np.repeat(starts - eptr[:-1], lens) + np.arange(eptr[-1])
        eind = np.asarray(node_glb_flat, dtype=np.int64)[idx]
    return part_mesh_dual(nparts, eptr, eind, ncommon)[1]
