"""Opt-in matrix-free element-by-element operator - the reference's own operator form
(calcMatVecProd, pcg_solver.py:263-300) on the GPU instead of the assembled CSR matrix.

Status: parity-green on B200 (tests/test_gpu_ebe.py: hex vs CSR/oracle, concrete vs the reference's 1085 iterations, several
live operators with different pattern matrices); [B200] 128^3: 0.27 ms per application, 2589 PCG it/s (`bench.py --operator ebe`,
profiles/bench_r2p_n1_ebe.json) against 0.80 ms / 1116 it/s for the CSR node-block kernel.  Scatter-adds with fp64 atomics
(reproducible to rounding only); `EbeMatrixColored` is the bit-reproducible coloured variant (3x slower).  `to_operator(kind="ebe")`.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib


class EbeMatrix:
    """Pattern groups of one subdomain on the device, indices in the free-dof numbering (-1 = clamped)."""

    def __init__(self, groups, loc_dof_eff: np.ndarray, ndof: int, device="cuda"):
        self.device = torch.device(device)
        n = int(len(loc_dof_eff))
        self.shape = (n, n)
        eff_map = np.full(ndof, -1, dtype=np.int32)
        eff_map[loc_dof_eff] = np.arange(n, dtype=np.int32)
        self._keep = []          # device tensors and host arrays the C side points into
        cgroups = (_lib.EbeGroup * len(groups))()
        diag = np.zeros(n)
        self.nnz_equivalent = 0
        for k, g in enumerate(groups):
            nd, ne = g.loc_dof.shape
            idx = eff_map[g.loc_dof].astype(np.int32)                       # (nd, ne)
            ke = np.ascontiguousarray(g.ke, dtype=np.float64)
            if not np.allclose(ke, ke.T, rtol=1e-12, atol=1e-12 * np.abs(ke).max()):
                raise ValueError("EbeMatrix: pattern matrices must be symmetric")
            d_idx = torch.from_numpy(np.ascontiguousarray(idx)).to(self.device)
            d_ck = torch.from_numpy(np.ascontiguousarray(g.ck, dtype=np.float64)).to(self.device)
            d_sign = torch.from_numpy(np.ascontiguousarray(g.sign.astype(np.uint8))).to(self.device) if g.sign.any() else None
            self._keep += [d_idx, d_ck, d_sign, ke]
            cg = cgroups[k]
            cg.nd, cg.ne = nd, ne
            cg.d_idx, cg.d_ck = d_idx.data_ptr(), d_ck.data_ptr()
            cg.d_sign = d_sign.data_ptr() if d_sign is not None else None
            cg.ke_host = ke.ctypes.data
            # diagonal for the Jacobi preconditioner: sum_e Ck * diag(Ke)  (pcg_solver.py:282-287); signs cancel (s_i^2 = 1)
            contrib = (g.ck[None, :] * np.diag(ke)[:, None]).ravel()
            flat = idx.ravel()
            ok = flat >= 0
            diag += np.bincount(flat[ok], weights=contrib[ok], minlength=n)
            self.nnz_equivalent += int(nd) * int(nd) * int(ne)
        self._diag = diag
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().pcgb_ebe_create(n, len(groups), cgroups, ctypes.byref(self._h)), "pcgb_ebe_create")

    @property
    def handle(self):
        return self._h

    def apply_local(self, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """y = K_i x, no interface sum."""
        y = out if out is not None else torch.empty(self.shape[0], dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().pcgb_ebe_apply(self._h, _lib.ptr(x), _lib.ptr(y), _lib.stream_ptr()), "pcgb_ebe_apply")
        return y

    spmv = apply_local

    def diagonal(self) -> torch.Tensor:
        return torch.from_numpy(self._diag).to(self.device)

    def spmv_bytes(self) -> int:
        return int(_lib.load().pcgb_ebe_bytes(self._h))

    def __del__(self):
        try:
            if self._h:
                _lib.load().pcgb_ebe_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass


class EbeMatrixColored:
    """Operator-level variant (green on B200, tests/test_gpu_ebe_colored.py): the same operator with an atomics-free,
    bit-reproducible scatter.  All elements of the subdomain are coloured together (coloring.color_elements: no two
    elements of a colour share a node), every (pattern group, colour) slice becomes its own group, slices are passed
    in colour order and the kernels of one colour use plain `y[dof] += v` (csrc/ebe_color.cuh)."""

    def __init__(self, groups, loc_dof_eff: np.ndarray, ndof: int, device="cuda", colors_per_group=None):
        from .coloring import color_elements
        self.device = torch.device(device)
        n = int(len(loc_dof_eff))
        self.shape = (n, n)
        eff_map = np.full(ndof, -1, dtype=np.int32)
        eff_map[loc_dof_eff] = np.arange(n, dtype=np.int32)
        if colors_per_group is None:
            # element -> node lists over ALL groups (dofs come in triples per node: node = dof // 3)
            flat = np.concatenate([(g.loc_dof[0::3, :] // 3).T.ravel() for g in groups])
            counts = np.concatenate([np.full(g.loc_dof.shape[1], g.loc_dof.shape[0] // 3, dtype=np.int64) for g in groups])
            ptr = np.concatenate([[0], np.cumsum(counts)])
            colors, ncolors = color_elements(flat, ptr, ndof // 3 + 1)
            split = np.cumsum([g.loc_dof.shape[1] for g in groups])[:-1]
            colors_per_group = np.split(colors, split)
        else:
            ncolors = int(max(int(c.max()) for c in colors_per_group if c.size) + 1)
        self.ncolors = ncolors
        self._keep = []
        entries = []     # (colour, nd, ne, d_idx, d_sign, d_ck, ke)
        for g, col in zip(groups, colors_per_group):
            ke = np.ascontiguousarray(g.ke, dtype=np.float64)   # ONE array per pattern group: its address identifies the pattern
            self._keep.append(ke)
            idx_all = eff_map[g.loc_dof].astype(np.int32)
            has_sign = bool(g.sign.any())
            for c in range(ncolors):
                sel = np.nonzero(col == c)[0]
                if sel.size == 0:
                    continue
                d_idx = torch.from_numpy(np.ascontiguousarray(idx_all[:, sel])).to(self.device)
                d_ck = torch.from_numpy(np.ascontiguousarray(g.ck[sel], dtype=np.float64)).to(self.device)
                d_sign = torch.from_numpy(np.ascontiguousarray(g.sign[:, sel].astype(np.uint8))).to(self.device) if has_sign else None
                self._keep += [d_idx, d_ck, d_sign]
                entries.append((c, g.loc_dof.shape[0], sel.size, d_idx, d_sign, d_ck, ke))
        entries.sort(key=lambda t: t[0])                        # stable: colour order, pattern groups in their order inside a colour
        cgroups = (_lib.EbeGroup * len(entries))()
        phase = (ctypes.c_int32 * len(entries))()
        for k, (c, nd, ne, d_idx, d_sign, d_ck, ke) in enumerate(entries):
            cg = cgroups[k]
            cg.nd, cg.ne = nd, ne
            cg.d_idx, cg.d_ck = d_idx.data_ptr(), d_ck.data_ptr()
            cg.d_sign = d_sign.data_ptr() if d_sign is not None else None
            cg.ke_host = ke.ctypes.data
            phase[k] = c
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().pcgb_ebe2_create(n, len(entries), cgroups, phase, ctypes.byref(self._h)), "pcgb_ebe2_create")

    def apply_local(self, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        y = out if out is not None else torch.empty(self.shape[0], dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().pcgb_ebe2_apply(self._h, _lib.ptr(x), _lib.ptr(y), _lib.stream_ptr()), "pcgb_ebe2_apply")
        return y

    def launches(self) -> int:
        return int(_lib.load().pcgb_ebe2_launches(self._h))

    def __del__(self):
        try:
            if self._h:
                _lib.load().pcgb_ebe2_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass
