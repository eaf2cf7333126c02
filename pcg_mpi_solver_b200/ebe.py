"""EXPERIMENTAL matrix-free element-by-element operator - the reference's own operator form
(calcMatVecProd, pcg_solver.py:263-300) on the GPU instead of the assembled CSR matrix.

Status: written at the end of round 1; its parity tests (tests/test_gpu_ebe.py: hex vs CSR/oracle, concrete vs the
reference's 1085 iterations) are green on a B200 and a first timing exists (profiles/ebe_quick_r1.json: 0.27 ms per
application and 2550 PCG it/s on the 128^3 hex box, vs 0.93 ms / 964 it/s for the CSR kernel), but it has not been
profiled or tuned and scatter-adds with fp64 atomics (not bit-reproducible).  Opt-in: `to_operator(kind="ebe")`.
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
