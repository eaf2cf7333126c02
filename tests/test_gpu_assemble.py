"""Device assembly of K_i[Eff,Eff] (csrc/assemble.cuh) against the host COO assembly of the same pattern groups
(partition._assemble: sum_e P_e^T (Ck_e S_e Ke S_e) P_e, the operator calcMatVecProd applies, pcg_solver.py:263-300)."""
import os

import numpy as np
import pytest

from oracle.hex_mdf import write_hex_mdf

pytestmark = pytest.mark.gpu


def _compare(sub, cuda):
    import torch
    from pcg_mpi_solver_b200.partition import _assemble, assemble_csr_device
    K = _assemble(sub.groups, sub.ndof)
    A = K[sub.loc_dof_eff][:, sub.loc_dof_eff].tocsr()
    A.sort_indices()
    rowptr, col, val = assemble_csr_device(sub, cuda)
    assert np.array_equal(rowptr.cpu().numpy().astype(np.int64), A.indptr.astype(np.int64))      # structure: exact
    assert np.array_equal(col.cpu().numpy(), A.indices)
    v = val.cpu().numpy()
    assert np.abs(v - A.data).max() <= 1e-14 * np.abs(A.data).max()                              # values: summation order only
    r2, c2, v2 = assemble_csr_device(sub, cuda)                                                   # bit-reproducible
    assert torch.equal(v2, val) and torch.equal(c2, col) and torch.equal(r2, rowptr)
    return A


def test_device_assembly_hex_parts(cuda, tmp_path):
    from pcg_mpi_solver_b200.model import load_mdf
    from pcg_mpi_solver_b200.partition import partition_mesh
    write_hex_mdf(str(tmp_path), (7, 6, 5))
    model = load_mdf(str(tmp_path), name="hexmodel")
    for sub in partition_mesh(model, 1, assemble=False) + partition_mesh(model, 3, assemble=False):
        _compare(sub, cuda)


def test_device_assembly_mixed_pattern_groups_with_signs(cuda):
    """Several pattern sizes, sign flips, clamped dofs, rows touched by many elements (synthetic)."""
    from pcg_mpi_solver_b200.partition import SubdomainData, TypeGroup
    rng = np.random.default_rng(3)
    ndof = 400
    groups = []
    for t, (nd, ne) in enumerate([(24, 60), (33, 25), (78, 9), (3, 40)]):
        loc = np.stack([rng.choice(ndof, nd, replace=False) for _ in range(ne)], axis=1).astype(np.int64)     # (nd, ne), distinct per element
        ke = rng.standard_normal((nd, nd))
        ke = ke + ke.T
        groups.append(TypeGroup(t, loc, rng.random((nd, ne)) < 0.2, rng.random(ne) + 0.5, ke, np.arange(ne)))
    eff = np.sort(rng.choice(ndof, 350, replace=False))
    sub = SubdomainData(0, 1, np.arange(ndof), np.arange(ndof // 3), eff, groups, [], [], [], np.ones(ndof), np.zeros(ndof), np.zeros(ndof), eff.size, ndof)
    _compare(sub, cuda)


def test_device_assembly_concrete_part(cuda):
    from pcg_mpi_solver_b200.partition import partition_mesh
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    zp = os.path.join(root, "oracle", "_ref", "concrete.zip")
    if not os.path.exists(zp):
        pytest.skip("concrete.zip not staged")
    ep = np.load(os.path.join(root, "tests", "golden", "concrete_elepart_8.npy")).astype(np.int64)
    sub = partition_mesh(zp, 8, elepart=ep, assemble=False)[3]
    A = _compare(sub, cuda)
    assert A.nnz > 5_000_000
