"""GPU parity on the reference's own model (data/concrete.zip, config C4) against golden vectors produced
by the UNMODIFIED reference (tests/golden/concrete_*): product builder -> device assembly -> CUDA PCG."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _zip():
    for p in (os.path.join(ROOT, "oracle", "_ref", "concrete.zip"), "/root/reference/data/concrete.zip"):
        if os.path.exists(p):
            return p
    pytest.skip("data/concrete.zip not staged")


@pytest.fixture(scope="module")
def concrete(cuda):
    from pcg_mpi_solver_b200.partition import partition_mesh
    sub = partition_mesh(_zip(), 1, assemble=False)[0]
    op = sub.to_operator(device=cuda)
    return sub, op


def test_concrete_assembly_invariants(concrete):
    """Golden G4: nnz, row lengths, diagonal range, symmetry of A = K[Eff,Eff]."""
    import torch
    sub, op = concrete
    A = op.A
    assert A.shape == (616413, 616413) and A.nnz == 73073097
    lens = torch.diff(A.rowptr.to(torch.int64))
    assert int(lens.min()) == 24 and int(lens.max()) == 324
    d = A.diagonal()
    assert abs(float(d.min()) - 1.2322e6) < 1e2 and abs(float(d.max()) - 1.5383e8) < 1e4
    x = torch.randn(A.shape[0], dtype=torch.float64, device=A.device)
    y = torch.randn_like(x)
    assert abs(float(torch.dot(y, A.spmv(x)) - torch.dot(x, A.spmv(y)))) <= 1e-11 * float(torch.dot(y.abs(), A.spmv(x).abs()))


def test_concrete_operator_matches_reference_probe(concrete):
    """SpMV of the assembled matrix == the reference's own calcMPFint on the same input vector."""
    import torch
    sub, op = concrete
    with open(os.path.join(GOLD, "concrete_probe.json")) as f:
        g = json.load(f)
    arr = np.load(os.path.join(GOLD, "concrete_probe.npz"))
    dofv = sub.dof_vector
    v = np.sin(0.001 * dofv) + 0.25 * np.cos(0.37 * dofv)
    v_eff = v[sub.loc_dof_eff]
    y = op.apply(torch.from_numpy(v_eff).to(op.device)).cpu().numpy()
    yfull = np.zeros(sub.ndof)
    yfull[sub.loc_dof_eff] = y
    sel = np.isin(arr["idx"], sub.loc_dof_eff)
    scale = np.abs(arr["y"]).max()
    assert np.abs(yfull[arr["idx"]][sel] - arr["y"][sel]).max() <= 1e-12 * scale
    minv = op.jacobi().cpu().numpy()
    np.testing.assert_allclose(minv[::101], arr["minv"], rtol=1e-13)
    np.testing.assert_allclose(np.linalg.norm(minv), g["norm_minv"], rtol=1e-13)


def test_concrete_solve_matches_reference(concrete):
    """Golden G1: the reference's run (1 part, Tol 1e-7): Flag 0, Iter 1085, RelRes 9.653e-08, ||U||."""
    import torch
    sub, op = concrete
    with open(os.path.join(GOLD, "concrete_ref.json")) as f:
        g = json.load(f)
    run = g["runs"]["1"]
    b = torch.from_numpy(sub.b).to(op.device)
    minv = op.jacobi()
    x, info = op.solve(b, minv, g["Tol"], g["MaxIter"])
    assert info.flag == run["Flag"] == 0
    assert abs(info.iters - run["Iter"]) <= 2, (info.iters, run["Iter"])
    assert info.relres <= g["Tol"]
    u = np.zeros(g["GlobNDof"])
    u[sub.dof_eff_global] = x.cpu().numpy()      # Un = X_unq + Udi, Ud = 0 (pcg_solver.py:598)
    # both runs stop at RelRes ~1e-7; the reference differs from ITSELF by 6e-11 between 1 and 8 parts
    assert abs(np.linalg.norm(u) - run["norm2_U"]) <= 1e-7 * run["norm2_U"]
    s = np.load(os.path.join(GOLD, "concrete_ref_samples.npz"))
    assert np.abs(u[s["idx"]] - s["U1"]).max() <= 1e-6 * np.abs(s["U1"]).max()
    if info.iters == run["Iter"]:
        assert abs(info.relres - run["RelRes"]) <= 2e-2 * run["RelRes"]   # the reference's own 1- and 8-part runs differ by 0.8 % here
        assert np.abs(u[s["idx"]] - s["U1"]).max() <= 1e-8 * np.abs(s["U1"]).max()
    # true residual with the device operator
    r = b - op.apply(x)
    assert float(torch.linalg.norm(r) / torch.linalg.norm(b)) <= g["Tol"] * (1 + 1e-6)
    # config C4's tolerance (golden G3 of the survey: ~1154 iterations at Tol 1e-8)
    x8, info8 = op.solve(b, minv, 1e-8, g["MaxIter"])
    assert info8.flag == 0 and info8.relres <= 1e-8 and abs(info8.iters - 1154) <= 5
