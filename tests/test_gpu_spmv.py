"""GPU parity tests of the merge-path SpMV and the vector kernels against the CPU oracle
(oracle/ref_pcg.py generators + scipy), called through the C ABI (libpcgb200.so via ctypes)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import ref_pcg as R

pytestmark = pytest.mark.gpu

RTOL = 1e-13  # per-entry |y - y_ref| <= RTOL * (|A| |x|)_i : fp64 summation-order noise only


def _check_spmv(A, cuda, seed=0, **env):
    import torch
    from pcg_mpi_solver_b200.csr import CsrMatrix
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(A.shape[1])
    y_ref = A @ x
    scale = abs(A) @ np.abs(x) + 1e-300
    M = CsrMatrix.from_scipy(A, device=cuda)
    y = M.spmv(torch.from_numpy(x).to(cuda)).cpu().numpy()
    err = np.abs(y - y_ref) / scale
    assert err.max() <= RTOL, (err.max(), M.plan_info())
    return M


def test_spmv_poisson27_c1(cuda):
    A = R.poisson27(32)  # config C1: n = 32768, nnz = 830584
    assert A.nnz == 830584
    M = _check_spmv(A, cuda)
    assert M.plan_info()["snap"] == 1 and M.plan_info()["split_rows"] == 0


def test_spmv_hex_box(cuda):
    A = R.hex_box_csr((12, 10, 8), (0, 0, 0), (12, 10, 8))
    _check_spmv(A, cuda)


@pytest.mark.parametrize("lanes", [4, 8, 16, 32])
@pytest.mark.parametrize("variant", ["ldg", "tma", "staged", "persist"])
def test_spmv_variants(cuda, monkeypatch, lanes, variant):
    """The three kernels (plain loads / TMA-staged val+col with L1 gather / TMA + x staged in shared memory
    with 16-bit local indices) and every lanes-per-row instantiation agree with scipy."""
    monkeypatch.setenv("PCGB_SPMV_LANES", str(lanes))
    monkeypatch.setenv("PCGB_SPMV_TMA", "0" if variant == "ldg" else "1")
    monkeypatch.setenv("PCGB_SPMV_STAGE", "1" if variant in ("staged", "persist") else "0")
    monkeypatch.setenv("PCGB_SPMV_PERSIST", "1" if variant == "persist" else "0")
    monkeypatch.setenv("PCGB_SPMV_T3", "0")     # one 16-bit index per non-zero: the row-group kernels
    monkeypatch.setenv("PCGB_SPMV_BSR", "0")    # (the node-block kernel, default for 3-dofs-per-node matrices, has its own tests)
    A = R.hex_box_csr((9, 7, 5), (0, 0, 0), (9, 7, 5))
    M = _check_spmv(A, cuda, seed=lanes)
    info = M.plan_info()
    assert info["lanes"] == lanes and info["tma"] == (variant != "ldg")
    assert info["staged"] == {"ldg": 0, "tma": 0, "staged": 1, "persist": 2}[variant]
    if variant in ("staged", "persist"):
        assert M.stream_bytes() < M.spmv_bytes()  # 10 B/nnz instead of 12


@pytest.mark.parametrize("tile", [256, 512, 2048])
@pytest.mark.parametrize("gap", [0, 8, 64])
def test_spmv_staged_tiles_and_gaps(cuda, monkeypatch, tile, gap):
    monkeypatch.setenv("PCGB_SPMV_TILE", str(tile))
    monkeypatch.setenv("PCGB_SPMV_GAP", str(gap))
    for A in (R.poisson27(11), R.hex_box_csr((7, 6, 5), (2, 0, 0), (5, 6, 5))):
        M = _check_spmv(A, cuda, seed=tile + gap)
        assert M.plan_info()["staged"] == 2  # persistent pipelined kernel


@pytest.mark.parametrize("stages,ctas", [(4, 1), (4, 2), (8, 1)])
def test_spmv_persist_pipeline_shapes(cuda, monkeypatch, stages, ctas):
    """Ring depth / CTAs per SM of the persistent kernel; many more tiles than CTAs so every stage wraps."""
    monkeypatch.setenv("PCGB_SPMV_STAGES", str(stages))
    monkeypatch.setenv("PCGB_SPMV_CTAS", str(ctas))
    monkeypatch.setenv("PCGB_SPMV_TILE", "256")
    A = R.hex_box_csr((24, 20, 16), (0, 0, 0), (24, 20, 16))
    M = _check_spmv(A, cuda, seed=stages)
    info = M.plan_info()
    assert info["staged"] == 2 and info["ntiles"] > 2 * stages * 148 * ctas


def test_spmv_staged_falls_back_when_not_stageable(cuda, monkeypatch):
    """A tile whose columns are scattered over more than the shared-memory budget keeps the L1-gather kernel."""
    n = 400000
    rng = np.random.default_rng(12)
    rows = np.repeat(np.arange(2000), 40)
    cols = rng.integers(0, n, size=rows.size)
    A = sp.csr_matrix((rng.standard_normal(rows.size), (rows, cols)), shape=(2000, n))
    A.sum_duplicates()
    M = _check_spmv(A, cuda)
    assert M.plan_info()["staged"] == 0 and M.stream_bytes() >= M.spmv_bytes()


@pytest.mark.parametrize("tile", [256, 1024, 4096])
def test_spmv_split_rows(cuda, monkeypatch, tile):
    """Rows longer than a tile: pure merge-path split with the carry fix-up pass."""
    monkeypatch.setenv("PCGB_SPMV_TILE", str(tile))
    rng = np.random.default_rng(5)
    n = 3000
    A = sp.random(n, n, density=0.002, random_state=7, format="lil")
    for r in (0, 17, 1500, n - 1):  # dense rows spanning many tiles
        A[r, :] = rng.standard_normal(n)
    A[5, :] = 0  # empty rows
    A[6, :] = 0
    A = A.tocsr()
    M = _check_spmv(A, cuda)
    info = M.plan_info()
    assert info["snap"] == 0 and info["split_rows"] > 0


def test_spmv_ragged_and_empty(cuda):
    rng = np.random.default_rng(3)
    n = 1237
    A = sp.random(n, n, density=0.01, random_state=11, format="csr")
    A.data[:] = rng.standard_normal(A.nnz)
    _check_spmv(A, cuda)
    # rectangular with trailing empty rows, nnz not a multiple of 4
    B = sp.vstack([A[:100], sp.csr_matrix((7, n))]).tocsr()
    _check_spmv(B, cuda)
    # a single entry
    C = sp.csr_matrix(([2.5], ([3], [2])), shape=(6, 6))
    _check_spmv(C, cuda)
    # no entries at all
    Z = sp.csr_matrix((5, 5))
    _check_spmv(Z, cuda)


def test_spmv_int64_offsets(cuda):
    import torch
    from pcg_mpi_solver_b200.csr import CsrMatrix
    A = R.hex_box_csr((6, 6, 6), (0, 0, 0), (6, 6, 6))
    x = np.random.default_rng(1).standard_normal(A.shape[1])
    M32 = CsrMatrix.from_scipy(A, device=cuda, index64=False)
    M64 = CsrMatrix.from_scipy(A, device=cuda, index64=True)
    xd = torch.from_numpy(x).to(cuda)
    y32, y64 = M32.spmv(xd), M64.spmv(xd)
    assert torch.equal(y32, y64)  # same plan -> bit-identical
    assert M64.spmv_bytes() - M32.spmv_bytes() == 4 * (A.shape[0] + 1)


def test_spmv_deterministic(cuda):
    import torch
    from pcg_mpi_solver_b200.csr import CsrMatrix
    A = R.poisson27(20)
    M = CsrMatrix.from_scipy(A, device=cuda)
    x = torch.randn(A.shape[1], dtype=torch.float64, device=cuda)
    y0 = M.spmv(x).clone()
    for _ in range(5):
        assert torch.equal(M.spmv(x), y0)


def test_diag_and_vector_kernels(cuda):
    import ctypes
    import torch
    from pcg_mpi_solver_b200 import _lib
    from pcg_mpi_solver_b200.csr import CsrMatrix
    A = R.hex_box_csr((5, 5, 5), (0, 0, 0), (5, 5, 5))
    M = CsrMatrix.from_scipy(A, device=cuda)
    np.testing.assert_allclose(M.diagonal().cpu().numpy(), A.diagonal(), rtol=1e-14)
    lib = _lib.load()
    rng = np.random.default_rng(0)
    n = 100003
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    w = (rng.random(n) > 0.3).astype(float)
    ad, bd, wd = (torch.from_numpy(v).to(cuda) for v in (a, b, w))
    out = torch.zeros(1, dtype=torch.float64, device=cuda)
    _lib.check(lib.pcgb_dot_w(n, _lib.ptr(ad), _lib.ptr(bd), _lib.ptr(wd), _lib.ptr(out), _lib.stream_ptr()))
    ref = np.dot(a, b * w)  # np.dot(a, b*w) as pcg_solver.py:462
    assert abs(out.item() - ref) <= 1e-12 * np.dot(np.abs(a), np.abs(b) * w)
    _lib.check(lib.pcgb_dot_w(n, _lib.ptr(ad), _lib.ptr(bd), None, _lib.ptr(out), _lib.stream_ptr()))
    assert abs(out.item() - np.dot(a, b)) <= 1e-12 * np.dot(np.abs(a), np.abs(b))
    yd = bd.clone()
    _lib.check(lib.pcgb_axpby(n, 0.5, _lib.ptr(ad), -2.0, _lib.ptr(yd), _lib.stream_ptr()))
    np.testing.assert_allclose(yd.cpu().numpy(), 0.5 * a - 2.0 * b, rtol=1e-15, atol=1e-15)
    _lib.check(lib.pcgb_dot_w(0, None, None, None, _lib.ptr(out), _lib.stream_ptr()))  # empty input
    assert out.item() == 0.0


@pytest.mark.parametrize("lanes3", [4, 8, 16, 32])
@pytest.mark.parametrize("tile", [512, 2304])
def test_spmv_triple_index(cuda, monkeypatch, lanes3, tile):
    """Column-triple index: 3 dofs per node make every row a sequence of aligned triples of consecutive columns; the
    persistent kernel then streams ONE 16-bit staged position per triple (8 + 2/3 B per non-zero instead of 10)."""
    monkeypatch.setenv("PCGB_SPMV_T3", "1")          # opt-in mode (slower than the per-non-zero index on B200 so far)
    monkeypatch.setenv("PCGB_SPMV_BSR", "0")
    monkeypatch.setenv("PCGB_SPMV_LANES3", str(lanes3))
    monkeypatch.setenv("PCGB_SPMV_TILE", str(tile))
    A = R.hex_box_csr((9, 7, 5), (0, 0, 0), (9, 7, 5))
    M = _check_spmv(A, cuda, seed=lanes3)
    info = M.plan_info()
    assert info["index_mode"] == 1 and info["staged"] == 2 and info["lanes"] == lanes3
    assert M.stream_bytes() < 8.8 * A.nnz + 40 * A.shape[0] + 64 * info["ntiles"]
    # an interior (unclamped) box and a box with a ragged last tile
    _check_spmv(R.hex_box_csr((8, 6, 4), (4, 0, 2), (4, 3, 2)), cuda, seed=1)
    # the fused dot epilogue of the PCG loop goes through the same kernel: compare x.(A x)
    import torch
    from pcg_mpi_solver_b200 import _lib
    x = torch.from_numpy(np.random.default_rng(3).standard_normal(A.shape[0])).to(cuda)
    M.set_boundary_rows(torch.arange(0, A.shape[0], 7, dtype=torch.int32, device=cuda))
    y, d = M.spmv_split(x, with_dot=True)
    y_ref = A @ x.cpu().numpy()
    assert np.abs(y.cpu().numpy() - y_ref).max() <= 1e-12 * np.abs(y_ref).max()
    assert abs(float(d) - float(x.cpu().numpy() @ y_ref)) <= 1e-12 * float(np.abs(x.cpu().numpy()) @ np.abs(y_ref))


def test_spmv_triple_index_not_applicable(cuda, monkeypatch):
    """Rows that are not made of column triples keep the per-non-zero index (Poisson: 27 single columns per row)."""
    monkeypatch.setenv("PCGB_SPMV_T3", "1")
    monkeypatch.setenv("PCGB_SPMV_BSR", "0")
    M = _check_spmv(R.poisson27(12), cuda)
    assert M.plan_info()["index_mode"] == 0
    # triples by count but not consecutive columns
    rng = np.random.default_rng(5)
    n = 300
    rows = np.repeat(np.arange(n), 6)
    cols = np.concatenate([np.sort(rng.choice(n, 6, replace=False)) for _ in range(n)])
    A = sp.csr_matrix((rng.standard_normal(rows.size), (rows, cols)), shape=(n, n))
    A.sort_indices()
    M = _check_spmv(A, cuda, seed=2)
    assert M.plan_info()["index_mode"] == 0


@pytest.mark.parametrize("t3", ["0", "1", "bsr"])
def test_spmv_interface_first_split(cuda, monkeypatch, t3):
    """Interface-first split (multi-GPU overlap): tiles owning a registered row run in a first launch, the rest in a
    second one; y is bit-identical to the single launch (same tiles, same arithmetic), x.y agrees to rounding."""
    import torch
    from pcg_mpi_solver_b200.csr import CsrMatrix
    monkeypatch.setenv("PCGB_SPMV_T3", "1" if t3 == "1" else "0")
    monkeypatch.setenv("PCGB_SPMV_BSR", "1" if t3 == "bsr" else "0")
    monkeypatch.setenv("PCGB_BSR_MIN_UNIFORM_PCT", "0")
    monkeypatch.setenv("PCGB_SPMV_TILE", "512" if t3 != "bsr" else "768")
    A = R.hex_box_csr((12, 10, 8), (0, 0, 0), (12, 10, 8))
    n = A.shape[0]
    M = CsrMatrix.from_scipy(A, device=cuda)
    assert M.plan_info()["interface_tiles"] == -1
    x = torch.from_numpy(np.random.default_rng(6).standard_normal(n)).to(cuda)
    y0 = M.spmv(x).clone()
    # the x = max face of the box: every 12th node -> scattered tiles; plus one whole z plane -> contiguous tiles
    node = np.arange(n // 3)
    face = np.nonzero((node % 12 == 11) | (node // (12 * 11) == 3))[0]
    rows = (3 * face[:, None] + np.arange(3)[None, :]).ravel().astype(np.int32)
    M.set_boundary_rows(torch.from_numpy(rows).to(cuda))
    info = M.plan_info()
    assert 0 < info["interface_tiles"] < info["ntiles"]
    y1, d = M.spmv_split(x, with_dot=True)
    assert torch.equal(y0, y1)
    ref = float(torch.dot(x, y0))
    assert abs(float(d) - ref) <= 1e-12 * float(torch.dot(x.abs(), y0.abs()))
    # degenerate registrations: nothing / everything is interface
    M.set_boundary_rows(torch.zeros(0, dtype=torch.int32, device=cuda))
    assert M.plan_info()["interface_tiles"] == 0 and torch.equal(M.spmv_split(x), y0)
    M.set_boundary_rows(torch.arange(n, dtype=torch.int32, device=cuda))
    assert M.plan_info()["interface_tiles"] == M.plan_info()["ntiles"] and torch.equal(M.spmv_split(x), y0)


def test_release_col(cuda):
    """The persistent kernel never reads the 4-byte column array: after release_col() SpMV and the Jacobi diagonal still work."""
    import torch
    from pcg_mpi_solver_b200 import _lib
    from pcg_mpi_solver_b200.csr import CsrMatrix
    A = R.hex_box_csr((9, 7, 5), (0, 0, 0), (9, 7, 5))
    M = CsrMatrix.from_scipy(A, device=cuda)
    x = torch.from_numpy(np.random.default_rng(7).standard_normal(A.shape[0])).to(cuda)
    y0, d0 = M.spmv(x).clone(), M.diagonal().clone()
    assert M.release_col() and M.col is None and M.plan_info()["col_released"] == 1
    torch.cuda.empty_cache()
    junk = torch.full((A.nnz,), -1, dtype=torch.int32, device=cuda)   # likely reuses the freed block
    assert torch.equal(M.spmv(x), y0) and torch.equal(M.diagonal(), d0)
    del junk
    with pytest.raises(_lib.PcgbError):
        M.to_scipy()
    # a plan that gathers through L1 needs the columns: release is refused, nothing changes
    P = CsrMatrix.from_scipy(R.poisson27(6), device=cuda)
    if P.plan_info()["staged"] != 2:
        assert not P.release_col() and P.col is not None


@pytest.mark.parametrize("inplace,cw", [(1, 6), (0, 6), (1, 8), (0, 8)])
@pytest.mark.parametrize("uni", [1, 0])
@pytest.mark.parametrize("tile", [0, 600, 1200, 2058])
def test_spmv_node_block_kernel(cuda, monkeypatch, tile, uni, inplace, cw):
    """Node-block ("BSR-3") kernel, the default for 3-dofs-per-node matrices: one thread per 3x3 block, one 16-bit staged position
    per block (8 + 2/9 B per non-zero), rows summed from shared-memory partials in block order.  Several tile sizes (one and
    several passes of 256 blocks, 2- and 4-stage rings), clamped / interior boxes, fused dot, bit-reproducibility."""
    import torch
    from pcg_mpi_solver_b200.csr import CsrMatrix
    monkeypatch.setenv("PCGB_BSR_MIN_UNIFORM_PCT", "0")     # small boxes are mostly boundary: do not let the regularity gate decide
    monkeypatch.setenv("PCGB_BSR_UNI", str(uni))            # uniform-tile fast path on / off (off: per-block node search everywhere)
    monkeypatch.setenv("PCGB_BSR_INPLACE", str(inplace))    # row partials in the stage itself (larger stages) / in a separate scratch
    monkeypatch.setenv("PCGB_BSR_CW", str(cw))              # consumer warps per CTA
    if tile:
        monkeypatch.setenv("PCGB_SPMV_TILE", str(tile))
    for box in [((9, 7, 5), (0, 0, 0), (9, 7, 5)), ((8, 6, 4), (4, 0, 2), (4, 3, 2)), ((14, 12, 10), (0, 0, 0), (14, 12, 10))]:
        A = R.hex_box_csr(*box)
        M = _check_spmv(A, cuda, seed=tile + uni)
        info = M.plan_info()
        assert info["index_mode"] == 2 and info["staged"] == 2 and info["split_rows"] == 0, info
        assert M.stream_bytes() < 8.3 * A.nnz + 40 * A.shape[0] + 64 * info["ntiles"]
        x = torch.from_numpy(np.random.default_rng(1).standard_normal(A.shape[0])).to(cuda)
        y = M.spmv(x).clone()
        assert torch.equal(M.spmv(x), y)                                   # fixed summation order
        M.set_boundary_rows(torch.arange(0, A.shape[0], 11, dtype=torch.int32, device=cuda))
        y2, d = M.spmv_split(x, with_dot=True)
        assert torch.equal(y2, y)
        ref = float(torch.dot(x, y))
        assert abs(float(d) - ref) <= 1e-12 * float(torch.dot(x.abs(), y.abs()))


def test_spmv_node_block_irregular_nodes(cuda, monkeypatch):
    """Nodes with different numbers of blocks per row (a random node graph expanded to 3x3 blocks) - the octree / concrete shape.
    Such matrices stay on the row-group kernel by default (regularity gate); forced here to cover the per-block node search."""
    monkeypatch.setenv("PCGB_BSR_MIN_UNIFORM_PCT", "0")
    monkeypatch.setenv("PCGB_SPMV_TILE", "1500")           # random columns: one x window per block - stay below the 256-window budget
    rng = np.random.default_rng(12)
    nn = 700
    rows, cols = [], []
    for n in range(nn):
        k = int(rng.integers(1, 40))
        nb = np.unique(np.concatenate([[n], rng.choice(nn, k, replace=False)]))
        rows.append(np.full(nb.size, n)); cols.append(nb)
    G = sp.csr_matrix((np.ones(sum(c.size for c in cols)), (np.concatenate(rows), np.concatenate(cols))), shape=(nn, nn))
    A = sp.kron(G, np.ones((3, 3))).tocsr()
    A.data = rng.standard_normal(A.nnz)
    A.sort_indices()
    M = _check_spmv(A, cuda, seed=4)
    assert M.plan_info()["index_mode"] == 2
    # one row of one node differs from its siblings -> not a node-block matrix -> row-group kernel, same answer
    B = A.tolil()
    assert B[4, 5] != 0                       # (node 1, node 1) block exists: drop one entry of its middle row
    B[4, 5] = 0.0
    B = B.tocsr(); B.eliminate_zeros(); B.sort_indices()
    M = _check_spmv(B, cuda, seed=5)
    assert M.plan_info()["index_mode"] == 0
    # default gate: an irregular node graph is not sent to the node-block kernel
    monkeypatch.delenv("PCGB_BSR_MIN_UNIFORM_PCT")
    monkeypatch.delenv("PCGB_SPMV_TILE")
    assert _check_spmv(A, cuda, seed=6).plan_info()["index_mode"] == 0
