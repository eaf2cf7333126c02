"""GPU run of the displacement-controlled two-step ramp through the file-compatible stage (CUDA backend) against the
UNMODIFIED reference's golden frames.  (Sorted last on purpose: added after the last GPU session of round 1.)"""
import json
import os

import numpy as np
import pytest

from oracle import run_reference as rr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_cli_gpu_ramp_matches_reference(cuda, tmp_path):
    from pcg_mpi_solver_b200.partition import partition_mesh
    from pcg_mpi_solver_b200.pcg_solver import export_mesh_parts, run
    from tests.test_pull_ramp import _model, _workdir
    with open(os.path.join(GOLD, "hex_pull_ref.json")) as f:
        meta = json.load(f)
    arr = np.load(os.path.join(GOLD, "hex_pull_ref.npz"))
    model, info = _model(tmp_path, meta)
    sub = partition_mesh(model, 1, assemble=False)[0]
    work = _workdir(tmp_path, meta)
    export_mesh_parts(os.path.join(work, "data", "ModelData", "MPI") + "/", [sub])
    out = run(1, 0, workdir=work, quiet=True)
    run1 = meta["runs"]["p1"]
    assert list(out["Flag"]) == [0, 0, 0]
    assert all(abs(int(a) - b) <= 2 for a, b in zip(out["Iter"], run1["Iter"]))
    for frame in (1, 2):
        _, u = rr.read_results(work, "hexpull", 1, 1, info["ndof"], frame=frame)
        ref = arr[f"U{frame}_p1"]
        assert np.linalg.norm(u - ref) <= 1e-8 * np.linalg.norm(ref)
