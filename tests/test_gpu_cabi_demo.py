"""The drop-in boundary exercised from plain C (examples/cabi_demo.c): no Python, no torch in the process that solves."""
import subprocess

import pytest

pytestmark = pytest.mark.gpu


def test_c_demo_solves_c1_sized_poisson(cuda, tmp_path, build_c_demo):
    exe = build_c_demo(str(tmp_path / "cabi_demo"))
    r = subprocess.run([exe, "32"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)   # config C1: 32^3, n = 32768
    assert r.returncode == 0 and "CABI_DEMO_OK" in r.stdout, r.stdout
    assert "n=32768 nnz=830584" in r.stdout
