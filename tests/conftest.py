import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long CPU test, enabled with PCGB_SLOW=1")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture
def build_c_demo():
    """Builds examples/cabi_demo.c with gcc (plain C, no nvcc, no Python) against include/pcgb200.h + libpcgb200.so + the CUDA runtime."""
    import subprocess

    def build(out):
        so_dir = os.path.join(ROOT, "pcg_mpi_solver_b200", "csrc")
        cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include", os.path.join(ROOT, "examples", "cabi_demo.c"),
               "-o", out, "-L" + so_dir, "-lpcgb200", "-L/usr/local/cuda/lib64", "-lcudart", "-lm", "-Wl,-rpath," + so_dir]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout
        return out

    return build
