"""Stand-in for the `mpi4py` package (TEST INFRASTRUCTURE, not product code).

The upstream reference imports `mpi4py` and `from mpi4py import MPI`
(pcg_solver.py:29-30, partition_mesh.py:20-21, file_operations.py:10-11).
mpi4py / libmpi are absent from this image, so the oracle harness puts this
directory first on PYTHONPATH to let the *unmodified* reference scripts run.
See MPI.py for the (small) surface that is implemented.
"""
from . import MPI  # noqa: F401
