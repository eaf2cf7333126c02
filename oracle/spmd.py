"""ORACLE helper (test infrastructure): run the restated reference SPMD - one OS process per mesh part, like
`mpiexec -np P python3 pcg_solver.py` - with the reductions and the interface exchange going through POSIX shared
memory and barriers instead of MPI.  Used by the CPU reference arm of bench.py (`--impl reference`, `cpu_baseline`)
so that the timed CPU run contains the same communication pattern as the reference
(3 allreduces + 1 neighbour exchange per iteration, pcg_solver.py:303-334, 622-628).

    ShmComm.allreduce(v)                 <- MPI_SUM (pcg_solver.py:622-628), sum in rank order (deterministic)
    ShmComm.exchange_add_full(part, y)   <- pcg_solver.py:303-334 (pack, Isend/Recv, += in neighbour order)
"""
from __future__ import annotations

import multiprocessing as mp
from multiprocessing import shared_memory

import numpy as np


class ShmComm:
    MAXRED = 8
    TIMEOUT = 600.0   # seconds; a dead rank breaks the barrier instead of hanging the others

    def __init__(self, rank, size, barrier, red_name, table_name):
        self.rank, self.size, self.barrier = rank, size, barrier
        self._red_shm = shared_memory.SharedMemory(name=red_name)
        self.red = np.ndarray((size, self.MAXRED), dtype=np.float64, buffer=self._red_shm.buf)
        self._tab_shm = shared_memory.SharedMemory(name=table_name)
        self.table = np.ndarray((size, size), dtype=np.int64, buffer=self._tab_shm.buf)   # table[p, q] = doubles p sends to q
        self._halo_shm = None
        self.halo = None
        self.offsets = None

    # ---- MPI_SUM
    def allreduce(self, v):
        a = np.atleast_1d(np.asarray(v, dtype=np.float64))
        self.red[self.rank, :a.size] = a
        self.barrier.wait(self.TIMEOUT)
        tot = self.red[0, :a.size].copy()
        for r in range(1, self.size):
            tot = tot + self.red[r, :a.size]
        self.barrier.wait(self.TIMEOUT)
        return float(tot[0]) if np.isscalar(v) or np.ndim(v) == 0 else tot

    # ---- interface exchange
    def setup_halo(self, part, halo_name_box):
        """Collective: publish the send lengths, derive every pair's offset, attach the shared halo buffer."""
        self.table[self.rank, :] = 0
        for nb, idx in zip(part.nbr, part.ovrlp_full):
            self.table[self.rank, nb] = len(idx)
        self.barrier.wait(self.TIMEOUT)
        flat = self.table.ravel()
        self.offsets = np.concatenate([[0], np.cumsum(flat)]).reshape(-1)[:-1].reshape(self.size, self.size)
        total = int(flat.sum())
        if self.rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=max(8, 8 * total))
            halo_name_box.value = shm.name.encode()
            self._halo_owner = shm
        self.barrier.wait(self.TIMEOUT)
        self._halo_shm = shared_memory.SharedMemory(name=halo_name_box.value.decode())
        self.halo = np.ndarray((max(1, total),), dtype=np.float64, buffer=self._halo_shm.buf)
        self.barrier.wait(self.TIMEOUT)

    def exchange_add_full(self, part, y):
        for nb, idx in zip(part.nbr, part.ovrlp_full):                       # pack + "Isend" (:307-322)
            o = self.offsets[self.rank, nb]
            self.halo[o:o + len(idx)] = y[idx]
        self.barrier.wait(self.TIMEOUT)
        for nb, idx in zip(part.nbr, part.ovrlp_full):                       # "Recv" + += in neighbour order (:324-334)
            o = self.offsets[nb, self.rank]
            y[idx] += self.halo[o:o + len(idx)]
        self.barrier.wait(self.TIMEOUT)
        return y

    def close(self):
        for s in (self._red_shm, self._tab_shm, self._halo_shm):
            try:
                if s is not None:
                    s.close()
            except Exception:
                pass


def _worker(rank, size, barrier, red_name, table_name, halo_box, fn, args, conn):
    import os
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[rank % len(os.sched_getaffinity(0))]})
    except Exception:
        pass
    comm = ShmComm(rank, size, barrier, red_name, table_name)
    comm._halo_box = halo_box
    try:
        conn.send(("ok", fn(rank, size, comm, *args)))
    except Exception as e:  # pragma: no cover
        import traceback
        conn.send(("error", traceback.format_exc()))
        try:
            barrier.abort()
        except Exception:
            pass
    finally:
        comm.close()


def run_spmd(size, fn, args=()):
    """Fork `size` processes running fn(rank, size, comm, *args); returns the list of their return values."""
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(size)
    red = shared_memory.SharedMemory(create=True, size=8 * size * ShmComm.MAXRED)
    table = shared_memory.SharedMemory(create=True, size=8 * size * size)
    halo_box = ctx.Array("c", 64)
    procs, conns = [], []
    for r in range(size):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_worker, args=(r, size, barrier, red.name, table.name, halo_box, fn, args, b))
        p.start()
        procs.append(p)
        conns.append(a)
    out = []
    import time as _time
    deadline = _time.time() + 1800.0
    for c, p in zip(conns, procs):
        while not c.poll(1.0):
            if not p.is_alive() or _time.time() > deadline:
                break
        out.append(c.recv() if c.poll(0) else ("error", f"rank process exited without a result (exitcode {p.exitcode})"))
    for p in procs:
        p.join(5)
        if p.is_alive():
            p.terminate()
    for s in (red, table):
        s.close()
        s.unlink()
    try:
        if halo_box.value:
            h = shared_memory.SharedMemory(name=halo_box.value.decode())
            h.close()
            h.unlink()
    except Exception:
        pass
    errs = [o[1] for o in out if o[0] == "error"]
    if errs:
        raise RuntimeError("SPMD worker failed:\n" + errs[0])
    return [o[1] for o in out]
