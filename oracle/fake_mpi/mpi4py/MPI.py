"""Fake `mpi4py.MPI` for the oracle harness (TEST INFRASTRUCTURE ONLY).

Implements exactly the surface the upstream reference touches
(SURVEY.md section 2.2): COMM_WORLD / COMM_SELF with
Get_rank, Get_size, barrier, Split_type, gather, bcast, scatter, allreduce,
Allgather, Isend, Recv, isend, recv; Request.Waitall; Win.Allocate_shared /
Shared_query; File.Open / Write / Write_at / Read / Read_at / Close; the
datatype objects LONG, DOUBLE, BOOL (Get_size only); SUM; MODE_*;
COMM_TYPE_SHARED.

Two modes:
  * single rank (default)  - everything is local.
  * multi rank             - `oracle/run_reference.py` forks one OS process per
    rank and calls `_attach(rank, size, inboxes, barrier)` before the reference
    script is executed with runpy; point-to-point traffic goes through one
    multiprocessing.Queue inbox per rank, collectives are built from it with
    rank 0 as root and a fixed (rank-order) summation for allreduce.
Nothing here is used by the product path.
"""
import os
import pickle

import numpy as np

SUM = "SUM"
COMM_TYPE_SHARED = 1
MODE_RDONLY, MODE_WRONLY, MODE_CREATE, MODE_RDWR = 1, 2, 4, 8
ANY_SOURCE, ANY_TAG = -1, -1


class _Datatype:
    def __init__(self, nbytes):
        self._n = nbytes

    def Get_size(self):
        return self._n


LONG, DOUBLE, BOOL = _Datatype(8), _Datatype(8), _Datatype(1)

_state = {"rank": 0, "size": 1, "inboxes": None, "barrier": None, "pending": [], "coll": 0}


def _attach(rank, size, inboxes, barrier):
    _state.update(rank=rank, size=size, inboxes=inboxes, barrier=barrier, pending=[], coll=0)


class Request:
    def Wait(self):
        return None

    wait = Wait

    @staticmethod
    def Waitall(reqs):
        return None


_COLL_TAG0 = 1 << 40


def _send(dest, tag, payload):
    _state["inboxes"][dest].put((_state["rank"], tag, payload))


def _recv(source, tag):
    pend = _state["pending"]
    for k, (s, t, p) in enumerate(pend):
        if (source in (ANY_SOURCE, s)) and (tag in (ANY_TAG, t)):
            pend.pop(k)
            return p
    inbox = _state["inboxes"][_state["rank"]]
    while True:
        s, t, p = inbox.get()
        if (source in (ANY_SOURCE, s)) and (tag in (ANY_TAG, t)):
            return p
        pend.append((s, t, p))


class _SelfComm:
    """COMM_SELF / result of Split_type in single-node runs restricted to one rank."""

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def barrier(self):
        return None

    Barrier = barrier


class _WorldComm:
    def Get_rank(self):
        return _state["rank"]

    def Get_size(self):
        return _state["size"]

    def barrier(self):
        if _state["size"] > 1:
            _state["barrier"].wait()

    Barrier = barrier

    def Split_type(self, kind):
        # all fake ranks live on one node -> the shared communicator is the world
        return self

    # ---- point to point -------------------------------------------------
    def Isend(self, buf, dest, tag=0):
        _send(dest, tag, np.array(buf, copy=True))
        return Request()

    def Send(self, buf, dest, tag=0):
        self.Isend(buf, dest, tag)

    def Recv(self, buf, source=ANY_SOURCE, tag=ANY_TAG):
        data = _recv(source, tag)
        buf[...] = np.asarray(data).reshape(buf.shape)

    def isend(self, obj, dest, tag=0):
        _send(dest, tag, pickle.dumps(obj, pickle.HIGHEST_PROTOCOL))
        return Request()

    def send(self, obj, dest, tag=0):
        self.isend(obj, dest, tag)

    def recv(self, source=ANY_SOURCE, tag=ANY_TAG):
        return pickle.loads(_recv(source, tag))

    # ---- collectives (root = 0, deterministic order) ----------------------
    def _ctag(self):
        _state["coll"] += 1
        return _COLL_TAG0 + _state["coll"]

    def gather(self, obj, root=0):
        size, rank = _state["size"], _state["rank"]
        if size == 1:
            return [obj]
        tag = self._ctag()
        if rank == root:
            out = [None] * size
            out[root] = obj
            for _ in range(size - 1):
                src, o = pickle.loads(_recv(ANY_SOURCE, tag))
                out[src] = o
            return out
        _send(root, tag, pickle.dumps((rank, obj), pickle.HIGHEST_PROTOCOL))
        return None

    def bcast(self, obj, root=0):
        size, rank = _state["size"], _state["rank"]
        if size == 1:
            return obj
        tag = self._ctag()
        if rank == root:
            blob = pickle.dumps(obj, pickle.HIGHEST_PROTOCOL)
            for r in range(size):
                if r != root:
                    _send(r, tag, blob)
            return obj
        return pickle.loads(_recv(root, tag))

    def scatter(self, objs, root=0):
        size, rank = _state["size"], _state["rank"]
        if size == 1:
            return objs[0]
        tag = self._ctag()
        if rank == root:
            for r in range(size):
                if r != root:
                    _send(r, tag, pickle.dumps(objs[r], pickle.HIGHEST_PROTOCOL))
            return objs[root]
        return pickle.loads(_recv(root, tag))

    def allreduce(self, obj, op=SUM):
        if _state["size"] == 1:
            return obj
        parts = self.gather(obj, root=0)
        tot = None
        if _state["rank"] == 0:
            tot = parts[0]
            for p in parts[1:]:
                tot = tot + p
        return self.bcast(tot, root=0)

    def Allgather(self, sendbuf, recvbuf):
        size, rank = _state["size"], _state["rank"]
        send = np.asarray(sendbuf)
        if size == 1:
            recvbuf[...] = send.reshape(recvbuf.shape)
            return
        parts = self.gather(send, root=0)
        full = self.bcast(np.concatenate([np.asarray(p).ravel() for p in parts]) if rank == 0 else None, root=0)
        recvbuf[...] = full.reshape(recvbuf.shape)


COMM_WORLD = _WorldComm()
COMM_SELF = _SelfComm()


class Win:
    """Shared window: with one rank per node-group a private bytearray is enough.

    Multi-rank runs of the *builder* are not needed (the reference builder handles
    all parts on one rank through its MPGSize path, partition_mesh.py:113-116).
    """

    def __init__(self, nbytes, itemsize):
        self._buf = bytearray(int(nbytes))
        self._itemsize = itemsize

    @classmethod
    def Allocate_shared(cls, nbytes, itemsize, comm=None):
        if comm is not None and comm.Get_size() > 1:
            raise NotImplementedError("fake MPI: shared windows only for single-rank runs")
        return cls(nbytes, itemsize)

    def Shared_query(self, rank):
        return self._buf, self._itemsize


class File:
    def __init__(self, fd):
        self._fd = fd

    @classmethod
    def Open(cls, comm, filename, amode=MODE_RDONLY):
        flags = 0
        if amode & MODE_CREATE:
            flags |= os.O_CREAT
        if amode & MODE_RDWR:
            flags |= os.O_RDWR
        elif amode & MODE_WRONLY:
            flags |= os.O_WRONLY
        else:
            flags |= os.O_RDONLY
        return cls(os.open(filename, flags, 0o644))

    def Write(self, buf):
        os.write(self._fd, np.ascontiguousarray(buf).tobytes())

    def Write_at(self, offset, buf):
        os.pwrite(self._fd, np.ascontiguousarray(buf).tobytes(), int(offset))

    def Read(self, buf):
        raw = os.read(self._fd, buf.nbytes)
        buf[...] = np.frombuffer(raw, dtype=buf.dtype).reshape(buf.shape)

    def Read_at(self, offset, buf):
        raw = os.pread(self._fd, buf.nbytes, int(offset))
        buf[...] = np.frombuffer(raw, dtype=buf.dtype).reshape(buf.shape)

    def Close(self):
        os.close(self._fd)
