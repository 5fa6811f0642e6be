"""Multi-GPU host plumbing: one process per GPU, walkers sharded by row block.

Rank ``r`` of ``R`` owns walkers ``[r*N/R, (r+1)*N/R)``.  In each split it
updates the active walkers it owns; because the active set is listed in
ascending walker order (``moves/red_blue.py:85``), the owned ones are one
contiguous range ``[i_lo, i_hi)`` of *active ranks* -- the indices the
counter-based draws are keyed by, so results do not depend on ``R``.  After
every split the updated rows are exchanged (``ncclAllGather`` of the owned row
blocks, or -- ``mode="p2p"`` -- nothing at all: partner rows are read from the
owner's HBM over NVLink inside the kernel, with a peer-memory flag barrier).

The host-side rendezvous (broadcast of the NCCL id, all-gather of the IPC
handles, barriers, max-over-ranks of timings) is a ~100-line TCP star over
the loopback interface, driven by the environment variables ``torchrun`` (or
any launcher) sets: ``RANK``, ``WORLD_SIZE``, ``LOCAL_RANK``, ``MASTER_PORT``.
No PyTorch: the data path is NCCL / NVLink peer memory inside
``libemcee_b200.so``, and this module moves a few hundred bytes.
"""

import os
import pickle
import socket
import struct
import time

import numpy as np

from . import _lib

__all__ = ["row_block", "active_range", "Rendezvous", "attach", "combine_moments"]


def row_block(nwalkers, rank, nranks):
    """``[w_lo, w_hi)`` owned by ``rank``."""
    if nwalkers % nranks:
        raise ValueError("nwalkers must be divisible by the number of ranks")
    rows = nwalkers // nranks
    return rank * rows, (rank + 1) * rows


def active_range(active_walkers, w_lo, w_hi):
    """``[i_lo, i_hi)``: positions in the ascending ``active_walkers`` list whose
    walker id lies in ``[w_lo, w_hi)`` (what ``split_table_kernel`` computes on
    the device for every (step, set))."""
    a = np.asarray(active_walkers)
    return int(np.searchsorted(a, w_lo, side="left")), int(np.searchsorted(a, w_hi, side="left"))


def _send(sock, obj):
    blob = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    sock.sendall(struct.pack("<Q", len(blob)) + blob)


def _recv(sock):
    def exact(n):
        buf = bytearray()
        while len(buf) < n:
            part = sock.recv(n - len(buf))
            if not part:
                raise ConnectionError("rendezvous peer closed the connection")
            buf += part
        return bytes(buf)

    (n,) = struct.unpack("<Q", exact(8))
    return pickle.loads(exact(n))


class Rendezvous(object):
    """Host-side collectives between the ranks of one node (a single process
    when ``WORLD_SIZE`` is unset).  Rank 0 listens on an ephemeral loopback
    port and publishes it in a small file named after the launcher's
    ``MASTER_PORT`` and parent pid (``EB_RDV_FILE`` overrides the path); the
    other ranks connect to it.  Every collective is a gather to rank 0
    followed by a broadcast of the gathered list."""

    def __init__(self, backend=None, timeout=120.0):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", str(self.rank)))
        self._peers = []  # rank 0: sockets of ranks 1..world-1, in rank order
        self._sock = None  # other ranks: socket to rank 0
        self._server = None
        if self.world == 1:
            return
        path = os.environ.get("EB_RDV_FILE") or os.path.join(
            os.environ.get("TMPDIR", "/tmp"),
            "emcee_b200_rdv_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid()),
        )
        self._path = path
        deadline = time.time() + timeout
        if self.rank == 0:
            try:
                os.unlink(path)
            except OSError:
                pass
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind(("127.0.0.1", 0))
            srv.listen(self.world)
            srv.settimeout(timeout)
            self._server = srv
            tmp = "%s.%d" % (path, os.getpid())
            with open(tmp, "w") as f:
                f.write("%d %d\n" % (srv.getsockname()[1], os.getpid()))
            os.replace(tmp, path)
            byrank = {}
            while len(byrank) < self.world - 1:
                conn, _ = srv.accept()
                conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                conn.settimeout(timeout)
                byrank[_recv(conn)] = conn
            self._peers = [byrank[r] for r in range(1, self.world)]
        else:
            while True:
                try:
                    port = int(open(path).read().split()[0])
                    s = socket.create_connection(("127.0.0.1", port), timeout=5.0)
                    break
                except (OSError, ValueError, IndexError):
                    if time.time() > deadline:
                        raise RuntimeError("rendezvous: rank 0 did not appear at %s" % path)
                    time.sleep(0.05)
            s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            s.settimeout(timeout)
            _send(s, self.rank)
            self._sock = s

    # -- collectives ---------------------------------------------------------
    def allgather(self, obj):
        if self.world == 1:
            return [obj]
        if self.rank == 0:
            out = [obj] + [_recv(p) for p in self._peers]
            for p in self._peers:
                _send(p, out)
            return out
        _send(self._sock, obj)
        return _recv(self._sock)

    def barrier(self):
        self.allgather(None)

    def bcast(self, obj, src=0):
        return self.allgather(obj if self.rank == src else None)[src]

    def max(self, x):
        return float(max(self.allgather(float(x))))

    def close(self):
        if self.world > 1:
            try:
                self.barrier()
            except Exception:
                pass
        for p in self._peers:
            p.close()
        if self._sock is not None:
            self._sock.close()
        if self._server is not None:
            self._server.close()
            try:
                os.unlink(self._path)
            except OSError:
                pass
        self._peers, self._sock, self._server = [], None, None


def attach(engine, rdv, mode="allgather"):
    """Join ``engine`` (an ``_lib.Engine`` holding the GLOBAL ensemble size) to
    the communicator of ``rdv``.  Call before the state is set: the engine then
    uploads / evaluates only the rows it owns."""
    if rdv.world == 1:
        return
    m = {"allgather": _lib.EB_COMM_ALLGATHER, "p2p": _lib.EB_COMM_P2P}[mode]
    cid = rdv.bcast(_lib.Engine.comm_id() if rdv.rank == 0 else None)
    engine.comm_init(cid, rdv.rank, rdv.world, m)
    if m == _lib.EB_COMM_P2P:
        blobs = rdv.allgather(engine.comm_export())
        engine.comm_import(b"".join(blobs))
    rdv.barrier()


def combine_moments(parts):
    """Merge per-rank ``(mean, cov, count, naccepted)`` tuples (``Engine.moments()``)
    into the moments of the whole ensemble (pairwise update of Chan et al.)."""
    mean, m2, n, na = None, None, 0, 0
    for mu, cov, cnt, acc in parts:
        na += acc
        if cnt == 0:
            continue
        s2 = np.asarray(cov) * (cnt - 1)
        if n == 0:
            mean, m2, n = np.array(mu, dtype=np.float64), s2, cnt
            continue
        delta = np.asarray(mu) - mean
        tot = n + cnt
        m2 = m2 + s2 + np.outer(delta, delta) * (n * cnt / tot)
        mean = mean + delta * (cnt / tot)
        n = tot
    if n == 0:
        return None, None, 0, na
    return mean, m2 / (n - 1), n, na
