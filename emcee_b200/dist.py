"""Multi-GPU host plumbing: one process per GPU, walkers sharded by row block.

Rank ``r`` of ``R`` owns walkers ``[r*N/R, (r+1)*N/R)``.  In each split it
updates the active walkers it owns; because the active set is listed in
ascending walker order (``moves/red_blue.py:85``), the owned ones are one
contiguous range ``[i_lo, i_hi)`` of *active ranks* -- the indices the
counter-based draws are keyed by, so results do not depend on ``R``.  After
every split the updated rows are exchanged (``ncclAllGather`` of the owned row
blocks, or -- ``mode="p2p"`` -- nothing at all: partner rows are read from the
owner's HBM over NVLink inside the kernel, with a peer-memory flag barrier).

``torch.distributed`` (gloo) is used for the host-side rendezvous only
(broadcast of the NCCL id, all-gather of the IPC handles, barriers); the data
path is NCCL / NVLink peer memory inside ``libemcee_b200.so``.
"""

import os

import numpy as np

from . import _lib

__all__ = ["row_block", "active_range", "Rendezvous", "attach"]


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


class Rendezvous(object):
    """Thin wrapper over an initialised ``torch.distributed`` process group (or a
    single process when ``WORLD_SIZE`` is unset)."""

    def __init__(self, backend="gloo"):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", str(self.rank)))
        self.td = None
        if self.world > 1:
            import torch.distributed as td

            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if not td.is_initialized():
                td.init_process_group(backend, rank=self.rank, world_size=self.world)
            self.td = td

    def barrier(self):
        if self.td:
            self.td.barrier()

    def bcast(self, obj, src=0):
        if not self.td:
            return obj
        box = [obj]
        self.td.broadcast_object_list(box, src=src)
        return box[0]

    def allgather(self, obj):
        if not self.td:
            return [obj]
        out = [None] * self.world
        self.td.all_gather_object(out, obj)
        return out

    def max(self, x):
        if not self.td:
            return x
        import torch

        t = torch.tensor([float(x)], dtype=torch.float64)
        self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return float(t[0])

    def close(self):
        if self.td and self.td.is_initialized():
            self.td.destroy_process_group()


def attach(engine, rdv, mode="allgather"):
    """Join ``engine`` (an ``_lib.Engine`` holding the GLOBAL ensemble size) to
    the communicator of ``rdv``."""
    if rdv.world == 1:
        return
    m = {"allgather": _lib.EB_COMM_ALLGATHER, "p2p": _lib.EB_COMM_P2P}[mode]
    cid = rdv.bcast(_lib.Engine.comm_id() if rdv.rank == 0 else None)
    engine.comm_init(cid, rdv.rank, rdv.world, m)
    if m == _lib.EB_COMM_P2P:
        blobs = rdv.allgather(engine.comm_export())
        engine.comm_import(b"".join(blobs))
    rdv.barrier()
