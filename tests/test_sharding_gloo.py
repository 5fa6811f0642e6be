"""The multi-GPU scheme on CPU: world_size-2 gloo processes emulate two ranks with
the oracle (row-block ownership, update only owned active walkers, exchange the
owned row blocks after every split) and must reproduce the single-process oracle
bit for bit -- the property the CUDA path relies on (draws keyed by active rank,
results independent of the number of GPUs)."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _moves(rb, spec):
    return {
        "stretch": lambda: [(rb.Stretch(), 1.0)],
        "de+snooker": lambda: [(rb.DE(), 0.6), (rb.Snooker(), 0.4)],
        "stretch3fixed": lambda: [(rb.Stretch(nsplits=3, randomize_split=False), 1.0)],
    }[spec]()


def _worker(rank, world, port, cases, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      EB_RDV_FILE=out % (99, 99) + ".rdv")
    import torch
    import torch.distributed as td

    from emcee_b200 import dist
    from oracle import redblue as rb
    from oracle import targets as T

    td.init_process_group("gloo", rank=rank, world_size=world)  # the emulated data-path collective
    # the product's own host rendezvous (TCP star, no torch): exercise its collectives as well
    rdv = dist.Rendezvous()
    assert rdv.allgather(("r", rank)) == [("r", r) for r in range(world)]
    assert rdv.bcast(b"id-%d" % rank, src=0) == b"id-0"
    assert rdv.max(10.0 + rank) == 10.0 + world - 1
    rdv.barrier()
    for ci, (name, N, D, moves, steps) in enumerate(cases):
        target, p0 = T.make_config(name, N, D)
        o = rb.OracleSampler(N, D, target, _moves(rb, moves), seed=77)
        o.set_state(p0)
        w_lo, w_hi = dist.row_block(N, rank, world)
        o.owner_range = (w_lo, w_hi)

        def exchange(coords, log_prob, accepted, N=N, w_lo=w_lo, w_hi=w_hi):
            # what ncclAllGather does in place on the device: every rank contributes its row block
            for arr in (coords, log_prob, accepted):
                t = torch.from_numpy(np.ascontiguousarray(arr[w_lo:w_hi]))
                parts = [torch.empty_like(t) for _ in range(world)]
                td.all_gather(parts, t)
                for r, p in enumerate(parts):
                    lo, hi = dist.row_block(N, r, world)
                    arr[lo:hi] = p.numpy()

        o.exchange = exchange
        o.run(steps)
        np.savez(out % (ci, rank), coords=o.coords, log_prob=o.log_prob)
    rdv.close()
    td.destroy_process_group()


CASES = [
    ("gauss_dense", 64, 8, "stretch", 12),
    ("rosenbrock", 48, 4, "de+snooker", 12),
    ("gauss_iso", 30, 3, "stretch3fixed", 8),
]


def test_two_ranks_reproduce_one(tmp_path):
    """One pair of gloo processes runs every case (importing torch in a fresh
    process dominates the cost)."""
    import torch.multiprocessing as mp

    from oracle import redblue as rb
    from oracle import targets as T

    out = str(tmp_path / "case%d_rank%d.npz")
    mp.spawn(_worker, args=(2, _free_port(), CASES, out), nprocs=2, join=True)
    for ci, (name, N, D, moves, steps) in enumerate(CASES):
        target, p0 = T.make_config(name, N, D)
        ref = rb.OracleSampler(N, D, target, _moves(rb, moves), seed=77)
        ref.set_state(p0)
        ref.run(steps)
        for rank in range(2):
            got = np.load(out % (ci, rank))
            assert np.array_equal(got["coords"], ref.coords), (name, rank)
            assert np.array_equal(got["log_prob"], ref.log_prob), (name, rank)


def test_active_range_matches_definition():
    from emcee_b200 import dist
    from oracle import philox as px

    N, P = 96, 4
    inds = px.split_assignment(5, 3, N, P, True)
    for world in (1, 2, 3, 4, 8):
        for split in range(P):
            act = np.flatnonzero(inds == split)
            covered = []
            for r in range(world):
                lo, hi = dist.row_block(N, r, world)
                i_lo, i_hi = dist.active_range(act, lo, hi)
                assert np.all((act[i_lo:i_hi] >= lo) & (act[i_lo:i_hi] < hi))
                covered.extend(range(i_lo, i_hi))
            assert covered == list(range(len(act)))
