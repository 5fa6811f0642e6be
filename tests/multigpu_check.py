#!/usr/bin/env python
"""Multi-GPU parity check, launched one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multigpu_check.py

Every rank steps the same global ensemble (row-block sharded) in both exchange
modes and compares its final replica with the single-process oracle: stretch
coordinates and accept counts bit-exact, log-probabilities to 1e-11 -- i.e. the
result does not depend on the number of GPUs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import emcee_b200  # noqa: E402
from emcee_b200 import dist, models, moves  # noqa: E402
from oracle import redblue as rb  # noqa: E402
from oracle import targets as T  # noqa: E402


def run_case(rdv, mode, name, N, D, omoves, dmoves, steps, seed=0xD157):
    target, p0 = T.make_config(name, N, D)
    model = {"gauss_dense": lambda: models.GaussianDense(target.icov), "rosenbrock": lambda: models.Rosenbrock(),
             "ring": lambda: models.Ring()}[name]()
    s = emcee_b200.EnsembleSampler(N, D, model, moves=dmoves, seed=seed, device=rdv.local_rank)
    dist.attach(s._engine, rdv, mode)
    last = s.run_mcmc(p0, steps, store=False, skip_initial_state_check=True)
    nacc = s._engine.naccepted()
    o = rb.OracleSampler(N, D, target, omoves, seed=seed)
    o.set_state(p0)
    o.run(steps)
    exact = all(m.kind == "stretch" for m, _ in omoves)
    snooker = any(m.kind == "snooker" for m, _ in omoves)
    if exact:
        assert np.array_equal(last.coords, o.coords), "coords differ from the oracle"
    else:
        tol = 1e-6 if snooker else 1e-11
        np.testing.assert_allclose(last.coords, o.coords, rtol=tol, atol=tol)
    np.testing.assert_allclose(last.log_prob, o.log_prob, rtol=1e-6 if snooker else 1e-11, atol=1e-4 if snooker else 1e-11)
    assert np.array_equal(nacc, o.naccepted.astype(np.uint64)), "accept counts differ"
    # every rank must hold the same replica
    digest = (float(last.coords.sum()), float(last.log_prob.sum()), int(nacc.sum()))
    all_digests = rdv.allgather(digest)
    assert all(d == all_digests[0] for d in all_digests), all_digests
    if rdv.rank == 0:
        print("PASS %-9s %-12s %6dx%-4d %3d steps  kernel=%s  world=%d" % (
            mode, name, N, D, steps, s._engine.last_kernel_name(), rdv.world), flush=True)
    s._engine.close()
    rdv.barrier()


def main():
    rdv = dist.Rendezvous("gloo")
    for mode in ("allgather", "p2p"):
        run_case(rdv, mode, "gauss_dense", 4096, 128, [(rb.Stretch(), 1.0)], moves.StretchMove(), 24)
        run_case(rdv, mode, "gauss_dense", 1024, 24, [(rb.Stretch(nsplits=3), 1.0)], moves.StretchMove(nsplits=3), 16)
        run_case(rdv, mode, "rosenbrock", 2048, 64, [(rb.DE(), 0.7), (rb.Snooker(), 0.3)],
                 [(moves.DEMove(), 0.7), (moves.DESnookerMove(), 0.3)], 16)
        run_case(rdv, mode, "ring", 8192, 32, [(rb.Stretch(randomize_split=False), 1.0)],
                 moves.StretchMove(randomize_split=False), 16)
    if rdv.rank == 0:
        print("ALL MULTI-GPU CHECKS PASSED (world=%d)" % rdv.world, flush=True)
    rdv.close()


if __name__ == "__main__":
    main()
