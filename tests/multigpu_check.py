#!/usr/bin/env python
"""Multi-GPU parity check, one process per GPU (launched by tests/test_gpu_multi.py, or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multigpu_check.py

Every rank steps the same global ensemble (row-block sharded) in both exchange
modes and compares the result with the single-process oracle: stretch
coordinates and accept counts bit-exact, log-probabilities to 1e-11 -- i.e. the
result does not depend on the number of GPUs.  Covers the dense_dmma kernel
(one launch per half-step with the fused peer barrier, locality-sorted and natural tile
order), tma_rows, the generic kernel, a mixed schedule that
alternates fused and unfused kernels, stored chains, sharded read-back and the
device-side chain moments."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import emcee_b200  # noqa: E402
from emcee_b200 import dist, models, moves  # noqa: E402
from oracle import redblue as rb  # noqa: E402
from oracle import targets as T  # noqa: E402

QUICK = os.environ.get("EB_MG_QUICK", "0") == "1"


def build(rdv, mode, name, N, D, dmoves, seed, **kw):
    target, p0 = T.make_config(name, N, D)
    model = {"gauss_dense": lambda: models.GaussianDense(target.icov), "rosenbrock": lambda: models.Rosenbrock(),
             "ring": lambda: models.Ring(), "gauss_iso": lambda: models.GaussianIso()}[name]()
    s = emcee_b200.EnsembleSampler(N, D, model, moves=dmoves, seed=seed, device=rdv.local_rank)
    s.attach(rdv, mode, **kw)
    return s, target, p0


def compare(last_coords, last_lp, o, omoves, rows=slice(None)):
    exact = all(m.kind == "stretch" for m, _ in omoves)
    snooker = any(m.kind == "snooker" for m, _ in omoves)
    if exact:
        assert np.array_equal(last_coords[rows], o.coords[rows]), "coords differ from the oracle"
    else:
        tol = 1e-6 if snooker else 1e-11
        np.testing.assert_allclose(last_coords[rows], o.coords[rows], rtol=tol, atol=tol)
    np.testing.assert_allclose(last_lp[rows], o.log_prob[rows], rtol=1e-6 if snooker else 1e-11,
                               atol=1e-4 if snooker else 1e-11)


def run_case(rdv, mode, name, N, D, omoves, dmoves, steps, seed=0xD157, group=1, local_first=2):
    s, target, p0 = build(rdv, mode, name, N, D, dmoves, seed)
    if group > 1:
        s._engine.set_option("dmma_group", group)
    s._engine.set_option("dmma_local_first", local_first)
    last = s.run_mcmc(p0, steps, store=False, skip_initial_state_check=True)
    nacc = s._engine.naccepted()
    o = rb.OracleSampler(N, D, target, omoves, seed=seed)
    o.set_state(p0)
    o.run(steps)
    compare(last.coords, last.log_prob, o, omoves)
    assert np.array_equal(nacc, o.naccepted.astype(np.uint64)), "accept counts differ"
    # every rank must hold the same replica
    digest = (float(last.coords.sum()), float(last.log_prob.sum()), int(nacc.sum()))
    all_digests = rdv.allgather(digest)
    assert all(d == all_digests[0] for d in all_digests), all_digests
    if rdv.rank == 0:
        print("PASS %-9s %-12s %6dx%-4d %3d steps  kernel=%s group=%d world=%d" % (
            mode, name, N, D, steps, s._engine.last_kernel_name(), group, rdv.world), flush=True)
    s._engine.close()
    rdv.barrier()


def run_store_case(rdv, mode, name, N, D, omoves, dmoves, steps, thin_by, seed=0x570E):
    """store=True on a sharded ensemble: every stored step must hold every walker's row."""
    s, target, p0 = build(rdv, mode, name, N, D, dmoves, seed)
    s.run_mcmc(p0, steps, thin_by=thin_by, skip_initial_state_check=True)
    chain, lps = s.get_chain(), s.get_log_prob()
    o = rb.OracleSampler(N, D, target, omoves, seed=seed)
    o.set_state(p0)
    acc_tot = np.zeros(N)
    for k in range(steps):
        for _ in range(thin_by):
            last_acc = o.run(1)
        acc_tot += last_acc
        compare(chain[k], lps[k], o, omoves)
    assert np.array_equal(s.backend.accepted, acc_tot), "stored accept totals differ"
    if rdv.rank == 0:
        print("PASS %-9s store=True %-12s %6dx%-4d %d stored steps (thin_by=%d)" % (mode, name, N, D, steps, thin_by),
              flush=True)
    s._engine.close()
    rdv.barrier()


def run_sharded_case(rdv, mode, N, D, steps, seed=0x5AAD):
    """gather_results=False: only the owned block comes back; device-side chain moments combine across ranks."""
    omoves, dmoves = [(rb.Stretch(), 1.0)], moves.StretchMove()
    s, target, p0 = build(rdv, mode, "gauss_dense", N, D, dmoves, seed, gather_results=False)
    s.enable_moments(1)
    last = s.run_mcmc(p0, steps, store=False, skip_initial_state_check=True)
    o = rb.OracleSampler(N, D, target, omoves, seed=seed)
    o.set_state(p0)
    flat = []
    for _ in range(steps):
        o.run(1)
        flat.append(o.coords.copy())
    compare(last.coords, last.log_prob, o, omoves, rows=s.owned_rows)
    flat = np.concatenate(flat)
    mean, cov, n = s.moments()
    assert n == flat.shape[0]
    np.testing.assert_allclose(mean, flat.mean(0), rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(cov, np.cov(flat, rowvar=False), rtol=1e-9, atol=1e-12)
    if rdv.rank == 0:
        print("PASS %-9s sharded read-back + moments %6dx%-4d %d steps" % (mode, N, D, steps), flush=True)
    s._engine.close()
    rdv.barrier()


def main():
    rdv = dist.Rendezvous()
    S, DE, SN = rb.Stretch, rb.DE, rb.Snooker
    for mode in ("p2p", "allgather"):
        run_case(rdv, mode, "gauss_dense", 4096, 128, [(S(), 1.0)], moves.StretchMove(), 24)
        run_case(rdv, mode, "gauss_dense", 1024, 24, [(S(nsplits=3), 1.0)], moves.StretchMove(nsplits=3), 16)
        run_case(rdv, mode, "rosenbrock", 2048, 64, [(DE(), 0.7), (SN(), 0.3)],
                 [(moves.DEMove(), 0.7), (moves.DESnookerMove(), 0.3)], 16)
        run_case(rdv, mode, "ring", 8192, 32, [(S(randomize_split=False), 1.0)],
                 moves.StretchMove(randomize_split=False), 16)
        # fused dense_dmma launches alternating with unfused generic ones (DE on the dense Gaussian)
        run_case(rdv, mode, "gauss_dense", 2048, 64, [(S(), 0.5), (DE(), 0.5)],
                 [(moves.StretchMove(), 0.5), (moves.DEMove(), 0.5)], 24)
        run_store_case(rdv, mode, "gauss_dense", 1024, 32, [(S(), 1.0)], moves.StretchMove(), 6, 2)
        run_store_case(rdv, mode, "gauss_iso", 512, 8, [(S(), 0.6), (DE(), 0.4)],
                       [(moves.StretchMove(), 0.6), (moves.DEMove(), 0.4)], 5, 1)
        run_sharded_case(rdv, mode, 2048, 32, 6)
    # the option is accepted but sharded ensembles always run one half-step per launch
    run_case(rdv, "p2p", "gauss_dense", 1024, 24, [(S(nsplits=3), 1.0)], moves.StretchMove(nsplits=3), 16, group=5)
    # natural tile order (peer barrier at the head of every half-step) as well as the locality-sorted default
    run_case(rdv, "p2p", "gauss_dense", 4096, 128, [(S(), 1.0)], moves.StretchMove(), 24, local_first=0)
    if not QUICK:
        # >= 4 tiles per consumer warp on every rank: the steady state of the landing pipeline
        n_big = 8 * 148 * 8 * 4 * 2 * rdv.world
        run_case(rdv, "p2p", "gauss_dense", n_big, 128, [(S(), 1.0)], moves.StretchMove(), 3)
        run_case(rdv, "p2p", "gauss_dense", n_big, 128, [(S(), 1.0)], moves.StretchMove(), 3, local_first=0)
    if rdv.rank == 0:
        print("ALL MULTI-GPU CHECKS PASSED (world=%d)" % rdv.world, flush=True)
    rdv.close()


if __name__ == "__main__":
    main()
