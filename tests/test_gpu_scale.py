"""Parity at BASELINE.json's full sizes.  At these sizes every persistent CTA / warp works through several
tiles, so the multi-tile pipelines of dense_dmma (landing slots, meta double-buffer, mbarrier phase flips)
and tma_rows (two-stage ring, bulk stores) are exercised, which the small cases cannot do.

Two kinds of checks: a few steps against the oracle (exact for the stretch move), and size-independent
invariants after a longer run -- the stored log-probability of every walker equals a fresh evaluation of its
stored coordinates, walkers whose proposal was rejected in the last step did not move, accepted ones did."""
import numpy as np
import pytest

from oracle import redblue as rb
from oracle import targets as T

from gpu_util import device_model, device_moves, move_rows_from_oracle

import emcee_b200

pytestmark = pytest.mark.gpu

FULL = [
    # BASELINE.json configs[2] shape (one GPU), configs[4], configs[3]
    ("gauss_dense", 65536, 128, [(rb.Stretch(), 1.0)], 3, "dense_dmma"),
    ("gauss_dense", 40008, 128, [(rb.Stretch(nsplits=3), 1.0)], 2, "dense_dmma"),
    ("ring", 262144, 32, [(rb.Stretch(), 1.0)], 2, "tma_rows"),
    ("gauss_iso", 65536, 128, [(rb.Stretch(a=1.7), 1.0)], 2, "tma_rows"),
    ("rosenbrock", 16384, 256, [(rb.DE(), 0.8), (rb.Snooker(), 0.2)], 3, "tma_rows"),
]


def _sampler(name, N, D, omoves, seed, target):
    return emcee_b200.EnsembleSampler(
        N, D, device_model(name, target=target), moves=device_moves(move_rows_from_oracle(omoves)), seed=seed
    )


@pytest.mark.parametrize("name,N,D,omoves,nsteps,kernel", FULL, ids=[f"{c[0]}-{c[1]}x{c[2]}" for c in FULL])
def test_full_size_against_oracle(name, N, D, omoves, nsteps, kernel):
    target, p0 = T.make_config(name, N, D)
    seed = 0xF011 + N
    o = rb.OracleSampler(N, D, target, omoves, seed=seed)
    o.set_state(p0)
    o.run(nsteps)
    s = _sampler(name, N, D, omoves, seed, target)
    last = s.run_mcmc(p0, nsteps, store=False, skip_initial_state_check=True)
    assert s._engine.last_kernel_name() == kernel
    assert np.array_equal(s._engine.naccepted(), o.naccepted.astype(np.uint64))
    if all(m.kind == "stretch" for m, _ in omoves):
        assert np.array_equal(last.coords, o.coords)
    else:
        np.testing.assert_allclose(last.coords, o.coords, rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(last.log_prob, o.log_prob, rtol=1e-9, atol=1e-6)


@pytest.mark.parametrize("name,N,D,omoves,nsteps,kernel", FULL, ids=[f"{c[0]}-{c[1]}x{c[2]}" for c in FULL])
def test_full_size_invariants(name, N, D, omoves, nsteps, kernel):
    target, p0 = T.make_config(name, N, D)
    s = _sampler(name, N, D, omoves, 0x1234 + D, target)
    eng, sched = s._engine, s._schedule()
    eng.set_state(p0)
    eng.step(sched, 25, want_accepted=False)
    before, _ = eng.get_state()
    acc = eng.step(sched, 1)
    after, lp = eng.get_state()
    moved = np.any(after != before, axis=1)
    assert np.array_equal(moved, acc)  # accepted <=> the row changed (proposals never equal the old row)
    assert 0.001 < acc.mean() < 0.999
    assert np.all(np.isfinite(lp)) and np.all(np.isfinite(after))
    fresh = eng.compute_log_prob(after)
    if name == "gauss_dense":
        assert np.array_equal(fresh, lp)  # same tensor-pipe arithmetic in both kernels
    else:
        np.testing.assert_allclose(fresh, lp, rtol=1e-13, atol=1e-13)
    # counters: every walker proposed once per step
    n = eng.naccepted()
    assert n.max() <= 26 and n.sum() > 0
