"""Host chain store: the Backend protocol of the reference (``backends/backend.py``;
reference tests ``tests/unit/test_backends.py``) -- no GPU needed."""
import numpy as np
import pytest

import emcee_b200
from emcee_b200 import Backend, State


def filled(nsteps=12, nwalkers=6, ndim=3, seed=0):
    rng = np.random.default_rng(seed)
    b = Backend()
    assert not b.initialized
    b.reset(nwalkers, ndim)
    b.grow(nsteps, None)
    states, accs = [], []
    for k in range(nsteps):
        st = State(rng.standard_normal((nwalkers, ndim)), log_prob=rng.standard_normal(nwalkers), random_state=("r", k))
        acc = rng.random(nwalkers) < 0.5
        b.save_step(st, acc)
        states.append(st)
        accs.append(acc)
    return b, states, accs


def test_uninitialised_and_empty_access():
    b = Backend()
    with pytest.raises(AttributeError):
        b.get_last_sample()
    b.reset(4, 2)
    assert b.shape == (4, 2) and b.iteration == 0
    with pytest.raises(AttributeError):
        b.get_chain()  # backend.py:43-48


def test_save_and_slicing():
    b, states, accs = filled()
    assert b.iteration == 12 and b.get_chain().shape == (12, 6, 3) and b.get_log_prob().shape == (12, 6)
    np.testing.assert_array_equal(b.get_chain()[5], states[5].coords)
    # backend.py:53  v[discard + thin - 1 : iteration : thin]
    full = np.stack([s.coords for s in states])
    np.testing.assert_array_equal(b.get_chain(discard=2, thin=3), full[2 + 3 - 1 : 12 : 3])
    np.testing.assert_array_equal(b.get_chain(flat=True), full.reshape(-1, 3))
    np.testing.assert_array_equal(b.get_log_prob(flat=True, thin=2), np.stack([s.log_prob for s in states])[1::2].reshape(-1))
    np.testing.assert_array_equal(b.accepted, np.sum(accs, axis=0).astype(float))
    assert b.get_blobs() is None and not b.has_blobs()
    last = b.get_last_sample()
    np.testing.assert_array_equal(last.coords, states[-1].coords)
    assert last.random_state == ("r", 11)


def test_grow_keeps_data_and_reuses_room():
    b, states, _ = filled(nsteps=4)
    before = b.get_chain().copy()
    b.grow(3, None)  # backend.py:172-176: only the missing room is added
    assert len(b.chain) == 7
    np.testing.assert_array_equal(b.get_chain(), before)
    b.grow(2, None)
    assert len(b.chain) == 7


def test_shape_errors():
    b, _, _ = filled(nsteps=2)
    b.grow(1, None)
    with pytest.raises(ValueError):
        b.save_step(State(np.zeros((5, 3)), log_prob=np.zeros(6)), np.zeros(6, dtype=bool))
    with pytest.raises(ValueError):
        b.save_step(State(np.zeros((6, 3)), log_prob=np.zeros(5)), np.zeros(6, dtype=bool))
    with pytest.raises(ValueError):
        b.save_step(State(np.zeros((6, 3)), log_prob=np.zeros(6)), np.zeros(5, dtype=bool))
    with pytest.raises(ValueError):
        b.save_step(State(np.zeros((6, 3)), log_prob=np.zeros(6), blobs=np.zeros(6)), np.zeros(6, dtype=bool))
    with pytest.raises(NotImplementedError):
        Backend(dtype=np.float32)


def test_autocorr_time_from_backend():
    rng = np.random.default_rng(3)
    n, w, d = 6000, 4, 2
    x = np.empty((n, w, d))
    x[0] = 0
    e = rng.random((n, w, d))
    for i in range(1, n):
        x[i] = 0.9 * x[i - 1] + e[i]
    b = Backend()
    b.reset(w, d)
    b.grow(n, None)
    for i in range(n):
        b.save_step(State(x[i], log_prob=np.zeros(w)), np.ones(w, dtype=bool))
    tau = b.get_autocorr_time(quiet=True)
    assert tau.shape == (d,) and np.all(np.abs(tau - 19) / 19 < 0.3)
    np.testing.assert_allclose(b.get_autocorr_time(thin=2, quiet=True), 2 * emcee_b200.autocorr.integrated_time(x[1::2], quiet=True))


def test_walkers_independent_contract():
    # reference: tests/unit/test_sampler.py:237-321
    rng = np.random.default_rng(0)
    good = rng.standard_normal((20, 4))
    assert emcee_b200.walkers_independent(good)
    assert not emcee_b200.walkers_independent(np.ones((20, 4)))  # zero spread
    assert not emcee_b200.walkers_independent(np.tile(good[:, :1], (1, 4)) * [1, 2, 3, 4])  # rank 1
    bad = good.copy()
    bad[0, 0] = np.nan
    assert not emcee_b200.walkers_independent(bad)
    assert emcee_b200.walkers_independent(good * 1e-300 + 1.0) in (True, False)  # no exception on tiny spreads
    assert emcee_b200.walkers_independent(1e8 + good)  # offsets do not matter (centred)
