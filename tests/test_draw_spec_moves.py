"""CPU checks of the round-2 additions to the draw specification (oracle/philox.py) and of the host helpers
that combine per-rank chain moments -- no GPU needed."""
import numpy as np
import pytest

from oracle import philox as px
from oracle import redblue as rb


def test_normals_are_the_two_box_muller_branches_of_one_block():
    seed, step, split = 0xABCDEF, 7, 3
    idx = np.arange(5)
    z = px.normals(seed, step, split, idx, 7)  # odd count: the last block contributes its cosine branch only
    assert z.shape == (5, 7)
    for k in range(4):
        w0, w1, w2, w3 = px.draw_words(seed, step, px.sub_split(split, k), px.TAG_NORMAL, idx)
        r = np.sqrt(-2.0 * np.log(1.0 - px.u53(w0, w1)))
        t = 6.283185307179586 * px.u53(w2, w3)
        assert np.array_equal(z[:, 2 * k], r * np.cos(t))
        if 2 * k + 1 < 7:
            assert np.array_equal(z[:, 2 * k + 1], r * np.sin(t))
    big = px.normals(seed, step, 0, np.arange(200000), 2)
    assert abs(big.mean()) < 0.01 and abs(big.std() - 1.0) < 0.01 and abs(np.corrcoef(big.T)[0, 1]) < 0.01
    assert px.sub_split(5, 0) == 5 and px.sub_split(5, 3) == 5 | (3 << 6)


def test_subset_indices_are_prefixes_of_a_bijection():
    seed, step, split = 99, 4, 1
    for n in (7, 32, 1000):
        full = px.subset_indices(seed, step, split, 11, n, n - 1)  # n - 1 images of the keyed permutation
        assert len(set(full.tolist())) == n - 1 and full.min() >= 0 and full.max() < n
        for s in (2, 3, n // 2):
            assert np.array_equal(px.subset_indices(seed, step, split, 11, n, s), full[:s])  # prefix property
        assert np.array_equal(px.subset_indices(seed, step, split, 11, n, n), np.arange(n))  # whole set: identity
        other = px.subset_indices(seed, step, split, 12, n, n - 1)
        assert not np.array_equal(other, full)  # keyed by the active rank


def test_chol_psd_factorises_definite_and_semidefinite_matrices():
    rng = np.random.default_rng(3)
    a = rng.standard_normal((40, 6))
    cov = np.cov(a, rowvar=0)
    np.testing.assert_allclose(px.chol_psd(cov), np.linalg.cholesky(cov), rtol=1e-12, atol=1e-14)
    # covariance of 3 rows in 5-D: rank 2 -- two pivots, the rest of the factor is exactly zero
    x = rng.standard_normal((3, 5))
    c3 = np.cov(x, rowvar=0)
    L = px.chol_psd(c3, max_rank=2)
    assert np.count_nonzero(np.diag(L)) == 2 and np.all(np.triu(L, 1) == 0)
    np.testing.assert_allclose(L @ L.T, c3, rtol=0, atol=1e-12 * np.abs(c3).max() * 10)
    # without the rank bound a third "pivot" may be rounding noise; with it the factor is stable under tiny perturbations
    L2 = px.chol_psd(c3 * (1 + 1e-15), max_rank=2)
    np.testing.assert_allclose(L2, L, rtol=1e-9, atol=1e-12)
    assert np.all(px.chol_psd(np.zeros((3, 3))) == 0)


def test_gaussian_oracle_argument_checks_match_the_reference():
    # gaussian.py:36-79
    with pytest.raises(ValueError):
        rb.Gaussian(np.zeros((2, 3)))
    with pytest.raises(ValueError):
        rb.Gaussian(np.eye(3), mode="random")
    with pytest.raises(ValueError):
        rb.Gaussian(1.0, factor=0.5)
    g = rb.Gaussian(np.array([4.0, 9.0]))
    assert g.form == "diag" and np.array_equal(g.scale, [2.0, 3.0])
    assert rb.Gaussian(4.0).scale == 2.0 and rb.Gaussian(np.eye(2)).form == "full"


def test_combine_moments_matches_numpy_on_sharded_samples():
    from emcee_b200 import dist

    rng = np.random.default_rng(8)
    x = rng.standard_normal((1000, 4)) @ rng.standard_normal((4, 4)) + 5.0
    parts = []
    for lo, hi in ((0, 100), (100, 100), (100, 640), (640, 1000)):  # one empty shard
        blk = x[lo:hi]
        if len(blk) == 0:
            parts.append((np.full(4, np.nan), np.full((4, 4), np.nan), 0, 3))
        else:
            parts.append((blk.mean(0), np.cov(blk, rowvar=False), len(blk), 7))
    mean, cov, n, nacc = dist.combine_moments(parts)
    assert n == 1000 and nacc == 3 + 7 * 3
    np.testing.assert_allclose(mean, x.mean(0), rtol=1e-13)
    np.testing.assert_allclose(cov, np.cov(x, rowvar=False), rtol=1e-12)
    assert dist.combine_moments([(None, None, 0, 0)])[2] == 0
