"""Property tests of the draw specification (numpy statement): things every
consumer relies on, for arbitrary sizes -- bijectivity and balance of the split
permutation, range and reproducibility of the bounded integers, and the exact
reconstruction of DEMove's ordered-pair table."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import philox as px
from oracle import redblue as rb


@settings(max_examples=60, deadline=None)
@given(n=st.integers(2, 5000), nsplits=st.integers(2, 8), seed=st.integers(0, 2**64 - 1), step=st.integers(0, 2**40))
def test_split_assignment_is_a_balanced_partition(n, nsplits, seed, step):
    nsplits = min(nsplits, n)
    perm = px.split_permutation(seed, step, n)
    assert np.array_equal(np.sort(perm), np.arange(n))
    inds = px.split_assignment(seed, step, n, nsplits, True)
    want = np.bincount(np.arange(n) % nsplits, minlength=nsplits)
    assert np.array_equal(np.bincount(inds, minlength=nsplits), want)
    assert np.array_equal(inds, px.split_assignment(seed, step, n, nsplits, True))  # pure function of its key


@settings(max_examples=60, deadline=None)
@given(bound=st.integers(1, 2**62), seed=st.integers(0, 2**64 - 1), step=st.integers(0, 2**50), split=st.integers(0, 31))
def test_bounded_integers_in_range(bound, seed, step, split):
    w0, w1, w2, w3 = px.draw_words(seed, step, split, px.TAG_PROP_A, np.arange(64))
    for lo, hi in ((w0, w1), (w2, w3)):
        r = px.bounded64(lo, hi, bound)
        assert r.min() >= 0 and r.max() < bound
        # exact integer arithmetic check of the multiply-shift
        x = (hi.astype(object) << 32) | lo.astype(object)
        assert [int(v) for v in r] == [(int(xx) * bound) >> 64 for xx in x]
    u = px.u53(w0, w1)
    assert np.all((u >= 0) & (u < 1))


@settings(max_examples=40, deadline=None)
@given(n=st.integers(2, 90))
def test_de_pair_decode_matches_the_table(n):
    rows, cols = np.tril_indices(n, -1)  # moves/de.py:70-75
    table = np.column_stack([np.concatenate([rows, cols]), np.concatenate([cols, rows])])
    p0, p1 = rb.de_pair_decode(np.arange(len(table)), n)
    assert np.array_equal(np.stack([p0, p1], 1), table)


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2**63), weights=st.lists(st.floats(0.01, 10.0), min_size=1, max_size=6))
def test_move_choice_follows_the_cdf(seed, weights):
    w = np.asarray(weights) / np.sum(weights)
    picks = np.array([px.move_choice(seed, step, w) for step in range(400)])
    assert picks.min() >= 0 and picks.max() < len(w)
    if len(w) > 1:
        freq = np.bincount(picks, minlength=len(w)) / 400.0
        assert np.all(np.abs(freq - w) < 0.15)
