"""The C++ statement of the draw specification (emcee_b200/csrc/philox.cuh,
compiled for the host) against the numpy statement (oracle/philox.py) and the
reference's own pair table (moves/de.py:67-77 semantics, rebuilt with numpy)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import philox as px
from oracle import redblue as rb

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("probe") / "libphilox_probe.so")
    subprocess.run(
        ["g++", "-O2", "-shared", "-fPIC", "-o", out, os.path.join(HERE, "helpers", "philox_host.cpp")], check=True
    )
    return C.CDLL(out)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def test_draw_words_u53_bounded(probe):
    rng = np.random.default_rng(5)
    idx = rng.integers(0, 2**32, size=257, dtype=np.uint64).astype(np.uint32)
    for seed, step, split, tag in [(0, 0, 0, 1), (0x656D636565B200, 12345678901, 3, 5), (2**64 - 1, 2**40 + 7, 31, 3)]:
        out = np.zeros((len(idx), 4), dtype=np.uint32)
        probe.probe_draw_words(C.c_uint64(seed), C.c_uint64(step), C.c_uint32(split), C.c_uint32(tag),
                               _p(idx, C.c_uint32), len(idx), _p(out, C.c_uint32))
        ref = np.stack(px.draw_words(seed, step, split, tag, idx), axis=1)
        assert np.array_equal(out, ref)
        lo, hi = np.ascontiguousarray(out[:, 0]), np.ascontiguousarray(out[:, 1])
        u = np.zeros(len(idx))
        probe.probe_u53(_p(lo, C.c_uint32), _p(hi, C.c_uint32), len(idx), _p(u, C.c_double))
        assert np.array_equal(u, px.u53(lo, hi)) and np.all((u >= 0) & (u < 1))
        for bound in (1, 6, 32768, 131072 * 131071, 2**40 + 3):
            b = np.zeros(len(idx), dtype=np.uint64)
            probe.probe_bounded64(_p(lo, C.c_uint32), _p(hi, C.c_uint32), len(idx), C.c_uint64(bound), _p(b, C.c_uint64))
            assert np.array_equal(b.astype(np.int64), px.bounded64(lo, hi, bound))
            assert b.max() < bound


def test_split_permutation(probe):
    for n in (2, 5, 32, 37, 1000, 4096, 65536, 100003):
        for step in (0, 7):
            out = np.zeros(n, dtype=np.int64)
            probe.probe_split_permutation(C.c_uint64(99), C.c_uint64(step), C.c_uint64(n), _p(out, C.c_int64))
            assert np.array_equal(out, px.split_permutation(99, step, n))


def test_de_pair_decode(probe):
    def table(n):  # what moves/de.py:67-77 builds
        rows, cols = np.tril_indices(n, -1)
        return np.column_stack([np.concatenate([rows, cols]), np.concatenate([cols, rows])])

    for n in (2, 3, 4, 17, 64):
        t = table(n)
        m = np.arange(len(t), dtype=np.uint64)
        p0 = np.zeros(len(t), dtype=np.uint64)
        p1 = np.zeros(len(t), dtype=np.uint64)
        probe.probe_de_pair(_p(m, C.c_uint64), len(t), C.c_uint64(n), _p(p0, C.c_uint64), _p(p1, C.c_uint64))
        assert np.array_equal(np.stack([p0, p1], 1).astype(np.int64), t)
        q0, q1 = rb.de_pair_decode(m.astype(np.int64), n)
        assert np.array_equal(np.stack([q0, q1], 1), t)
    # large n: the two decoders agree, stay in range and never return a diagonal pair
    n = 131072
    rng = np.random.default_rng(1)
    m = rng.integers(0, n * (n - 1), size=4096, dtype=np.uint64)
    m[:4] = [0, n * (n - 1) // 2 - 1, n * (n - 1) // 2, n * (n - 1) - 1]
    p0 = np.zeros(len(m), dtype=np.uint64)
    p1 = np.zeros(len(m), dtype=np.uint64)
    probe.probe_de_pair(_p(m, C.c_uint64), len(m), C.c_uint64(n), _p(p0, C.c_uint64), _p(p1, C.c_uint64))
    q0, q1 = rb.de_pair_decode(m.astype(np.int64), n)
    assert np.array_equal(p0.astype(np.int64), q0) and np.array_equal(p1.astype(np.int64), q1)
    assert np.all(p0 != p1) and p0.max() < n and p1.max() < n
