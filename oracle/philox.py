"""Counter-based random draws for the oracle (TEST INFRASTRUCTURE ONLY).

The reference consumes one sequential MT19937 stream (``ensemble.py:166-167``),
whose length is data dependent, so it cannot be reproduced by parallel
hardware.  Parity is therefore defined on *counter-addressed* draws: every
random number the walker-update path needs is a pure function of
``(seed, step, split, index, purpose)`` through Philox4x32-10 (Salmon et al.,
SC'11; same constants as ``curand_philox4x32_x.h``).  This module is the
numpy statement of that draw specification; ``emcee_b200/csrc/philox.cuh`` is
the CUDA statement and ``emcee_b200/rng.py`` documents the layout for users.

Two things live here:

* the draw specification itself (``philox4x32_10``, ``draw_words``, ``u53``,
  ``bounded64``, ``split_permutation`` ...), used by the vectorised oracle
  ``oracle/redblue.py``;
* ``PhiloxRandom`` -- a duck-typed "numpy-compatible random number state"
  (``moves/red_blue.py:60``; proven duck-typed by ``tests/unit/test_stretch.py:24``)
  exposing exactly the methods the reference calls on ``model.random`` on this
  path: ``choice``, ``shuffle``, ``rand``, ``randint``,
  ``randn``, ``get_state``, ``set_state``.  Injected as ``sampler._random`` it
  makes the *unmodified reference code* run on the counter-addressed draws;
  that is how ``oracle/gen_golden.py`` produces the golden vectors.

Counter layout (all uint32):  ``ctr = (index, step_lo, step_hi, (split << 8) | tag)``,
``key = (seed_lo, seed_hi)``.
"""

import numpy as np

__all__ = [
    "philox4x32_10",
    "draw_words",
    "u53",
    "bounded64",
    "box_muller",
    "split_assignment",
    "split_permutation",
    "move_choice",
    "PhiloxRandom",
    "TAG_MOVE",
    "TAG_SHUFFLE",
    "TAG_PROP_A",
    "TAG_PROP_B",
    "TAG_ACCEPT",
    "TAG_NORMAL",
    "TAG_SUBSET",
    "sub_split",
    "normals",
    "subset_indices",
    "chol_psd",
]

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_M32 = np.uint64(0xFFFFFFFF)

# draw purposes ("tag" byte of counter word 3)
TAG_MOVE = 1  # which move of the mixture runs this step   (ensemble.py:406)
TAG_SHUFFLE = 2  # round keys of the split permutation       (red_blue.py:79-80)
TAG_PROP_A = 3  # first proposal draw block of active rank i
TAG_PROP_B = 4  # second proposal draw block of active rank i
TAG_ACCEPT = 5  # Metropolis uniform of active rank i        (red_blue.py:100)
TAG_NORMAL = 6  # bulk standard normals of row i: block k holds normals 2k, 2k+1 (walk.py:36, gaussian.py:97)
TAG_SUBSET = 7  # round keys of the helper-subset permutation of active rank i (walk.py:34)

FEISTEL_ROUNDS = 8


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32 with 10 rounds.  Counters are uint32 arrays (broadcastable),
    the key two Python ints.  Returns four uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(
        np.asarray(c0, dtype=np.uint64),
        np.asarray(c1, dtype=np.uint64),
        np.asarray(c2, dtype=np.uint64),
        np.asarray(c3, dtype=np.uint64),
    )
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for r in range(10):
        p0 = PHILOX_M0 * c0  # 32x32 -> 64, exact in uint64
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _M32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _M32
        c0, c1, c2, c3 = (
            hi1 ^ c1 ^ np.uint64(k0),
            lo1,
            hi0 ^ c3 ^ np.uint64(k1),
            lo0,
        )
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return (
        c0.astype(np.uint32),
        c1.astype(np.uint32),
        c2.astype(np.uint32),
        c3.astype(np.uint32),
    )


def draw_words(seed, step, split, tag, index):
    """The four 32-bit words of draw block ``(step, split, tag)`` for each
    element of ``index``."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    step = int(step) & 0xFFFFFFFFFFFFFFFF
    c3 = ((int(split) & 0xFFFFFF) << 8) | (int(tag) & 0xFF)
    return philox4x32_10(
        np.asarray(index, dtype=np.uint64) & _M32,
        step & 0xFFFFFFFF,
        step >> 32,
        c3,
        seed & 0xFFFFFFFF,
        seed >> 32,
    )


def u53(lo, hi):
    """Uniform double on [0, 1) from two 32-bit words: the top 53 bits of
    ``hi:lo`` times 2**-53 (numpy's ``random_sample`` convention)."""
    x = (np.asarray(hi, dtype=np.uint64) << np.uint64(32)) | np.asarray(
        lo, dtype=np.uint64
    )
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def bounded64(lo, hi, n):
    """Integer on [0, n): the high 64 bits of ``(hi:lo) * n`` (multiply-shift;
    bias <= n / 2**64, i.e. < 2e-9 relative even for n = 2**35)."""
    n = int(n)
    assert 0 < n < (1 << 63)
    xl = np.asarray(lo, dtype=np.uint64)
    xh = np.asarray(hi, dtype=np.uint64)
    nl = np.uint64(n & 0xFFFFFFFF)
    nh = np.uint64(n >> 32)
    ll = xl * nl
    lh = xl * nh
    hl = xh * nl
    hh = xh * nh
    mid = (ll >> np.uint64(32)) + (lh & _M32) + (hl & _M32)
    out = hh + (lh >> np.uint64(32)) + (hl >> np.uint64(32)) + (mid >> np.uint64(32))
    return out.astype(np.int64)


def box_muller(u1, u2):
    """Standard normal from two [0,1) uniforms (cosine branch)."""
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(6.283185307179586 * u2)


def sub_split(split, k):
    """The 24-bit ``split`` field of the counter when a draw block needs one more index: the split
    number in the low 6 bits (MAX_SPLITS = 32), the sub-index ``k`` (< 2**18) above."""
    return (int(split) & 0x3F) | (int(k) << 6)


def normals(seed, step, split, index, count):
    """``[len(index), count]`` standard normals: normals ``2k`` and ``2k+1`` of row ``i`` are the cosine
    and sine branches of the Box-Muller pair of block ``(index=i, sub-index=k, TAG_NORMAL)``."""
    index = np.atleast_1d(np.asarray(index, dtype=np.uint64))
    out = np.empty((len(index), int(count)), dtype=np.float64)
    for k in range((int(count) + 1) // 2):
        w0, w1, w2, w3 = draw_words(seed, step, sub_split(split, k), TAG_NORMAL, index)
        r = np.sqrt(-2.0 * np.log(1.0 - u53(w0, w1)))
        th = 6.283185307179586 * u53(w2, w3)
        out[:, 2 * k] = r * np.cos(th)
        if 2 * k + 1 < count:
            out[:, 2 * k + 1] = r * np.sin(th)
    return out


def chol_psd(cov, max_rank=None):
    """Lower factor ``L`` with ``L L^T = cov`` for a positive SEMI-definite matrix: plain column
    Cholesky in which a pivot at or below ``1e-12 * max(diag)`` zeroes its column (for a PSD matrix the
    rest of that column is zero as well), and which stops after ``max_rank`` pivots: the sample covariance
    of ``s`` rows has rank at most ``s - 1``, and whatever a later pivot shows is rounding noise that a
    small earlier pivot has amplified -- taking its square root would put O(1e-6) garbage into the factor.
    This is the factorisation the draw specification fixes for ``multivariate_normal(mean, cov) := mean +
    L z`` (numpy's own uses an SVD whose sign / ordering conventions cannot be reproduced bit-for-bit by
    independent hardware)."""
    a = np.atleast_2d(np.asarray(cov, dtype=np.float64))
    n = a.shape[0]
    L = np.zeros_like(a)
    tol = 1e-12 * max(float(np.max(np.diag(a))), 0.0)
    left = n if max_rank is None else int(max_rank)
    for j in range(n):
        if left <= 0:
            break
        d = a[j, j] - np.dot(L[j, :j], L[j, :j])
        if not d > tol:
            continue
        left -= 1
        L[j, j] = np.sqrt(d)
        if j + 1 < n:
            L[j + 1 :, j] = (a[j + 1 :, j] - L[j + 1 :, :j] @ L[j, :j]) / L[j, j]
    return L


# --------------------------------------------------------------------------
# split permutation: a keyed bijection on [0, N) that can be evaluated for one
# element independently of all others (Feistel network + cycle walking; cf.
# Mitchell et al., "Bandwidth-optimal random shuffling for GPUs", 2021).
# --------------------------------------------------------------------------
def _fmix32(x):
    x = x & _M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & _M32
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & _M32
    x ^= x >> np.uint64(16)
    return x


def feistel_keys(seed, step):
    """The FEISTEL_ROUNDS round keys of step ``step``."""
    w = draw_words(seed, step, 0, TAG_SHUFFLE, np.arange(FEISTEL_ROUNDS // 4))
    return np.stack(w, axis=1).reshape(-1).astype(np.uint64)  # k[4*i + word]


def _half_bits(n):
    bits = max(int(n - 1).bit_length(), 2)
    return (bits + 1) // 2


def subset_indices(seed, step, split, i, n, count):
    """``choice(n, count, replace=False)`` of active rank ``i`` (walk.py:34): the first ``count`` images
    of the Feistel permutation of ``[0, n)`` keyed by blocks ``(index=i, sub-index 0..1, TAG_SUBSET)``;
    ``count == n`` is defined as the identity order (the covariance does not depend on the order)."""
    n, count = int(n), int(count)
    if count == n:
        return np.arange(n, dtype=np.int64)
    ws = [draw_words(seed, step, sub_split(split, k), TAG_SUBSET, np.array([i])) for k in range(FEISTEL_ROUNDS // 4)]
    keys = np.array([int(w[j][0]) for w in ws for j in range(4)], dtype=np.uint64)
    return _feistel_apply(np.arange(count, dtype=np.uint64), n, keys)


def _feistel_apply(x, n, keys):
    h = np.uint64(_half_bits(n))
    mask = (np.uint64(1) << h) - np.uint64(1)

    def enc(x):
        left, right = x >> h, x & mask
        for r in range(FEISTEL_ROUNDS):
            left, right = right, left ^ (_fmix32(right ^ keys[r]) & mask)
        return (left << h) | right

    x = enc(np.asarray(x, dtype=np.uint64))
    while True:
        bad = x >= np.uint64(n)
        if not bad.any():
            break
        x[bad] = enc(x[bad])
    return x.astype(np.int64)


def split_permutation(seed, step, n):
    """pi(w) for w in [0, n): balanced Feistel on 2*h bits, cycle-walked into
    [0, n)."""
    keys = feistel_keys(seed, step)
    h = np.uint64(_half_bits(n))
    mask = (np.uint64(1) << h) - np.uint64(1)

    def enc(x):
        left, right = x >> h, x & mask
        for r in range(FEISTEL_ROUNDS):
            left, right = right, left ^ (_fmix32(right ^ keys[r]) & mask)
        return (left << h) | right

    x = enc(np.arange(n, dtype=np.uint64))
    while True:
        bad = x >= np.uint64(n)
        if not bad.any():
            break
        x[bad] = enc(x[bad])
    return x.astype(np.int64)


def split_assignment(seed, step, n, nsplits, randomize):
    """``inds`` of ``red_blue.py:77-80``: ``arange(n) % nsplits``, shuffled by
    the step's keyed permutation when ``randomize``:
    ``shuffle(x)  :=  x[:] = x[pi]``."""
    inds = np.arange(n, dtype=np.int64) % int(nsplits)
    if randomize:
        inds = inds[split_permutation(seed, step, n)]
    return inds


def move_choice(seed, step, weights):
    """Index of the move drawn for step ``step`` (``ensemble.py:406``):
    inverse-CDF on one uniform, as ``RandomState.choice(a, p=p)`` does."""
    w0, w1, _, _ = draw_words(seed, step, 0, TAG_MOVE, np.zeros(1, np.uint64))
    u = u53(w0, w1)[0]
    cdf = np.cumsum(np.asarray(weights, dtype=np.float64))
    cdf /= cdf[-1]
    return int(min(np.searchsorted(cdf, u, side="right"), len(cdf) - 1))


# --------------------------------------------------------------------------
# the shim
# --------------------------------------------------------------------------
class PhiloxRandom(object):
    """Duck-typed ``RandomState`` that maps the reference's sequential call
    pattern onto counter-addressed draws.

    It recognises the call sequence of one ``EnsembleSampler.sample`` step:
    ``choice(moves, p=w)`` opens a step; ``shuffle(int64[N])`` is the split
    permutation; the first proposal-type call after a step start or after an
    accept phase opens the next split; scalar ``rand()`` calls are the accept
    draws of consecutive active ranks.
    """

    def __init__(self, seed, step=0):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.step = int(step)  # step index the NEXT ``choice`` call will open
        self._cur = None  # step being executed
        self._split = -1
        self._phase = "idle"
        self._i = 0  # active rank cursor (snooker proposal loop)
        self._k = 0  # randint cursor inside one snooker walker
        self._j = 0  # accept cursor
        self.trace = None  # optional list collecting (kind, step, split, payload)

    # -- bookkeeping -------------------------------------------------------
    def get_state(self):
        return ("philox4x32-10", self.seed, self.step)

    def set_state(self, state):
        # ``ensemble.py:335`` calls this with None; the reference swallows the
        # resulting exception (``ensemble.py:235-238``).
        kind, seed, step = state
        if kind != "philox4x32-10":
            raise ValueError("not a Philox state")
        self.seed, self.step = int(seed), int(step)

    def _log(self, kind, payload):
        if self.trace is not None:
            self.trace.append((kind, self._cur, self._split, payload))

    def _open_split(self):
        if self._phase != "proposal":
            self._split += 1
            self._phase = "proposal"
            self._i = 0
            self._k = 0

    def _words(self, tag, index):
        return draw_words(self.seed, self._cur, self._split, tag, index)

    # -- the seven methods -------------------------------------------------
    def choice(self, a, size=None, replace=True, p=None):
        if isinstance(a, (int, np.integer)) and not replace:
            # WalkMove: ``random.choice(Nc, s0, replace=False)`` once per active walker  (walk.py:34)
            assert p is None
            self._open_split()
            out = subset_indices(self.seed, self._cur, self._split, self._i, int(a), int(size))
            self._walk_rank = int(size) - 1  # rank bound of the covariance the next call receives
            self._log("walk_subset", out.copy())
            return out
        if isinstance(a, (int, np.integer)):
            # DEMove: ``random.choice(nc*(nc-1), size=ns, replace=True)``  (de.py:49)
            assert replace and p is None
            self._open_split()
            n = int(size)
            w0, w1, _, _ = self._words(TAG_PROP_A, np.arange(n))
            out = bounded64(w0, w1, int(a))
            self._log("de_pair", out.copy())
            return out
        # move selection (ensemble.py:406): opens a step
        self._cur = self.step
        self.step += 1
        self._split = -1
        self._phase = "step"
        idx = move_choice(self.seed, self._cur, p) if p is not None else 0
        self._log("move", idx)
        return a[idx]

    def shuffle(self, x):
        if x.ndim == 1:
            # split assignment (red_blue.py:80)
            assert self._phase == "step"
            perm = split_permutation(self.seed, self._cur, len(x))
            x[:] = x[perm]
            self._log("inds", x.copy())
            return
        # DESnookerMove: ``random.shuffle(w)`` on a [3, ndim] array
        # (de_snooker.py:39) -> one of the 6 row orders, drawn from block B.
        assert x.shape[0] == 3 and self._phase == "proposal" and self._k == 3
        _, _, w2, w3 = self._words(TAG_PROP_B, np.array([self._i]))
        p = int(bounded64(w2, w3, 6)[0])
        order = SNOOKER_PERMS[p]
        x[:] = x[list(order)]
        self._log("snooker_perm", p)
        self._i += 1
        self._k = 0

    def multivariate_normal(self, mean, cov):
        mean = np.asarray(mean, dtype=np.float64)
        if self._phase in ("step", "mh"):
            L = chol_psd(cov)
            # GaussianMove with a full covariance: ONE draw per step, added to every walker (gaussian.py:116-118)
            self._phase = "mh"
            z = normals(self.seed, self._cur, 0, np.array([0]), len(mean))[0]
            self._log("mh_mvn", z.copy())
            return mean + L @ z
        # WalkMove: ``random.multivariate_normal(s[i], cov)`` for active rank i  (walk.py:36)
        assert self._phase == "proposal"
        L = chol_psd(cov, max_rank=self._walk_rank)
        z = normals(self.seed, self._cur, self._split, np.array([self._i]), len(mean))[0]
        self._log("walk_z", z.copy())
        self._i += 1
        return mean + L @ z

    def uniform(self, low, high):
        # GaussianMove ``factor``: ``rng.uniform(-log f, log f)`` once per step  (gaussian.py:91)
        assert self._phase in ("step", "mh")
        self._phase = "mh"
        w0, w1, _, _ = draw_words(self.seed, self._cur, 0, TAG_MOVE, np.array([1]))
        u = float(u53(w0, w1)[0])
        self._log("mh_factor", u)
        return low + (high - low) * u

    def rand(self, *shape):
        if len(shape) == 1 and self._phase == "mh":
            # MHMove accept draws: ``model.random.rand(nwalkers)``  (mh.py:58)
            (n,) = shape
            w0, w1, _, _ = draw_words(self.seed, self._cur, 0, TAG_ACCEPT, np.arange(n))
            out = u53(w0, w1)
            self._log("u_accept_mh", out.copy())
            self._phase = "idle"
            return out
        if len(shape) == 0:
            # accept draw (red_blue.py:100)
            if self._phase == "proposal":
                self._phase = "accept"
                self._j = 0
            w0, w1, _, _ = self._words(TAG_ACCEPT, np.array([self._j]))
            self._j += 1
            u = float(u53(w0, w1)[0])
            self._log("u_accept", u)
            return u
        # StretchMove: ``random.rand(Ns)``  (stretch.py:30)
        (n,) = shape
        self._open_split()
        w0, w1, _, _ = self._words(TAG_PROP_A, np.arange(n))
        out = u53(w0, w1)
        self._log("u_stretch", out.copy())
        return out

    def randint(self, low, high=None, size=None):
        assert high is None
        if self._phase == "mh":
            # GaussianMove mode="random": ``rng.randint(ndim, size=nw)``  (gaussian.py:100)
            n = int(size)
            w0, w1, _, _ = draw_words(self.seed, self._cur, 0, TAG_PROP_B, np.arange(n))
            out = bounded64(w0, w1, int(low))
            self._log("mh_dim", out.copy())
            return out
        if size is None:
            # DESnookerMove: ``random.randint(Nc[j])`` for j = 0, 1, 2
            # (de_snooker.py:38)
            self._open_split()
            i, k = self._i, self._k
            assert k < 3
            if k < 2:
                w = self._words(TAG_PROP_A, np.array([i]))
                lo, hi = w[2 * k], w[2 * k + 1]
            else:
                w = self._words(TAG_PROP_B, np.array([i]))
                lo, hi = w[0], w[1]
            self._k += 1
            r = int(bounded64(lo, hi, int(low))[0])
            self._log("snooker_rint", r)
            return r
        # StretchMove: ``random.randint(Nc, size=(Ns,))``  (stretch.py:32)
        (n,) = size
        assert self._phase == "proposal"
        _, _, w2, w3 = self._words(TAG_PROP_A, np.arange(n))
        out = bounded64(w2, w3, int(low))
        self._log("rint", out.copy())
        return out

    def randn(self, *shape):
        if self._phase in ("step", "mh"):
            # GaussianMove: ``rng.randn(nwalkers, ndim)`` -- normal (w, d) of row w  (gaussian.py:97)
            self._phase = "mh"
            nw, nd = shape
            out = normals(self.seed, self._cur, 0, np.arange(nw), nd)
            self._log("mh_randn", out.copy())
            return out
        # DEMove: ``random.randn(ns, 1)``  (de.py:56)
        n = int(np.prod(shape))
        assert self._phase == "proposal"
        w0, w1, w2, w3 = self._words(TAG_PROP_B, np.arange(n))
        out = box_muller(u53(w0, w1), u53(w2, w3)).reshape(shape)
        self._log("randn", out.copy())
        return out


# row orders of the 3 snooker helper rows: after ``shuffle(w)`` the rows are
# ``w[perm]``; index drawn uniformly on [0, 6).
SNOOKER_PERMS = (
    (0, 1, 2),
    (0, 2, 1),
    (1, 0, 2),
    (1, 2, 0),
    (2, 0, 1),
    (2, 1, 0),
)
