"""Vectorised numpy restatement of the reference's walker-update path
(TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``).

Follows, line by line:

* ``EnsembleSampler.sample`` step loop ........ ``src/emcee/ensemble.py:401-424``
  (one move drawn per step for the whole ensemble, ``:406``)
* ``RedBlueMove.propose`` ..................... ``src/emcee/moves/red_blue.py:52-106``
  (split assignment ``:76-80``, per-split gather ``:82-87``, accept ``:96-101``,
  update ``:103-104`` -> ``moves/move.py:29-34``)
* ``StretchMove.get_proposal`` ................ ``src/emcee/moves/stretch.py:26-33``
* ``DEMove.setup/get_proposal`` ............... ``src/emcee/moves/de.py:33-64``
  with ``_get_nondiagonal_pairs`` (``de.py:67-77``) decoded analytically
   instead of materialising the O(Nc^2) table
* ``DESnookerMove.get_proposal`` .............. ``src/emcee/moves/de_snooker.py:31-46``
* ``WalkMove.get_proposal`` ................... ``src/emcee/moves/walk.py:27-37``
* ``MHMove.propose`` + ``GaussianMove`` ....... ``src/emcee/moves/mh.py:35-65``,
  ``src/emcee/moves/gaussian.py:72-119``
* ``EnsembleSampler.compute_log_prob`` guards . ``src/emcee/ensemble.py:476-479,550-551``

Random numbers come from the counter-addressed draw specification in
``oracle/philox.py`` -- the same values ``PhiloxRandom`` hands to the unmodified
reference, so this file and the reference must agree bit for bit (pinned by
``tests/test_oracle_golden.py`` against ``tests/golden/*.npz``).
"""

import numpy as np

from . import philox as px

__all__ = ["Stretch", "DE", "Snooker", "Walk", "Gaussian", "OracleSampler", "de_pair_decode"]


class _RedBlue(object):
    def __init__(self, nsplits=2, randomize_split=True, live_dangerously=False):
        self.nsplits = int(nsplits)
        self.randomize_split = bool(randomize_split)
        self.live_dangerously = bool(live_dangerously)


class Stretch(_RedBlue):
    kind = "stretch"

    def __init__(self, a=2.0, **kw):
        self.a = a
        super().__init__(**kw)


class DE(_RedBlue):
    kind = "de"

    def __init__(self, sigma=1.0e-5, gamma0=None, **kw):
        self.sigma, self.gamma0 = sigma, gamma0
        super().__init__(**kw)


class Snooker(_RedBlue):
    kind = "snooker"

    def __init__(self, gammas=1.7, **kw):
        self.gammas = gammas
        kw["nsplits"] = 4  # de_snooker.py:28
        super().__init__(**kw)


class Walk(_RedBlue):
    kind = "walk"

    def __init__(self, s=None, **kw):
        self.s = s
        super().__init__(**kw)


class Gaussian(object):
    """``GaussianMove(cov, mode, factor)`` (``gaussian.py:32-69``): scalar, vector or matrix ``cov``."""

    kind = "gaussian"

    def __init__(self, cov, mode="vector", factor=None):
        c = np.asarray(cov, dtype=np.float64)
        if c.ndim == 0:
            self.form, self.scale = "iso", float(np.sqrt(c))  # gaussian.py:58
        elif c.ndim == 1:
            self.form, self.scale = "diag", np.sqrt(c)  # gaussian.py:45
        elif c.ndim == 2 and c.shape[0] == c.shape[1]:
            self.form, self.scale = "full", c  # gaussian.py:50 (the matrix itself)
            if mode != "vector":
                raise ValueError("'{0}' is not a recognized mode.".format(mode))  # gaussian.py:111
        else:
            raise ValueError("Invalid proposal scale dimensions")
        if factor is not None and factor < 1.0:
            raise ValueError("'factor' must be >= 1.0")
        self.mode, self.factor = mode, factor
        self.index = 0  # gaussian.py:64, advanced by mode="sequential"
        self.nsplits, self.randomize_split, self.live_dangerously = 1, False, True


def de_pair_decode(m, n):
    """Row ``m`` of ``_get_nondiagonal_pairs(n)`` (``de.py:67-77``) without the
    table: the first ``T = n(n-1)/2`` rows are ``np.tril_indices(n, -1)`` in
    row-major order ``(r, col)``, the last ``T`` the same with columns swapped."""
    m = np.asarray(m, dtype=np.int64)
    T = n * (n - 1) // 2
    upper = m >= T
    k = np.where(upper, m - T, m)
    r = ((1.0 + np.sqrt(1.0 + 8.0 * k.astype(np.float64))) * 0.5).astype(np.int64)
    # fix the float estimate: largest r with r(r-1)/2 <= k
    r = np.where(r * (r - 1) // 2 > k, r - 1, r)
    r = np.where((r + 1) * r // 2 <= k, r + 1, r)
    col = k - r * (r - 1) // 2
    p0 = np.where(upper, col, r)
    p1 = np.where(upper, r, col)
    return p0, p1


class OracleSampler(object):
    """State + step loop.  ``log_prob`` is a vectorised callable
    ``[M, D] -> [M]`` (``oracle/targets.py``)."""

    def __init__(self, nwalkers, ndim, log_prob, moves=None, seed=0, step=0):
        self.nwalkers, self.ndim = int(nwalkers), int(ndim)
        self.log_prob_fn = log_prob
        if moves is None:
            moves = [(Stretch(), 1.0)]
        elif isinstance(moves, (_RedBlue, Gaussian)):
            moves = [(moves, 1.0)]
        self.moves = [m for m, _ in moves]
        w = np.array([w for _, w in moves], dtype=np.float64)
        self.weights = w / np.sum(w)  # ensemble.py:128-129
        self.seed = int(seed)
        self.step = int(step)
        self.coords = None
        self.log_prob = None
        self.naccepted = np.zeros(self.nwalkers, dtype=np.int64)
        self.rowwise = False  # per-row np.dot for snooker (bit-exact vs reference)
        self.taps = None  # last half-step's draws, for known-answer tests
        # sharded emulation (tests of the multi-GPU scheme): only walkers in
        # [owner_range) are updated by this instance, and ``exchange(coords,
        # log_prob, accepted)`` is called after every split to merge the ranks
        self.owner_range = None
        self.exchange = None

    # -- ensemble.py:458-553 (only the parts that exist for a device model) --
    def compute_log_prob(self, coords):
        p = np.asarray(coords, dtype=np.float64)
        if np.any(np.isinf(p)):
            raise ValueError("At least one parameter value was infinite")
        if np.any(np.isnan(p)):
            raise ValueError("At least one parameter value was NaN")
        lp = np.asarray(self.log_prob_fn(p), dtype=np.float64)
        if np.any(np.isnan(lp)):
            raise ValueError("Probability function returned NaN")
        return lp

    def set_state(self, coords, log_prob=None):
        self.coords = np.array(coords, dtype=np.float64, copy=True)
        assert self.coords.shape == (self.nwalkers, self.ndim)
        if log_prob is None:
            log_prob = self.compute_log_prob(self.coords)
        self.log_prob = np.array(log_prob, dtype=np.float64, copy=True)
        if np.any(np.isnan(self.log_prob)):
            raise ValueError("The initial log_prob was NaN")

    # -- proposals -----------------------------------------------------------
    def _stretch(self, mv, s, sets_c, step, split):
        comp = np.concatenate(sets_c)  # stretch.py:27 (walker id by complement rank)
        Ns, Nc = len(s), len(comp)
        w0, w1, w2, w3 = px.draw_words(self.seed, step, split, px.TAG_PROP_A, np.arange(Ns))
        zz = ((mv.a - 1.0) * px.u53(w0, w1) + 1) ** 2.0 / mv.a  # stretch.py:30
        factors = (self.ndim - 1.0) * np.log(zz)  # stretch.py:31
        rint = px.bounded64(w2, w3, Nc)  # stretch.py:32
        c = self.coords[comp[rint]]
        q = c - (c - s) * zz[:, None]  # stretch.py:33
        self.taps = dict(zz=zz, rint=rint, partner=comp[rint])
        return q, factors

    def _de(self, mv, s, sets_c, step, split):
        comp = np.concatenate(sets_c)  # de.py:41
        Ns, Nc = len(s), len(comp)
        g0 = mv.gamma0
        if g0 is None:
            g0 = 2.38 / np.sqrt(2 * self.ndim)  # de.py:33-38
        i = np.arange(Ns)
        a0, a1, _, _ = px.draw_words(self.seed, step, split, px.TAG_PROP_A, i)
        m = px.bounded64(a0, a1, Nc * (Nc - 1))  # de.py:49
        p0, p1 = de_pair_decode(m, Nc)
        diffs = self.coords[comp[p1]] - self.coords[comp[p0]]  # de.py:53
        b0, b1, b2, b3 = px.draw_words(self.seed, step, split, px.TAG_PROP_B, i)
        n = px.box_muller(px.u53(b0, b1), px.u53(b2, b3))
        gamma = g0 * (1 + mv.sigma * n[:, None])  # de.py:56
        q = s + gamma * diffs  # de.py:62
        self.taps = dict(pair=m, p0=comp[p0], p1=comp[p1], gamma=gamma[:, 0])
        return q, np.zeros(Ns, dtype=np.float64)  # de.py:64

    def _snooker(self, mv, s, sets_c, step, split):
        assert len(sets_c) == 3
        Ns = len(s)
        i = np.arange(Ns)
        a0, a1, a2, a3 = px.draw_words(self.seed, step, split, px.TAG_PROP_A, i)
        b0, b1, b2, b3 = px.draw_words(self.seed, step, split, px.TAG_PROP_B, i)
        r = [
            px.bounded64(a0, a1, len(sets_c[0])),
            px.bounded64(a2, a3, len(sets_c[1])),
            px.bounded64(b0, b1, len(sets_c[2])),
        ]  # de_snooker.py:38
        perm = np.array(px.SNOOKER_PERMS)[px.bounded64(b2, b3, 6)]  # de_snooker.py:39
        w = np.stack([sets_c[j][r[j]] for j in range(3)], axis=1)  # walker ids [Ns, 3]
        zi = np.take_along_axis(w, perm, axis=1)
        z, z1, z2 = (self.coords[zi[:, k]] for k in range(3))
        delta = s - z  # de_snooker.py:41
        if self.rowwise:
            dot = lambda a, b: np.array([np.dot(x, y) for x, y in zip(a, b)])
        else:
            dot = lambda a, b: np.einsum("ij,ij->i", a, b)
        norm = np.sqrt(dot(delta, delta))  # de_snooker.py:42
        u = delta / norm[:, None]  # de_snooker.py:43
        q = s + u * mv.gammas * (dot(u, z1) - dot(u, z2))[:, None]  # de_snooker.py:44
        dq = q - z
        metropolis = np.log(np.sqrt(dot(dq, dq))) - np.log(norm)  # de_snooker.py:45
        self.taps = dict(z=zi[:, 0], z1=zi[:, 1], z2=zi[:, 2])
        return q, (self.ndim - 1.0) * metropolis  # de_snooker.py:46

    def _walk(self, mv, s, sets_c, step, split):
        comp = np.concatenate(sets_c)  # walk.py:28
        Ns, Nc = len(s), len(comp)
        s0 = Nc if mv.s is None else int(mv.s)  # walk.py:32
        q = np.empty_like(s)
        z = px.normals(self.seed, step, split, np.arange(Ns), self.ndim)
        c = self.coords[comp]
        shared = None
        for i in range(Ns):
            if s0 == Nc:  # the whole complement, in its own order: one covariance for the split
                if shared is None:
                    shared = px.chol_psd(np.atleast_2d(np.cov(c, rowvar=0)), max_rank=s0 - 1)  # walk.py:35
                L = shared
            else:
                inds = px.subset_indices(self.seed, step, split, i, Nc, s0)  # walk.py:34
                L = px.chol_psd(np.atleast_2d(np.cov(c[inds], rowvar=0)), max_rank=s0 - 1)  # walk.py:35
            q[i] = s[i] + L @ z[i]  # walk.py:36 with multivariate_normal := mean + chol(cov) z
        self.taps = dict(z=z)
        return q, np.zeros(Ns, dtype=np.float64)  # walk.py:37

    # -- mh.py:35-65 with gaussian.py:72-119 as the proposal ------------------
    def _propose_mh(self, mv, step):
        N, D = self.nwalkers, self.ndim
        x0 = self.coords
        f = 1.0
        if mv.factor is not None:  # gaussian.py:88-91
            w0, w1, _, _ = px.draw_words(self.seed, step, 0, px.TAG_MOVE, np.array([1]))
            lf = np.log(mv.factor)
            f = np.exp(-lf + (lf - (-lf)) * float(px.u53(w0, w1)[0]))
        if mv.form == "full":
            z = px.normals(self.seed, step, 0, np.array([0]), D)[0]
            xnew = x0 + f * (np.zeros(D) + px.chol_psd(mv.scale) @ z)  # gaussian.py:116-118: one draw for all walkers
        else:
            xnew = x0 + f * mv.scale * px.normals(self.seed, step, 0, np.arange(N), D)  # gaussian.py:97
        if mv.mode == "vector":
            q = xnew
        else:
            if mv.mode == "random":  # gaussian.py:100
                w0, w1, _, _ = px.draw_words(self.seed, step, 0, px.TAG_PROP_B, np.arange(N))
                dim = px.bounded64(w0, w1, D)
            else:  # sequential, gaussian.py:102-103
                dim = np.full(N, mv.index % D, dtype=np.int64)
                mv.index = (mv.index + 1) % D
            q = np.array(x0)
            q[np.arange(N), dim] = xnew[np.arange(N), dim]  # gaussian.py:106-107
        new_lp = self.compute_log_prob(q)  # mh.py:54
        u0, u1, _, _ = px.draw_words(self.seed, step, 0, px.TAG_ACCEPT, np.arange(N))
        with np.errstate(divide="ignore", invalid="ignore"):
            lnpdiff = new_lp - self.log_prob + np.zeros(N)  # mh.py:57
            acc = np.log(px.u53(u0, u1)) < lnpdiff  # mh.py:58
        self.taps = dict(q=q, new_lp=new_lp, u_accept=px.u53(u0, u1))
        self.coords[acc] = q[acc]  # mh.py:62 -> move.py:33
        self.log_prob[acc] = new_lp[acc]
        return acc

    # -- red_blue.py:52-106 --------------------------------------------------
    def _propose(self, mv, step):
        if mv.kind == "gaussian":
            return self._propose_mh(mv, step)
        N, D = self.nwalkers, self.ndim
        if N < 2 * D and not mv.live_dangerously:  # red_blue.py:64-70
            raise RuntimeError(
                "It is unadvisable to use a red-blue move "
                "with fewer walkers than twice the number of "
                "dimensions."
            )
        accepted = np.zeros(N, dtype=bool)
        inds = px.split_assignment(self.seed, step, N, mv.nsplits, mv.randomize_split)
        get = {"stretch": self._stretch, "de": self._de, "snooker": self._snooker, "walk": self._walk}[mv.kind]
        for split in range(mv.nsplits):
            sets = [np.flatnonzero(inds == j) for j in range(mv.nsplits)]  # red_blue.py:85
            act = sets[split]
            sets_c = sets[:split] + sets[split + 1 :]  # red_blue.py:87
            s = self.coords[act]
            q, factors = get(mv, s, sets_c, step, split)
            new_lp = self.compute_log_prob(q)  # red_blue.py:93
            u0, u1, _, _ = px.draw_words(
                self.seed, step, split, px.TAG_ACCEPT, np.arange(len(act))
            )
            uacc = px.u53(u0, u1)
            with np.errstate(divide="ignore", invalid="ignore"):
                lnpdiff = factors + new_lp - self.log_prob[act]  # red_blue.py:99
                acc = lnpdiff > np.log(uacc)  # red_blue.py:100
            self.taps.update(u_accept=uacc, active=act, q=q, new_lp=new_lp, factors=factors)
            if self.owner_range is not None:
                acc = acc & (act >= self.owner_range[0]) & (act < self.owner_range[1])
            won = act[acc]
            self.coords[won] = q[acc]  # move.py:33
            self.log_prob[won] = new_lp[acc]  # move.py:34
            accepted[won] = True
            if self.exchange is not None:
                self.exchange(self.coords, self.log_prob, accepted)
        return accepted

    def run(self, nsteps):
        """``nsteps`` iterations of ``ensemble.py:403-424`` with ``store=False``;
        returns the accept mask of the last step."""
        accepted = np.zeros(self.nwalkers, dtype=bool)
        for _ in range(int(nsteps)):
            mi = px.move_choice(self.seed, self.step, self.weights)  # ensemble.py:406
            accepted = self._propose(self.moves[mi], self.step)
            self.naccepted += accepted
            self.step += 1
        return accepted
