"""``EnsembleSampler`` on the B200 engine (reference:
``src/emcee/ensemble.py:32-623``).

Same constructor / ``sample`` / ``run_mcmc`` / ``compute_log_prob`` surface and
the same exceptions, but ``log_prob_fn`` is a registered device model
(``emcee_b200.models``) and every step runs inside the CUDA library through the
C ABI: the walker array stays in HBM between steps, ``run_mcmc`` is one call
for all iterations, and stored steps stream back through pinned buffers.
Features that need a host callback per walker (``pool``, ``args``/``kwargs``,
blobs, named parameters) raise ``NotImplementedError``."""

from collections.abc import Iterable

import numpy as np

from . import _lib
from .backend import Backend
from .model import Model
from .models import DeviceModel
from .moves import StretchMove
from .rng import DeviceRandom
from .state import State

__all__ = ["EnsembleSampler", "walkers_independent"]


def _seed_from_numpy():
    """Like the reference, start from numpy's global generator
    (``ensemble.py:140`` ``state = np.random.get_state()``) without consuming it."""
    keys, pos = np.random.get_state()[1:3]
    a, b = int(keys[pos % 624]), int(keys[(pos + 1) % 624])
    return ((a << 32) | b) ^ (int(pos) * 0x9E3779B97F4A7C15 & (2**64 - 1))


class EnsembleSampler(object):
    """An ensemble MCMC sampler whose walker update runs on one B200.

    Args mirror ``ensemble.py:79-98``.  Extra keyword-only arguments: ``seed``
    (Philox key; default derived from numpy's global state), ``device``, and
    ``pinned_results`` (the yielded / returned ``State`` arrays are views of two
    page-locked buffers owned by the sampler and are overwritten by the next
    step -- full-speed D2H for callers that consume each state before asking for
    the next; default ``False`` = fresh arrays, as the reference returns)."""

    def __init__(
        self,
        nwalkers,
        ndim,
        log_prob_fn,
        pool=None,
        moves=None,
        args=None,
        kwargs=None,
        backend=None,
        vectorize=False,
        blobs_dtype=None,
        parameter_names=None,
        # deprecated in the reference; rejected here
        a=None,
        postargs=None,
        threads=None,
        live_dangerously=None,
        runtime_sortingfn=None,
        *,
        seed=None,
        device=0,
        pinned_results=False,
    ):
        for name, val in (("a", a), ("postargs", postargs), ("threads", threads),
                          ("live_dangerously", live_dangerously), ("runtime_sortingfn", runtime_sortingfn)):
            if val is not None:
                raise NotImplementedError("the deprecated '%s' argument is not supported; use 'moves'" % name)
        if not isinstance(log_prob_fn, DeviceModel):
            raise TypeError(
                "log_prob_fn must be a registered device model (emcee_b200.models.*): the "
                "walker update runs on the GPU and cannot call back into Python"
            )
        if pool is not None:
            raise NotImplementedError("pool: log-probabilities are evaluated on the GPU, not through map()")
        if args or kwargs:
            raise NotImplementedError("args/kwargs: put the parameters into the device model")
        if parameter_names is not None:
            raise NotImplementedError("parameter_names need a host callable")
        if blobs_dtype is not None:
            raise NotImplementedError("blobs are not supported on the device path")

        # move schedule (ensemble.py:115-129)
        if moves is None:
            self._moves, weights = [StretchMove()], [1.0]
        elif isinstance(moves, Iterable):
            moves = list(moves)
            try:
                self._moves, weights = (list(t) for t in zip(*moves))
            except TypeError:
                self._moves, weights = moves, np.ones(len(moves))
        else:
            self._moves, weights = [moves], [1.0]
        self._raw_weights = np.atleast_1d(weights).astype(float)
        self._weights = self._raw_weights / np.sum(self._raw_weights)
        for m in self._moves:
            # the step loop runs inside the CUDA library: a move must be one of the device moves
            # (it only *describes* itself); a host ``Move`` with its own ``propose`` cannot be called
            # back from a kernel -- say so here, not with an AttributeError in the middle of sample()
            if not callable(getattr(m, "descriptor", None)):
                raise TypeError(
                    "moves must be emcee_b200.moves.* device moves; got {0!r}, which has no device "
                    "descriptor (host-side Move subclasses cannot run inside the fused step kernels)".format(m)
                )

        self.pool = None
        self.vectorize = True  # the device path is always batched
        self.blobs_dtype = None
        self.ndim = int(ndim)
        self.nwalkers = int(nwalkers)
        self.log_prob_fn = log_prob_fn
        self.params_are_named = False

        self._device = int(device)
        self._engine = _lib.Engine(self.nwalkers, self.ndim, _seed_from_numpy() if seed is None else seed,
                                   device=device)
        self._engine.set_model(log_prob_fn.kind, log_prob_fn.device_params(self.ndim))
        self._random = DeviceRandom(self._engine)
        self._pinned = None
        if pinned_results:
            self._pinned = (_lib.pinned_empty((self.nwalkers, self.ndim)), _lib.pinned_empty((self.nwalkers,)))
        self._rdv = None  # multi-GPU: the host rendezvous this sampler is attached to (``attach``)
        self._gather_results = True

        self.backend = Backend() if backend is None else backend
        if not self.backend.initialized:  # ensemble.py:137-141
            self._previous_state = None
            self.reset()
        else:
            if self.backend.shape != (self.nwalkers, self.ndim):
                raise ValueError(
                    "the shape of the backend ({0}) is incompatible with the "
                    "shape of the sampler ({1})".format(self.backend.shape, (self.nwalkers, self.ndim))
                )
            self.random_state = self.backend.random_state  # silently ignored if foreign
            if self.backend.iteration > 0:
                self._previous_state = self.get_last_sample()
            else:
                self._previous_state = None

    # ------------------------------------------------------------------ state
    @property
    def random_state(self):
        """``("philox4x32-10", seed, step)`` -- the complete random state."""
        return self._random.get_state()

    @random_state.setter
    def random_state(self, state):
        # like ensemble.py:228-238: try, and stay as we are if it is garbage
        try:
            self._random.set_state(state)
        except Exception:
            pass

    @property
    def iteration(self):
        return self.backend.iteration

    def reset(self):
        self.backend.reset(self.nwalkers, self.ndim)

    @property
    def model(self):
        """The 4-tuple the reference hands to ``move.propose`` (``ensemble.py:393-395``),
        for callers that drive a move directly through the plugin boundary."""
        return Model(self.log_prob_fn, self.compute_log_prob, map, self._random)

    def __getstate__(self):
        """Picklable like the reference (``ensemble.py:251-256``, pinned by
        ``tests/unit/test_sampler.py:225-234``): the GPU engine is dropped and
        rebuilt on unpickling from the model, the Philox ``(seed, step)`` and
        the device index; the walker state travels through ``_previous_state``
        / the backend as it does in the reference."""
        d = dict(self.__dict__)
        d["_saved_rng"] = self._engine.get_rng()
        d["_saved_pinned"] = self._pinned is not None
        for k in ("_engine", "_random", "_pinned"):
            d.pop(k, None)
        d["_rdv"] = None  # a communicator does not survive pickling: re-attach after loading
        d["pool"] = None
        return d

    def __setstate__(self, d):
        seed, step = d.pop("_saved_rng")
        pinned = d.pop("_saved_pinned")
        self.__dict__.update(d)
        self._engine = _lib.Engine(self.nwalkers, self.ndim, seed, device=self._device)
        self._engine.set_model(self.log_prob_fn.kind, self.log_prob_fn.device_params(self.ndim))
        self._engine.set_rng(seed, step)
        self._random = DeviceRandom(self._engine)
        self._pinned = None
        if pinned:
            self._pinned = (_lib.pinned_empty((self.nwalkers, self.ndim)), _lib.pinned_empty((self.nwalkers,)))

    # ------------------------------------------------------------ multi-GPU
    def attach(self, rdv, mode="p2p", gather_results=True):
        """Shard the ensemble by row block over the ranks of ``rdv``
        (:class:`emcee_b200.dist.Rendezvous`; one process per GPU, every rank
        builds the same sampler with the same seed and calls the same methods).

        ``gather_results=True``: yielded / returned states hold every walker
        (one collective replication per read-back).  ``False``: only the rows
        this rank owns (``owned_rows``) cross PCIe -- the other rows of the
        returned arrays are not refreshed."""
        from . import dist

        dist.attach(self._engine, rdv, mode)
        self._rdv = rdv
        self._gather_results = bool(gather_results)

    @property
    def owned_rows(self):
        """``slice`` of the walkers this process updates (all of them on one GPU)."""
        r0, n = self._engine.owned_rows()
        return slice(r0, r0 + n)

    def enable_moments(self, every=1):
        """Accumulate the chain mean / covariance on the device after every
        ``every``-th step (``0`` disables; calling it resets the accumulators)."""
        self._engine.set_option("moments_every", int(every))

    def moments(self):
        """``(mean[ndim], cov[ndim, ndim], count)`` over the ``(step, walker)``
        samples accumulated since :meth:`enable_moments` -- what
        ``np.mean`` / ``np.cov(rowvar=False)`` of ``get_chain(flat=True)`` give,
        without storing or downloading the chain (``store=False`` runs,
        ``ensemble.py:287-291``).  Collective on a sharded ensemble."""
        from . import dist

        part = self._engine.moments()
        parts = [part] if self._rdv is None else self._rdv.allgather(part)
        mean, cov, n, _ = dist.combine_moments(parts)
        return mean, cov, n

    # ------------------------------------------------------------- the driver
    def _schedule(self):
        return [(m.descriptor(), w) for m, w in zip(self._moves, self._raw_weights)]

    def _after_steps(self):
        """Advance the host mirrors of stateful moves by what the engine just ran (``GaussianMove``
        mode ``"sequential"`` keeps a running dimension index, ``gaussian.py:102-103``)."""
        stateful = [m for m in self._moves if hasattr(m, "_advance")]
        if stateful:
            picks = self._engine.move_picks(len(self._moves))
            for m, p in zip(self._moves, picks):
                if hasattr(m, "_advance"):
                    m._advance(int(p), self.ndim)

    def sample(
        self,
        initial_state,
        log_prob0=None,
        rstate0=None,
        blobs0=None,
        iterations=1,
        tune=False,
        skip_initial_state_check=False,
        thin_by=1,
        thin=None,
        store=True,
        progress=False,
        progress_kwargs=None,
        _bulk=False,
    ):
        """Advance the chain as a generator (``ensemble.py:258-424``): yields the
        live :class:`State` every ``thin_by`` steps.

        Device-detected errors (NaN log-probability, non-finite proposal -- the
        reference's ``ValueError`` s, ``ensemble.py:476-479,550-551``) are raised
        when the C-ABI call that contains the offending step returns: per yielded
        state here, after the whole run for :meth:`run_mcmc` (one call)."""
        if log_prob0 is not None or rstate0 is not None or blobs0 is not None:
            raise NotImplementedError("log_prob0/rstate0/blobs0 are deprecated in the reference; pass a State")
        pbar = None
        if progress:
            # the reference wraps tqdm (pbar.py:33-60); one tick per yielded state here
            try:
                import tqdm

                total = None if iterations is None else iterations
                pbar = tqdm.tqdm(total=total, **(progress_kwargs or {}))
            except ImportError:
                pbar = None
        if iterations is None and store:
            raise ValueError("'store' must be False when 'iterations' is None")

        # ``State(initial_state, copy=True)`` in the reference (ensemble.py:312): here the
        # upload to the device IS the copy -- the caller's arrays are only read, and
        # the yielded State gets its own arrays from the first device read-back
        state = State(initial_state)
        state = State(state.coords, log_prob=state.log_prob, blobs=state.blobs, random_state=state.random_state)
        state_shape = np.shape(state.coords)
        if state_shape != (self.nwalkers, self.ndim):
            raise ValueError("incompatible input dimensions {0}".format(state_shape))
        if state.blobs is not None:
            raise NotImplementedError("blobs are not supported on the device path")
        if (not skip_initial_state_check) and (not self._walkers_independent(state.coords)):
            raise ValueError(
                "Initial state has a large condition number. "
                "Make sure that your walkers are linearly independent for the "
                "best performance"
            )
        self.random_state = state.random_state  # ensemble.py:335 (ignored if None/foreign)

        if state.log_prob is not None and np.shape(state.log_prob) != (self.nwalkers,):
            raise ValueError("incompatible input dimensions")
        # upload; a missing log_prob is evaluated on the device (ensemble.py:350-358)
        self._engine.set_state(state.coords, state.log_prob)

        if thin is not None:  # deprecated form: store every `thin`-th, yield every step
            thin = int(thin)
            if thin <= 0:
                raise ValueError("Invalid thinning argument")
            yield_step, checkpoint_step = 1, thin
            if store:
                self.backend.grow(iterations // checkpoint_step, None)
        else:
            thin_by = int(thin_by)
            if thin_by <= 0:
                raise ValueError("Invalid thinning argument")
            yield_step = checkpoint_step = thin_by
            if store:
                self.backend.grow(iterations, None)

        native_store = store and type(self.backend) is Backend
        sched = self._schedule()
        eng = self._engine

        def refresh():
            bufs = self._pinned if self._pinned is not None else (
                np.empty((self.nwalkers, self.ndim)), np.empty(self.nwalkers))
            if self._rdv is not None and not self._gather_results:
                r0, n = eng.owned_rows()  # sharded: only the owned block crosses PCIe
                state.coords, state.log_prob = eng.get_state_rows(r0, n, *bufs)
            else:
                state.coords, state.log_prob = eng.get_state(*bufs)
            state.random_state = self.random_state

        if _bulk and iterations is not None and (not store or (native_store and thin is None)):
            # run_mcmc: the whole run is one C-ABI call
            total = iterations * yield_step
            if total > 0:
                if store:
                    b = self.backend
                    k0, k1 = b.iteration, b.iteration + iterations
                    eng.step_store(sched, total, checkpoint_step, b.chain[k0:k1], b.log_prob[k0:k1], b.accepted)
                    self._after_steps()
                    b.iteration = k1
                    b.random_state = self.random_state
                else:
                    eng.step(sched, total, want_accepted=False)
                    self._after_steps()
            refresh()
            if pbar is not None:
                pbar.update(iterations)
                pbar.close()
            if iterations > 0:
                yield state
            return

        i = 0
        counter = iter(int, 1) if iterations is None else range(iterations)
        for _ in counter:
            # the steps of this yield window; at most the last one is stored
            sched = self._schedule()  # stateful moves (GaussianMove "sequential") change between calls
            last_is_checkpoint = store and (i + yield_step) % checkpoint_step == 0
            if last_is_checkpoint and native_store:
                b = self.backend
                k = b.iteration
                eng.step_store(sched, yield_step, yield_step, b.chain[k : k + 1], b.log_prob[k : k + 1], b.accepted)
                self._after_steps()
                b.iteration = k + 1
                b.random_state = self.random_state
                refresh()
            else:
                accepted = eng.step(sched, yield_step, want_accepted=last_is_checkpoint)
                self._after_steps()
                refresh()
                if last_is_checkpoint:
                    self.backend.save_step(state, accepted)
            i += yield_step
            if pbar is not None:
                pbar.update(1)
            yield state
        if pbar is not None:
            pbar.close()

    def run_mcmc(self, initial_state, nsteps, **kwargs):
        """Iterate :func:`sample` for ``nsteps`` iterations and return the last
        state (``ensemble.py:426-456``); ``initial_state=None`` resumes."""
        if initial_state is None:
            if self._previous_state is None:
                raise ValueError("Cannot have `initial_state=None` if run_mcmc has never been called.")
            initial_state = self._previous_state
        results = None
        for results in self.sample(initial_state, iterations=nsteps, _bulk=True, **kwargs):
            pass
        self._previous_state = results
        return results

    def _walkers_independent(self, coords):
        """``walkers_independent`` (``ensemble.py:653-663``) with the O(N D^2) part on
        the device: ``eb_walkers_gram`` returns ``C^T C`` of the centred,
        column-normalised walkers, the host solves the ``D x D`` symmetric
        eigen-problem, ``cond(C) = sqrt(l_max / l_min)``.  Squaring halves the
        digits, so only a clearly well-conditioned ensemble (cond <= 1e6 by the
        Gram matrix) is accepted here; anything nearer the reference's 1e8
        threshold is decided by the reference's own SVD statement on the host."""
        coords = np.asarray(coords, dtype=np.float64)
        if coords.ndim != 2 or coords.shape[1] != self.ndim or self.ndim > 1024:
            return walkers_independent(coords)
        gram, flags = self._engine.walkers_gram(coords)
        if flags:
            return False  # non-finite coordinate or a column without span
        ev = np.linalg.eigvalsh(gram)
        if ev[0] > 0 and np.sqrt(ev[-1] / ev[0]) <= 1e6:
            return True
        return walkers_independent(coords)

    def compute_log_prob(self, coords):
        """``(log_prob, None)`` for ``coords[..., ndim]`` evaluated on the device
        (``ensemble.py:458-553``); raises ``ValueError`` for non-finite
        parameters or a NaN log-probability like the reference."""
        return self._engine.compute_log_prob(np.asarray(coords, dtype=np.float64)), None

    # ---------------------------------------------------------------- results
    @property
    def acceptance_fraction(self):
        return self.backend.accepted / float(self.backend.iteration)

    def get_chain(self, **kwargs):
        return self.get_value("chain", **kwargs)

    def get_blobs(self, **kwargs):
        return self.get_value("blobs", **kwargs)

    def get_log_prob(self, **kwargs):
        return self.get_value("log_prob", **kwargs)

    def get_last_sample(self, **kwargs):
        return self.backend.get_last_sample()

    def get_value(self, name, **kwargs):
        return self.backend.get_value(name, **kwargs)

    def get_autocorr_time(self, discard=0, thin=1, **kwargs):
        """Integrated autocorrelation time of the stored chain (``ensemble.py:619-623``
        -> ``backends/backend.py:130-150``), the FFTs on the GPU (``eb_autocorr``)."""
        from . import autocorr

        x = self.get_chain(discard=discard, thin=thin)
        return thin * autocorr.integrated_time(x, engine=self._engine, **kwargs)


def walkers_independent(coords):
    """Initial-state sanity check (``ensemble.py:653-663``): the centred,
    column-normalised walker matrix must have condition number <= 1e8.  Runs
    once per ``sample`` call on the host."""
    coords = np.asarray(coords, dtype=np.float64)
    if not np.all(np.isfinite(coords)):
        return False
    centred = coords - np.mean(coords, axis=0)[None, :]
    span = np.amax(np.abs(centred), axis=0)
    if np.any(span == 0):
        return False
    centred /= span
    centred /= np.sqrt(np.sum(centred**2, axis=0))
    return np.linalg.cond(centred) <= 1e8
