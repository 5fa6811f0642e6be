"""In-memory chain store with the reference's ``Backend`` protocol
(``src/emcee/backends/backend.py:12-237``): ``reset / grow / save_step /
get_chain / get_log_prob / get_last_sample / shape / iteration / accepted /
random_state``.  Blobs do not exist on the device path.

The engine's ``eb_step_store`` writes stored steps straight into the
``chain`` / ``log_prob`` arrays of this class (pinned double-buffered D2H), so
``store=True`` does not need a host round trip per step."""

import numpy as np

from .state import State

__all__ = ["Backend"]


class Backend(object):
    def __init__(self, dtype=None):
        self.initialized = False
        self.dtype = np.float64 if dtype is None else dtype
        if np.dtype(self.dtype) != np.float64:
            raise NotImplementedError("the device path stores float64 chains only")

    def reset(self, nwalkers, ndim):
        self.nwalkers, self.ndim = int(nwalkers), int(ndim)
        self.iteration = 0
        self.accepted = np.zeros(self.nwalkers, dtype=self.dtype)  # backend.py:31
        self.chain = np.empty((0, self.nwalkers, self.ndim), dtype=self.dtype)
        self.log_prob = np.empty((0, self.nwalkers), dtype=self.dtype)
        self.blobs = None
        self.random_state = None
        self.initialized = True

    def has_blobs(self):
        return False

    @property
    def shape(self):
        return self.nwalkers, self.ndim

    # -- growth / writes ------------------------------------------------------
    def grow(self, ngrow, blobs):
        """Room for ``ngrow`` more stored steps (``backend.py:164-185``)."""
        if blobs is not None:
            raise NotImplementedError("blobs are not supported on the device path")
        extra = int(ngrow) - (len(self.chain) - self.iteration)
        if extra <= 0:
            return
        total = len(self.chain) + extra
        chain = np.empty((total, self.nwalkers, self.ndim), dtype=self.dtype)
        chain[: len(self.chain)] = self.chain
        log_prob = np.empty((total, self.nwalkers), dtype=self.dtype)
        log_prob[: len(self.log_prob)] = self.log_prob
        self.chain, self.log_prob = chain, log_prob

    def save_step(self, state, accepted):
        """Append one step (``backend.py:214-231``)."""
        if state.coords.shape != self.shape:
            raise ValueError("invalid coordinate dimensions; expected {0}".format(self.shape))
        if state.log_prob.shape != (self.nwalkers,):
            raise ValueError("invalid log probability size; expected {0}".format(self.nwalkers))
        if accepted.shape != (self.nwalkers,):
            raise ValueError("invalid acceptance size; expected {0}".format(self.nwalkers))
        if state.blobs is not None:
            raise ValueError("unexpected blobs")
        self.chain[self.iteration] = state.coords
        self.log_prob[self.iteration] = state.log_prob
        self.accepted += accepted
        self.random_state = state.random_state
        self.iteration += 1

    # -- reads ---------------------------------------------------------------
    def get_value(self, name, flat=False, thin=1, discard=0):
        if self.iteration <= 0:
            raise AttributeError(
                "you must run the sampler with 'store == True' before accessing the results"
            )
        if name == "blobs":
            return None
        v = getattr(self, name)[discard + thin - 1 : self.iteration : thin]  # backend.py:53
        if flat:
            return v.reshape((v.shape[0] * v.shape[1],) + v.shape[2:])
        return v

    def get_chain(self, **kwargs):
        """``[nsteps, nwalkers, ndim]`` (or flattened over walkers)."""
        return self.get_value("chain", **kwargs)

    def get_log_prob(self, **kwargs):
        return self.get_value("log_prob", **kwargs)

    def get_blobs(self, **kwargs):
        return self.get_value("blobs", **kwargs)

    def get_last_sample(self):
        if (not self.initialized) or self.iteration <= 0:
            raise AttributeError(
                "you must run the sampler with 'store == True' before accessing the results"
            )
        k = self.iteration - 1
        return State(self.chain[k], log_prob=self.log_prob[k], blobs=None, random_state=self.random_state)

    def get_autocorr_time(self, discard=0, thin=1, **kwargs):
        """Integrated autocorrelation time per parameter, in steps
        (``backend.py:130-150``)."""
        from . import autocorr

        x = self.get_chain(discard=discard, thin=thin)
        return thin * autocorr.integrated_time(x, **kwargs)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        pass
