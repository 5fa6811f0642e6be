"""Metropolis step with a Gaussian proposal (reference: ``src/emcee/moves/gaussian.py:10-119``)."""

import numpy as np

from .mh import MHMove

__all__ = ["GaussianMove"]

_MODES = ("vector", "random", "sequential")


class GaussianMove(MHMove):
    """Args mirror ``gaussian.py:32``: ``cov`` (scalar, vector or matrix: isotropic, axis-aligned or general
    proposal), ``mode`` (``"vector"``, ``"random"``, ``"sequential"``) and ``factor`` (a per-step scale drawn
    log-uniformly in ``[1/factor, factor]``).  Raises the reference's ``ValueError`` s for bad arguments."""

    def __init__(self, cov, mode="vector", factor=None):
        try:
            float(cov)
        except TypeError:
            cov = np.atleast_1d(np.asarray(cov, dtype=np.float64))
            if cov.ndim == 1:
                ndim, allowed = len(cov), _MODES  # gaussian.py:43-45
            elif cov.ndim == 2 and cov.shape[0] == cov.shape[1]:
                ndim, allowed = cov.shape[0], ("vector",)  # gaussian.py:47-50,110
            else:
                raise ValueError("Invalid proposal scale dimensions")
        else:
            cov = np.array([float(cov)])
            ndim, allowed = None, _MODES  # gaussian.py:56-58
        if factor is not None and factor < 1.0:
            raise ValueError("'factor' must be >= 1.0")  # gaussian.py:69-70
        if mode not in allowed:
            raise ValueError(
                "'{0}' is not a recognized mode. Please select from: {1}".format(mode, list(allowed))
            )  # gaussian.py:73-79
        proposal = dict(family="gaussian", cov=np.ascontiguousarray(cov, dtype=np.float64), mode=_MODES.index(mode),
                        factor=factor)
        super(GaussianMove, self).__init__(proposal, ndim=ndim)
