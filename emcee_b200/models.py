"""Registered device-side log-probability models.

The reference takes an arbitrary Python callable ``log_prob_fn``
(``src/emcee/ensemble.py:79-83``) and evaluates it row by row or through
``pool.map`` (``ensemble.py:486-496``).  A GPU engine cannot call back into
Python per walker, so the drop-in takes a *registered model*: a small object
naming one of the log-probabilities compiled into the CUDA library plus its
parameters.  These classes only carry parameters -- they are deliberately not
callable, so no code path can silently evaluate a model on the host.
"""

import numpy as np

__all__ = ["DeviceModel", "GaussianIso", "GaussianDense", "Rosenbrock", "Ring"]


class DeviceModel(object):
    kind = None

    def device_params(self, ndim):
        """flat float64 parameter vector for ``eb_model_set``"""
        raise NotImplementedError

    def __call__(self, *a, **k):
        raise TypeError(
            "device models are evaluated by the CUDA engine "
            "(EnsembleSampler.compute_log_prob); they cannot be called on the host"
        )


class GaussianIso(DeviceModel):
    """``-0.5 * sum(x**2)`` (the target of the reference's move tests,
    ``tests/integration/test_proposal.py:21-22``)."""

    kind = "gauss_iso"

    def device_params(self, ndim):
        return np.zeros(0)


class GaussianDense(DeviceModel):
    """``-0.5 * (x - mean)^T icov (x - mean)`` with a dense precision matrix
    (``document/plots/oned.py:17-18``)."""

    kind = "gauss_dense"

    def __init__(self, icov, mean=None):
        self.icov = np.ascontiguousarray(icov, dtype=np.float64)
        if self.icov.ndim != 2 or self.icov.shape[0] != self.icov.shape[1]:
            raise ValueError("icov must be a square matrix")
        d = self.icov.shape[0]
        self.mean = np.zeros(d) if mean is None else np.ascontiguousarray(mean, dtype=np.float64)
        if self.mean.shape != (d,):
            raise ValueError("mean must have shape (ndim,)")

    def device_params(self, ndim):
        if self.icov.shape[0] != ndim:
            raise ValueError("icov is %dx%d but ndim = %d" % (self.icov.shape + (ndim,)))
        return np.concatenate([self.mean, self.icov.ravel()])


class Rosenbrock(DeviceModel):
    """``-sum_i [ b (x[i+1] - x[i]^2)^2 + (a - x[i])^2 ]``"""

    kind = "rosenbrock"

    def __init__(self, a=1.0, b=100.0):
        self.a, self.b = float(a), float(b)

    def device_params(self, ndim):
        return np.array([self.a, self.b])


class Ring(DeviceModel):
    """``-(|x| - radius)^2 / (2 sigma^2)``"""

    kind = "ring"

    def __init__(self, radius=5.0, sigma=0.5):
        self.radius, self.sigma = float(radius), float(sigma)

    def device_params(self, ndim):
        return np.array([self.radius, self.sigma])
