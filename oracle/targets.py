"""Synthetic targets of BASELINE.json's configs, as numpy (TEST INFRASTRUCTURE ONLY).

The reference ships no benchmark targets (grep finds neither
"rosenbrock" nor "ring"), so they are defined once here and in DESIGN.md.
Every ``log_prob`` below is *vectorised*: ``coords[M, D] -> float64[M]`` as
``EnsembleSampler(..., vectorize=True)`` expects (``ensemble.py:486-487``), and
also accepts a single ``[D]`` row (the per-walker ``map`` form,
``ensemble.py:492-496``).

* gauss_iso   : ``-0.5 * sum(x**2)``            (``tests/integration/test_proposal.py:21-22``)
* gauss_dense : ``-0.5 * (x-mu)^T icov (x-mu)`` (``document/plots/oned.py:17-18``),
                icov from ``random_cov`` (``document/plots/oned.py:21-25``)
* rosenbrock  : ``-sum_i [ b (x[i+1]-x[i]^2)^2 + (a-x[i])^2 ]``, a=1, b=100
* ring        : ``-(|x| - R)^2 / (2 sigma^2)``
"""

import numpy as np

__all__ = [
    "GaussIso",
    "GaussDense",
    "Rosenbrock",
    "Ring",
    "random_cov",
    "make_config",
]

MODEL_SEED = 20240
INIT_SEED = 20241
SAMPLER_SEED = 0x656D636565B200


class GaussIso(object):
    kind = "gauss_iso"

    def __init__(self, ndim):
        self.ndim = int(ndim)

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        return -0.5 * np.sum(x * x, axis=-1)


class GaussDense(object):
    kind = "gauss_dense"

    def __init__(self, icov, mean=None):
        self.icov = np.ascontiguousarray(icov, dtype=np.float64)
        self.ndim = self.icov.shape[0]
        self.mean = (
            np.zeros(self.ndim) if mean is None else np.asarray(mean, dtype=np.float64)
        )

    def __call__(self, x):
        d = np.asarray(x, dtype=np.float64) - self.mean
        return -0.5 * np.sum((d @ self.icov) * d, axis=-1)


class Rosenbrock(object):
    kind = "rosenbrock"

    def __init__(self, ndim, a=1.0, b=100.0):
        self.ndim, self.a, self.b = int(ndim), float(a), float(b)

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        x0, x1 = x[..., :-1], x[..., 1:]
        t = x1 - x0 * x0
        u = self.a - x0
        return -np.sum(self.b * (t * t) + u * u, axis=-1)


class Ring(object):
    kind = "ring"

    def __init__(self, ndim, radius=5.0, sigma=0.5):
        self.ndim, self.radius, self.sigma = int(ndim), float(radius), float(sigma)

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        r = np.sqrt(np.sum(x * x, axis=-1))
        d = r - self.radius
        return -(d * d) / (2.0 * self.sigma * self.sigma)


def random_cov(ndim, dof=1, rng=None):
    """The PASP paper's covariance generator (``document/plots/oned.py:21-25``):
    ``sum_i v_i v_i^T / (ndim + dof)`` over ``ndim + dof`` standard-normal
    vectors, drawn here from a seeded Generator instead of the global state."""
    rng = np.random.default_rng(MODEL_SEED) if rng is None else rng
    v = rng.standard_normal((ndim + dof, ndim))
    return (v.T @ v) / (ndim + dof)


def make_config(name, nwalkers=None, ndim=None, dof=None):
    """(target, p0) of one BASELINE.json config, optionally rescaled.

    name: "gauss_iso" | "gauss_dense" | "rosenbrock" | "ring".
    gauss_dense uses the paper's ``random_cov(ndim, dof=1)``: at
    D = 128 the covariance has condition number ~3e4, so different fp64
    summation orders of the quadratic form agree to ~1e-12 relative.
    """
    defaults = {
        "gauss_iso": (32, 5),
        "gauss_dense": (4096, 128),
        "rosenbrock": (16384, 256),
        "ring": (262144, 32),
    }
    n0, d0 = defaults[name]
    n = int(nwalkers or n0)
    d = int(ndim or d0)
    rng_m = np.random.default_rng(MODEL_SEED)
    rng_p = np.random.default_rng(INIT_SEED)
    if name == "gauss_iso":
        return GaussIso(d), rng_p.standard_normal((n, d))
    if name == "gauss_dense":
        cov = random_cov(d, dof=1 if dof is None else dof, rng=rng_m)
        icov = np.linalg.inv(cov)
        icov = 0.5 * (icov + icov.T)
        return GaussDense(icov), rng_p.standard_normal((n, d))
    if name == "rosenbrock":
        return Rosenbrock(d), 1.0 + 0.1 * rng_p.standard_normal((n, d))
    if name == "ring":
        t = Ring(d)
        return t, rng_p.standard_normal((n, d)) * (t.radius / np.sqrt(d))
    raise ValueError(name)
