"""Build the CUDA-side sampler for a golden case / an oracle configuration."""
import numpy as np

import emcee_b200
from emcee_b200 import models, moves


def device_model(kind, g=None, target=None):
    if kind == "gauss_iso":
        return models.GaussianIso()
    if kind == "gauss_dense":
        if g is not None:
            return models.GaussianDense(g["model_icov"], g["model_mean"])
        return models.GaussianDense(target.icov, target.mean)
    if kind == "rosenbrock":
        p = g["model_params"] if g is not None else (target.a, target.b)
        return models.Rosenbrock(*p)
    if kind == "ring":
        p = g["model_params"] if g is not None else (target.radius, target.sigma)
        return models.Ring(*p)
    raise ValueError(kind)


GAUSS_MODES = ("vector", "random", "sequential")


def device_moves(rows, g=None):
    out = []
    for k, (kind, w, nsplits, rand, p0, p1) in enumerate(rows):
        kw = dict(randomize_split=bool(rand))
        if kind == 0:
            m = moves.StretchMove(a=p0, nsplits=int(nsplits), **kw)
        elif kind == 1:
            m = moves.DEMove(sigma=p0, gamma0=None if np.isnan(p1) else p1, nsplits=int(nsplits), **kw)
        elif kind == 2:
            m = moves.DESnookerMove(gammas=p0, **kw)
        elif kind == 3:
            m = moves.WalkMove(s=None if np.isnan(p0) else int(p0), nsplits=int(nsplits), **kw)
        else:
            cov = g["move%d_cov" % k]
            m = moves.GaussianMove(cov if cov.ndim else float(cov), mode=GAUSS_MODES[int(p0)],
                                   factor=None if np.isnan(p1) else float(p1))
        out.append((m, w))
    return out


def golden_sampler(g):
    return emcee_b200.EnsembleSampler(
        int(g["nwalkers"]), int(g["ndim"]), device_model(str(g["model_kind"]), g=g),
        moves=device_moves(g["moves"], g), seed=int(g["seed"]),
    )


def move_rows_from_oracle(oracle_moves):
    rows = []
    for m, w in oracle_moves:
        if m.kind == "stretch":
            rows.append([0, w, m.nsplits, m.randomize_split, m.a, np.nan])
        elif m.kind == "de":
            rows.append([1, w, m.nsplits, m.randomize_split, m.sigma, np.nan if m.gamma0 is None else m.gamma0])
        else:
            rows.append([2, w, m.nsplits, m.randomize_split, m.gammas, np.nan])
    return np.array(rows, dtype=np.float64)
