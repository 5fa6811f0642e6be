#!/usr/bin/env python
"""Write tests/fixtures/autocorr_reference.npz from the UNMODIFIED reference's
``emcee.autocorr`` (authoring container only): ``python -m oracle.gen_autocorr_fixture``."""
import os
import sys
import types

import numpy as np


def chain(seed, n, w, d):
    rng = np.random.default_rng(seed)
    x = np.empty((n, w, d))
    x[0] = 0
    e = rng.random((n, w, d))
    for i in range(1, n):
        x[i] = x[i - 1] * 0.9 + e[i]
    return x


def main():
    stub = types.ModuleType("emcee.emcee_version")
    stub.__version__ = "0+reference.8ab6c0f"
    sys.modules["emcee.emcee_version"] = stub
    sys.path.insert(0, "/root/reference/src")
    from emcee import autocorr as ref

    out = {}
    for name, (seed, n, w, d) in {"a": (1, 20000, 4, 3), "b": (2, 50000, 1, 2), "c": (3, 8000, 16, 1)}.items():
        out["tau_" + name] = ref.integrated_time(chain(seed, n, w, d), quiet=True)
        out["cfg_" + name] = np.array([seed, n, w, d])
    out["acf_head"] = ref.function_1d(chain(5, 3000, 2, 2)[:, 0, 0])[:16]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    np.savez(os.path.join(root, "tests", "fixtures", "autocorr_reference.npz"), **out)


if __name__ == "__main__":
    main()
