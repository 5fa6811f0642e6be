"""Host autocorrelation analysis vs values computed by the reference's
``emcee.autocorr`` on the same seeded AR(1) chains (fixture written in the
authoring container by ``oracle/gen_autocorr_fixture.py``; reference tests:
``tests/unit/test_autocorr.py``)."""
import os

import numpy as np
import pytest

from emcee_b200 import autocorr

FIX = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures", "autocorr_reference.npz"))


def chain(seed, n, w, d):
    rng = np.random.default_rng(seed)
    x = np.empty((n, w, d))
    x[0] = 0
    e = rng.random((n, w, d))
    for i in range(1, n):
        x[i] = x[i - 1] * 0.9 + e[i]
    return x


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_matches_reference(name):
    seed, n, w, d = (int(v) for v in FIX["cfg_" + name])
    tau = autocorr.integrated_time(chain(seed, n, w, d), quiet=True)
    np.testing.assert_allclose(tau, FIX["tau_" + name], rtol=1e-8)
    assert np.all(np.abs(tau - 19.0) / 19.0 < 0.25)  # AR(1), a = 0.9: tau = (1+a)/(1-a) = 19


def test_acf_and_shapes():
    x = chain(5, 3000, 2, 2)
    np.testing.assert_allclose(autocorr.function_1d(x[:, 0, 0])[:16], FIX["acf_head"], rtol=1e-9, atol=1e-12)
    t1 = autocorr.integrated_time(x[:, 0, :][:, None], quiet=True)
    t2 = autocorr.integrated_time(x[:, 0, :], has_walkers=False, quiet=True)
    assert np.allclose(t1, t2)
    with pytest.raises(autocorr.AutocorrError) as err:
        autocorr.integrated_time(chain(6, 100, 2, 3))
    assert err.value.tau.shape == (3,)
    with pytest.raises(ValueError):
        autocorr.function_1d(np.zeros((3, 3)))
