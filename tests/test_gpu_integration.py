"""INTEGRATION.md section 4, executed: the UNMODIFIED reference sampler (``baseline/_ref/emcee_reference.zip``)
drives the engine through the reference-side ctypes binding ``tests/helpers/reference_b200_move.py`` -- its own
``EnsembleSampler.sample`` loop, its own ``Backend``, our ``propose`` and ``log_prob_fn``.  The chain must equal
the oracle's for the same Philox key, bit for bit (stretch move)."""
import os
import sys

import numpy as np
import pytest

from oracle import redblue as rb
from oracle import targets as T

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ZIP = os.path.join(ROOT, "baseline", "_ref", "emcee_reference.zip")


@pytest.mark.skipif(not os.path.exists(REF_ZIP), reason="baseline/_ref/emcee_reference.zip not packaged (baseline/make_ref.py)")
def test_reference_sampler_drives_the_engine():
    sys.path.insert(0, REF_ZIP)
    sys.path.insert(0, os.path.join(ROOT, "tests", "helpers"))
    try:
        import emcee  # the reference package

        assert REF_ZIP in emcee.__file__
        import reference_b200_move as binding

        N, D, steps, seed = 512, 16, 25, 0x1B200
        target, p0 = T.make_config("gauss_dense", N, D)
        model = binding.DeviceGaussian(N, target.icov, seed=seed)
        sampler = emcee.EnsembleSampler(N, D, model, moves=binding.B200StretchMove(), vectorize=True)
        state = sampler.run_mcmc(p0, steps, skip_initial_state_check=True)  # the reference's own loop and Backend
        o = rb.OracleSampler(N, D, target, [(rb.Stretch(), 1.0)], seed=seed)
        o.set_state(p0)
        acc_total = np.zeros(N)
        for k in range(steps):
            acc_total += o.run(1)
            assert np.array_equal(sampler.get_chain()[k], o.coords), k
        assert np.array_equal(state.coords, o.coords)
        np.testing.assert_allclose(state.log_prob, o.log_prob, rtol=1e-12, atol=1e-12)
        assert np.array_equal(sampler.backend.accepted, acc_total)
        assert 0.1 < sampler.acceptance_fraction.mean() < 0.9
        # the reference's guards still fire through the binding (red_blue.py:64-70)
        few = emcee.EnsembleSampler(8, D, binding.DeviceGaussian(8, target.icov), moves=binding.B200StretchMove(),
                                    vectorize=True)
        with pytest.raises(RuntimeError):
            few.run_mcmc(p0[:8], 1, skip_initial_state_check=True)
        model.close()
    finally:
        sys.path.remove(REF_ZIP)
        sys.path.remove(os.path.join(ROOT, "tests", "helpers"))
        for name in [m for m in sys.modules if m == "emcee" or m.startswith("emcee.")]:
            del sys.modules[name]
