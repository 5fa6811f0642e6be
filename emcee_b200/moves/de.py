"""Differential-evolution move (reference: ``src/emcee/moves/de.py:11-77``)."""

from .red_blue import RedBlueMove

__all__ = ["DEMove"]


class DEMove(RedBlueMove):
    """Args mirror ``de.py:28``: ``sigma`` (default 1e-5) and ``gamma0``
    (default ``2.38 / sqrt(2 ndim)``, resolved by the engine, ``de.py:33-38``)."""

    kind = "de"

    def __init__(self, sigma=1.0e-5, gamma0=None, **kwargs):
        self.sigma = sigma
        self.gamma0 = gamma0
        super().__init__(**kwargs)

    def _params(self):
        return float(self.sigma), float("nan") if self.gamma0 is None else float(self.gamma0)
