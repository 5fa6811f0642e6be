"""Goodman & Weare walk move (reference: ``src/emcee/moves/walk.py:10-37``)."""

from .red_blue import RedBlueMove

__all__ = ["WalkMove"]


class WalkMove(RedBlueMove):
    """:param s: the number of helper walkers (``walk.py:20-22``); ``None`` (default) uses the whole
        complement, whose covariance is then computed once per split on the FP64 tensor pipe.

    Every proposal is ``N(x_k, cov(helpers))`` (``walk.py:34-36``); the draw specification fixes
    ``multivariate_normal(mean, cov) := mean + L z`` with ``L`` the lower Cholesky factor of ``cov`` (zero
    columns for the null space of a rank-deficient covariance).  With a helper subset each walker's
    covariance is factorised by one thread block: ``ndim <= 64`` and ``s <= 4096``."""

    kind = "walk"

    def __init__(self, s=None, **kwargs):
        self.s = s
        super(WalkMove, self).__init__(**kwargs)

    def _params(self):
        return float("nan") if self.s is None else float(int(self.s)), float("nan")
