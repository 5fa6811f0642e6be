"""Snooker differential-evolution move (reference:
``src/emcee/moves/de_snooker.py:10-46``)."""

from .red_blue import RedBlueMove

__all__ = ["DESnookerMove"]


class DESnookerMove(RedBlueMove):
    """:param gammas: mean stretch factor (default ``1.7``, ``de_snooker.py:26``).
    Always uses four sub-ensembles (``de_snooker.py:28``)."""

    kind = "snooker"

    def __init__(self, gammas=1.7, **kwargs):
        self.gammas = gammas
        kwargs["nsplits"] = 4
        super(DESnookerMove, self).__init__(**kwargs)

    def _params(self):
        return float(self.gammas), float("nan")
