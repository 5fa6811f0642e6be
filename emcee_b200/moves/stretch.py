"""Goodman & Weare stretch move (reference: ``src/emcee/moves/stretch.py:11-33``)."""

from .red_blue import RedBlueMove

__all__ = ["StretchMove"]


class StretchMove(RedBlueMove):
    """:param a: the stretch scale parameter (default ``2.0``, ``stretch.py:22``)."""

    kind = "stretch"

    def __init__(self, a=2.0, **kwargs):
        self.a = a
        super(StretchMove, self).__init__(**kwargs)

    def _params(self):
        return float(self.a), float("nan")
