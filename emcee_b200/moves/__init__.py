"""Moves of the device path (reference: ``src/emcee/moves/__init__.py``).

Only the red-blue family named by the hot path is provided; the reference's
``MHMove``, ``GaussianMove``, ``WalkMove`` and ``KDEMove`` are out of scope
(DESIGN.md)."""

from .de import DEMove
from .de_snooker import DESnookerMove
from .move import Move
from .red_blue import RedBlueMove
from .stretch import StretchMove

__all__ = ["Move", "RedBlueMove", "StretchMove", "DEMove", "DESnookerMove"]
