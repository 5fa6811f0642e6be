"""Moves of the device path (reference: ``src/emcee/moves/__init__.py``).

The red-blue family (``StretchMove``, ``DEMove``, ``DESnookerMove``, ``WalkMove``) and the
Metropolis family with Gaussian proposals (``MHMove``, ``GaussianMove``); the reference's
``KDEMove`` (SciPy kernel-density proposals) is out of scope."""

from .de import DEMove
from .de_snooker import DESnookerMove
from .gaussian import GaussianMove
from .mh import MHMove
from .move import Move
from .red_blue import RedBlueMove
from .stretch import StretchMove
from .walk import WalkMove

__all__ = ["Move", "RedBlueMove", "StretchMove", "DEMove", "DESnookerMove", "WalkMove", "MHMove", "GaussianMove"]
