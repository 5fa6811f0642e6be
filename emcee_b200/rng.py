"""The sampler's random state on the device path.

The reference owns a private MT19937 ``RandomState`` (``ensemble.py:166-167``)
and snapshots it every step (``ensemble.py:410``).  The engine is counter
based: every draw is a pure function of ``(seed, step, split, active rank,
purpose)`` through Philox4x32-10 ("Draw specification"), so the
whole random state is the pair ``(seed, step)`` and resuming a chain is exact.
``DeviceRandom`` is the object stored as ``sampler._random`` and handed to
moves as ``model.random``; it exposes ``get_state`` / ``set_state`` like the
object it replaces, bound to the engine that consumes the draws.
"""

__all__ = ["DeviceRandom", "STATE_TAG"]

STATE_TAG = "philox4x32-10"


class DeviceRandom(object):
    def __init__(self, engine):
        self.engine = engine

    def get_state(self):
        seed, step = self.engine.get_rng()
        return (STATE_TAG, seed, step)

    def set_state(self, state):
        tag, seed, step = state  # raises on None / foreign states, like RandomState.set_state
        if tag != STATE_TAG:
            raise ValueError("not a %s state" % STATE_TAG)
        self.engine.set_rng(int(seed), int(step))
