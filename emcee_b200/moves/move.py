"""Move protocol of the device path.

Reference interface: ``src/emcee/moves/move.py:8-45`` -- ``tune(state, accepted)`` and
``update(old_state, new_state, accepted, subset=None)``.  On the device path both
happen inside the fused half-step kernel (the accepted rows are written in place and
red-blue moves have nothing to tune); the host versions below exist so that code written
against the protocol -- e.g. a user-defined host move operating on ``State`` objects --
keeps working."""

import numpy as np

__all__ = ["Move"]


class Move(object):
    def tune(self, state, accepted):
        """Nothing to adapt for the red-blue moves."""
        return None

    def update(self, old_state, new_state, accepted, subset=None):
        """Scatter the accepted proposals of ``new_state`` (one row per member of
        ``subset``, in ascending walker order) into ``old_state`` and return it."""
        nwalkers = len(old_state.coords)
        members = np.arange(nwalkers) if subset is None else np.flatnonzero(subset)
        won = np.asarray(accepted, dtype=bool)[members]  # per proposal row
        dst, src = members[won], np.flatnonzero(won)
        old_state.coords[dst] = new_state.coords[src]
        old_state.log_prob[dst] = new_state.log_prob[src]
        if new_state.blobs is not None:
            if old_state.blobs is None:
                raise ValueError("the proposal carries blobs but the current state has none")
            old_state.blobs[dst] = new_state.blobs[src]
        return old_state
