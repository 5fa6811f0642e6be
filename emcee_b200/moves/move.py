"""Move protocol (reference: ``src/emcee/moves/move.py:8-45``)."""

import numpy as np

__all__ = ["Move"]


class Move(object):
    def tune(self, state, accepted):
        """No-op for every red-blue move (``move.py:9-10``)."""
        pass

    def update(self, old_state, new_state, accepted, subset=None):
        """Masked scatter of accepted proposals into a host ``State``
        (``move.py:12-45``).  The engine performs this update in the fused
        half-step kernel; the method exists for host-side callers that build
        their own moves on the ``State`` container."""
        n = len(old_state.coords)
        subset = np.ones(n, dtype=bool) if subset is None else subset
        take = accepted[subset]
        put = subset & accepted
        old_state.coords[put] = new_state.coords[take]
        old_state.log_prob[put] = new_state.log_prob[take]
        if new_state.blobs is not None:
            if old_state.blobs is None:
                raise ValueError(
                    "If you start sampling with a given log_prob, "
                    "you also need to provide the current list of "
                    "blobs at that position."
                )
            old_state.blobs[put] = new_state.blobs[take]
        return old_state
