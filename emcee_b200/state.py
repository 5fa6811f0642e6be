"""Host mirror of the ensemble state (reference: ``src/emcee/state.py:10-75``).

The live walker positions and log-probabilities stay in HBM inside the engine;
a ``State`` is the host-side snapshot handed to / received from the user, with
the reference's attribute names, copy semantics and tuple-unpacking
back-compat."""

import copy as _copy

import numpy as np

__all__ = ["State"]


class State(object):
    """Snapshot of the ensemble: ``coords[nwalkers, ndim]``, ``log_prob[nwalkers]``,
    ``blobs`` (always ``None`` on the device path) and ``random_state``.

    Iterating yields ``coords, log_prob, random_state`` (plus ``blobs`` when
    present), as the reference does for pre-3.0 callers (``state.py:47-75``)."""

    __slots__ = ("coords", "log_prob", "blobs", "random_state")

    def __init__(self, coords, log_prob=None, blobs=None, random_state=None, copy=False):
        other = coords if hasattr(coords, "coords") else None  # state.py:35-40
        if other is not None:
            fields = (other.coords, other.log_prob, other.blobs, other.random_state)
        else:
            fields = (np.atleast_2d(coords), log_prob, blobs, random_state)  # state.py:42
        if copy:
            fields = tuple(_copy.deepcopy(f) for f in fields)
        self.coords, self.log_prob, self.blobs, self.random_state = fields

    def _as_tuple(self):
        head = (self.coords, self.log_prob, self.random_state)
        return head if self.blobs is None else head + (self.blobs,)

    def __len__(self):
        return len(self._as_tuple())

    def __iter__(self):
        return iter(self._as_tuple())

    def __getitem__(self, index):
        items = self._as_tuple()
        if not -len(items) <= index < len(items):
            raise IndexError("Invalid index '{0}'".format(index))
        return items[index]

    def __repr__(self):
        return "State({0}, log_prob={1}, blobs={2}, random_state={3})".format(
            self.coords, self.log_prob, self.blobs, self.random_state
        )
