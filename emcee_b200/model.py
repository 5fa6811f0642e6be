"""The 4-tuple a move receives (reference: ``src/emcee/model.py:8-10``)."""

from collections import namedtuple

__all__ = ["Model"]

# log_prob_fn: the registered device model; compute_log_prob_fn: the sampler's
# batched evaluator (C ABI eb_compute_log_prob); map_fn: unused on the device
# path (kept for signature parity); random: the sampler's DeviceRandom.
Model = namedtuple("Model", ("log_prob_fn", "compute_log_prob_fn", "map_fn", "random"))
