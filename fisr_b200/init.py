"""Parameter initialisation of the reference (ops.py:8-9) for runs without a checkpoint (random-init benchmarks,
``--phase train`` from scratch)."""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

from .engine import param_inventory


def xavier_params(seed: int = 0, bias_std: float = 0.0) -> "OrderedDict[str, np.ndarray]":
    """``tf.contrib.layers.xavier_initializer(uniform=False)``: truncated normal (+-2 sigma), sigma = sqrt(1.3 / n),
    n = (fan_in + fan_out) / 2 with fan = 9 * C (TF-1.13); biases ``constant_initializer(0)`` unless bias_std > 0."""
    rng = np.random.default_rng(seed)
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, shape in param_inventory().items():
        if name.endswith("/w"):
            n = (9 * shape[2] + 9 * shape[3]) / 2.0
            sigma = math.sqrt(1.3 / n)
            w = rng.standard_normal(shape)
            bad = np.abs(w) > 2.0
            while bad.any():                               # resample the tails like tf.truncated_normal
                w[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(w) > 2.0
            out[name] = (w * sigma).astype(np.float32)
        else:
            out[name] = (rng.standard_normal(shape) * bias_std).astype(np.float32)
    return out
