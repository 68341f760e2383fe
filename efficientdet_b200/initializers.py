"""PriorProbability -- mirror of the reference's initializers.py:6-31."""
import math

import numpy as np


class PriorProbability:
    """Constant bias -log((1-p)/p) for the classification head's last conv (focal-loss prior)."""

    def __init__(self, probability=0.01):
        self.probability = probability

    def __call__(self, shape, dtype=None):
        scalar = -math.log((1 - self.probability) / self.probability)
        return np.full(shape, scalar, dtype=dtype or np.float32)

    def get_config(self):
        return {"probability": self.probability}
