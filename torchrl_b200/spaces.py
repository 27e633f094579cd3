"""Observation / action space descriptors (the two gym.spaces the reference reads).

The reference tests ``isinstance(env.action_space, gym.spaces.Box)``
(/root/reference/torchrl/collector/base.py:26, algo/rl_algo.py:35); the product keeps its
own tiny classes so that it does not depend on gym being installed.
"""
import numpy as np


class Space:
    shape = ()


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            low = np.asarray(low, dtype=np.float64)
            high = np.asarray(high, dtype=np.float64)
            shape = low.shape
        else:
            low = np.full(shape, low, dtype=np.float64)
            high = np.full(shape, high, dtype=np.float64)
        self.low, self.high = low, high
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)

    def __repr__(self):
        return "Box%s" % (self.shape,)


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def __repr__(self):
        return "Discrete(%d)" % self.n


def is_box(space):
    """True for our Box and for any gym-like Box (has low/high and is not discrete)."""
    return hasattr(space, "low") and hasattr(space, "high") and not hasattr(space, "n")
