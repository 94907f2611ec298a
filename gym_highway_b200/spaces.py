"""Minimal gym.spaces stand-ins (gym is not a dependency): just what trainers read from a VecEnv —
`shape`, `dtype`, `low`, `high`, `n`, `sample()`, `contains()`.  Box casts its bounds to `dtype`
like gym does (the single-agent observation bounds are float32, highway_env.py:51)."""
import numpy as np


class Space(object):
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)

    def __repr__(self):
        return "%s(shape=%s, dtype=%s)" % (type(self).__name__, self.shape, self.dtype)


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        Space.__init__(self, (), np.int64)

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __eq__(self, o):
        return isinstance(o, Discrete) and o.n == self.n


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        dtype = np.dtype(dtype)
        if shape is None:
            low, high = np.asarray(low), np.asarray(high)
            shape = low.shape
        else:
            low = np.full(shape, low, dtype=np.float64)
            high = np.full(shape, high, dtype=np.float64)
        self.low = low.astype(dtype)
        self.high = high.astype(dtype)
        Space.__init__(self, shape, dtype)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return np.random.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __eq__(self, o):
        return isinstance(o, Box) and np.array_equal(o.low, self.low) and np.array_equal(o.high, self.high)


def highway_obs_bounds(num_vehicles, dtype=np.float64):
    """low/high of the observation vector for `num_vehicles` vehicles (ego + others):
    [acc, steer, raw_x, y, (dx, dy) x (V-1), (vx, vy) x V]  — highway_env.py:33-49,
    multiagent_envs/highway_env.py:34-48."""
    v = int(num_vehicles)
    low = np.concatenate(([-5.0, -30.0], [0.0, 0.0] + [-60.0, -8.0] * (v - 1), [-20.0, -20.0] * v))
    high = np.concatenate(([5.0, 30.0], [1200.0, 8.0] + [60.0, 8.0] * (v - 1), [20.0, 20.0] * v))
    return low.astype(dtype), high.astype(dtype)
