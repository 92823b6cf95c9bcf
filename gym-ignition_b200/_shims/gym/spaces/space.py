from ..utils import seeding


class Space:
    def __init__(self, shape=None, dtype=None):
        import numpy as np
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self.np_random = None
        self.seed()

    def sample(self):
        raise NotImplementedError

    def seed(self, seed=None):
        self.np_random, seed = seeding.np_random(seed)
        return [seed]

    def contains(self, x):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)
