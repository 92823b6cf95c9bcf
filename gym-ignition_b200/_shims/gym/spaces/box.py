import numpy as np

from .space import Space


class Box(Space):
    """gym 0.17 Box: bounds are stored in ``dtype``; ``contains`` compares the input as given (a float64
    vector is compared against the float32-rounded bounds, inclusive) — the reference's done masks rely on it."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        assert dtype is not None, "dtype must be explicitly provided."
        self.dtype = np.dtype(dtype)
        if shape is None:
            assert np.shape(low) == np.shape(high), "box dimension mismatch."
            self.shape = np.shape(low)
            self.low, self.high = np.asarray(low), np.asarray(high)
        else:
            assert np.isscalar(low) and np.isscalar(high), "box requires scalar bounds."
            self.shape = tuple(shape)
            self.low, self.high = np.full(self.shape, low), np.full(self.shape, high)
        self.low = self.low.astype(self.dtype)
        self.high = self.high.astype(self.dtype)
        self.bounded_below = -np.inf < self.low
        self.bounded_above = np.inf > self.high
        super().__init__(self.shape, self.dtype)

    def is_bounded(self, manner="both"):
        below, above = np.all(self.bounded_below), np.all(self.bounded_above)
        return {"both": below and above, "below": below, "above": above}[manner]

    def sample(self):
        high = self.high if self.dtype.kind == "f" else self.high.astype("int64") + 1
        sample = np.empty(self.shape)
        unbounded = ~self.bounded_below & ~self.bounded_above
        upp_bounded = ~self.bounded_below & self.bounded_above
        low_bounded = self.bounded_below & ~self.bounded_above
        bounded = self.bounded_below & self.bounded_above
        sample[unbounded] = self.np_random.normal(size=unbounded[unbounded].shape)
        sample[low_bounded] = self.np_random.exponential(size=low_bounded[low_bounded].shape) + self.low[low_bounded]
        sample[upp_bounded] = -self.np_random.exponential(size=upp_bounded[upp_bounded].shape) + self.high[upp_bounded]
        sample[bounded] = self.np_random.uniform(low=self.low[bounded], high=high[bounded], size=bounded[bounded].shape)
        if self.dtype.kind == "i":
            sample = np.floor(sample)
        return sample.astype(self.dtype)

    def contains(self, x):
        if isinstance(x, list):
            x = np.array(x)
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    def __eq__(self, other):
        return isinstance(other, Box) and self.shape == other.shape and np.allclose(self.low, other.low) \
            and np.allclose(self.high, other.high)
