"""Uniform density on the open box (low, high)^ndim (hepmc/core/densities/uniform.py)."""
import numpy as np
import torch

from ..density import BuiltinDensity, Distribution
from ... import _lib, rng


class Uniform(BuiltinDensity, Distribution):
    kind = _lib.UNIFORM

    def __init__(self, ndim, sample_range=(0, 1)):
        BuiltinDensity.__init__(self, ndim, True)
        self.low = sample_range[0]
        self.high = sample_range[1]
        self.vol = np.prod(np.diff(sample_range, axis=0))

    def _fill_block(self, blk):
        blk.s0 = float(self.low)
        blk.s1 = float(self.high)
        blk.s2 = float(self.vol)

    def pdf_gradient(self, xs):
        return torch.zeros_like(xs) if isinstance(xs, torch.Tensor) else np.zeros(np.shape(xs))

    def pot_gradient(self, xs):
        return self.pdf_gradient(xs)

    def rvs(self, sample_size):
        """(n, ndim) uniform draws on the box with the native generator (uniform.py:33-34)."""
        from ..integration.engine import uniform_rvs
        u = uniform_rvs(int(sample_size), self.ndim)
        if self.low != 0 or self.high != 1:
            u = self.low + (self.high - self.low) * u
        return u

    def __repr__(self):
        return type(self).__name__ + "(ndim=%s, sample_range=%s)" % (self.ndim, (self.low, self.high))
