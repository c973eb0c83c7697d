"""Gaussian density / distribution with a diagonal covariance
(hepmc/core/densities/gaussian.py).  Also the HMC momentum distribution."""
import ctypes
import math

import numpy as np
import torch

from ..density import BuiltinDensity, Distribution
from ... import _lib, rng


class Gaussian(BuiltinDensity, Distribution):
    kind = _lib.GAUSSIAN

    def __init__(self, ndim, mu=0, cov=None, scale=None):
        from ..proposals.gaussian import diagonal_covariance
        BuiltinDensity.__init__(self, ndim, False)
        self._mean = self._cov = self._cov_inv = None
        self.mean = mu
        self.cov = diagonal_covariance(ndim, cov, scale)

    def _diag(self):
        cov = np.asarray(self._cov)
        if np.count_nonzero(cov - np.diag(np.diagonal(cov))):
            raise RuntimeError("hepmc_b200.Gaussian supports diagonal covariances only")
        return np.diagonal(cov)

    def _fill_block(self, blk):
        diag = self._diag()
        for d in range(self.ndim):
            blk.a[d] = float(self._mean[d])
            blk.b[d] = math.sqrt(1.0 / float(diag[d]))
            blk.c[d] = 1.0 / float(diag[d])
        blk.s1 = self.ndim * math.log(2 * math.pi) + float(np.sum(np.log(diag)))

    def rvs(self, sample_size):
        """`sample_size` draws N(mean, cov) with the native generator -> (n, ndim) CUDA tensor.
        (The reference draws np.random.multivariate_normal, gaussian.py:49-51.)"""
        from ..markov.ensemble import gaussian_rvs
        return gaussian_rvs(self._mean, np.sqrt(self._diag()), int(sample_size))

    @property
    def mean(self):
        return self._mean

    @mean.setter
    def mean(self, value):
        if np.isscalar(value):
            self._mean = np.full(self.ndim, float(value))
        else:
            self._mean = np.array(value, dtype=np.float64, ndmin=1)

    @property
    def variance(self):
        return np.diagonal(self._cov)

    @property
    def cov(self):
        return self._cov

    @cov.setter
    def cov(self, value):
        # element-wise product with the identity (gaussian.py:73-76): the stored matrix is diagonal
        self._cov = np.eye(self.ndim, self.ndim) * np.asanyarray(value)
        self._cov_inv = np.diag(1.0 / np.diagonal(self._cov))

    def __repr__(self):
        return type(self).__name__ + "(ndim=%s, mu=%s, cov=%s)" % (self.ndim, self.mean, self.cov)
