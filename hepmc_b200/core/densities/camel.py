"""Camel densities: mixture of two isotropic Gaussians (hepmc/core/densities/camel.py).

`UnconstrainedCamel.pdf = (N(x; mu_a, cov I) + N(x; mu_b, cov I)) / 2`, `cov = a^2/2`;
`Camel` restricts it to the open unit hypercube (0 outside, so pot_gradient is inf there).
"""
import math

import numpy as np

from ..density import BuiltinDensity
from ... import _lib


def _vec(value, ndim):
    if np.isscalar(value):
        return np.full(ndim, float(value))
    return np.array(value, dtype=np.float64, ndmin=1)


class UnconstrainedCamel(BuiltinDensity):
    kind = _lib.UCAMEL

    def __init__(self, ndim, mu_a=1/3, mu_b=2/3, a=0.1):
        super().__init__(ndim, False)
        self._mu_a = self._mu_b = None
        self.mu_a = mu_a
        self.mu_b = mu_b
        self.cov = a**2/2

    def _fill_block(self, blk):
        n = self.ndim
        for d in range(n):
            blk.a[d] = float(self._mu_a[d])
            blk.b[d] = float(self._mu_b[d])
        cov = float(self.cov)
        # SciPy prepares: U = sqrt(1/eigval), log_pdet = sum(log(eigval)) over n equal eigenvalues
        blk.s0 = math.sqrt(1.0 / cov)
        blk.s1 = n * math.log(2 * math.pi) + float(np.sum(np.log(np.full(n, cov))))
        blk.s2 = cov

    @property
    def mean(self):
        return (self.mu_a + self.mu_b) / 2

    @property
    def variance(self):
        # reference value (camel.py:52-54), kept for API compatibility
        return self.cov + (self.mu_a - self.mu_b) ** 2

    @property
    def mu_a(self):
        return self._mu_a

    @mu_a.setter
    def mu_a(self, value):
        self._mu_a = _vec(value, self.ndim)

    @property
    def mu_b(self):
        return self._mu_b

    @mu_b.setter
    def mu_b(self, value):
        self._mu_b = _vec(value, self.ndim)

    def __repr__(self):
        return type(self).__name__ + "(ndim=%s, mu_a=%s, mu_b=%s, a=%s)" % (
            self.ndim, self.mu_a, self.mu_b, np.sqrt(2*self.cov))


class Camel(UnconstrainedCamel):
    kind = _lib.CAMEL
