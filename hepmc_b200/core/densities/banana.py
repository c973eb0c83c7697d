"""Banana density (hepmc/core/densities/banana.py): a Gaussian N(0, diag(100,1,..,1)) in
the twisted coordinates phi = (x0, x1 + b x0^2 - 100 b, x2, ...)."""
import math

from ..density import BuiltinDensity
from ... import _lib


class Banana(BuiltinDensity):
    kind = _lib.BANANA

    def __init__(self, ndim, bananicity=0.1):
        super().__init__(ndim, False)
        self.b = bananicity

    def _fill_block(self, blk):
        n = self.ndim
        blk.s0 = float(self.b)
        logdet = 0.0
        for d in range(n):
            cov = 100.0 if d == 0 else 1.0
            blk.a[d] = math.sqrt(1.0 / cov)
            logdet += math.log(cov)
        blk.s1 = n * math.log(2 * math.pi) + logdet
        # constant of Banana.pot as the reference writes it (banana.py:35; note the sign, B3)
        blk.s3 = -math.log(math.sqrt((2 * math.pi) ** n * 100))

    def pot(self, xs):
        """Reference-compatible value: banana.py:35 returns +log pdf."""
        return super().pot(xs)

    def pot_correct(self, xs):
        """-log pdf (what `pot` means everywhere else)."""
        return -self.pot(xs)

    def pdf_gradient(self, xs):
        raise NotImplementedError()

    def __repr__(self):
        return type(self).__name__ + "(ndim=%s, bananaicity=%s)" % (self.ndim, self.b)
