"""Gaussian random-walk proposal (mirror of hepmc/core/proposals/gaussian.py:7-39).
Inside the chain kernels the candidate is  state + sqrt(diag cov) * z."""
import numpy as np

from ..density import Proposal


class Gaussian(Proposal):

    def __init__(self, ndim=1, cov=None, scale=None):
        super().__init__(ndim, True)
        self._cov = None
        if cov is None:
            if scale is None:
                self.cov = np.ones(ndim)
            else:
                if not np.isscalar(scale):
                    scale = np.atleast_1d(scale)
                self.cov = scale ** 2
        else:
            self.cov = cov
            if scale is not None:
                raise RuntimeWarning("Specified both cov and scale (using cov)")

    def proposal(self, state):
        """One candidate for one state (host convenience; the kernels draw in-register)."""
        from ..markov.ensemble import gaussian_rvs
        state = np.atleast_1d(np.asarray(state, dtype=np.float64))
        return gaussian_rvs(state, np.sqrt(np.diagonal(self._cov)), 1)[0].cpu().numpy()

    def proposal_pdf(self, state, candidate):
        from ..densities import Gaussian as GaussianDensity
        return float(GaussianDensity(self.ndim, mu=np.asarray(state), cov=self._cov).pdf(candidate)[0])

    @property
    def cov(self):
        return self._cov

    @cov.setter
    def cov(self, value):
        self._cov = np.eye(self.ndim, self.ndim) * np.asanyarray(value)
