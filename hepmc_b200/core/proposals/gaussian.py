"""Gaussian random-walk proposal, `proposals.Gaussian(ndim, cov=None, scale=None)`
(reference: hepmc/core/proposals/gaussian.py:7-39).

Inside the chain kernels the candidate is `state + sqrt(diag cov) * z` with z from
Box-Muller on Philox draws; this object only carries the (diagonal) covariance.
"""
import numpy as np

from ..density import Proposal


def diagonal_covariance(ndim, cov, scale):
    """(ndim, ndim) covariance from the reference's two spellings: `cov` (scalar / per-axis /
    matrix, multiplied element-wise into the identity) or `scale` (standard deviations)."""
    if cov is not None and scale is not None:
        raise RuntimeWarning("Specified both cov and scale (using cov)")
    if cov is None:
        cov = np.ones(ndim) if scale is None else np.square(scale if np.isscalar(scale) else np.atleast_1d(scale))
    return np.eye(ndim, ndim) * np.asanyarray(cov)


class Gaussian(Proposal):

    def __init__(self, ndim=1, cov=None, scale=None):
        Proposal.__init__(self, ndim, True)      # symmetric: plain Metropolis ratio
        self._cov = diagonal_covariance(ndim, cov, scale)

    cov = property(lambda self: self._cov,
                   lambda self, value: setattr(self, "_cov", diagonal_covariance(self.ndim, value, None)))

    def proposal(self, state):
        """One candidate for one state (host convenience; the kernels draw in registers)."""
        from ..markov.ensemble import gaussian_rvs
        centre = np.atleast_1d(np.asarray(state, dtype=np.float64))
        return gaussian_rvs(centre, np.sqrt(np.diagonal(self._cov)), 1)[0].cpu().numpy()

    def proposal_pdf(self, state, candidate):
        from ..densities import Gaussian as GaussianDensity
        return float(GaussianDensity(self.ndim, mu=np.asarray(state), cov=self._cov).pdf(candidate)[0])
