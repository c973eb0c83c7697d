"""Metropolis updates (mirror of hepmc/core/markov/metropolis.py:20-127).

`DefaultMetropolis(ndim, target, proposal=None)`: accept with min(1, pdf(cand)/pdf(state)).
Device kernel: target = a built-in density (or its bound `pdf`), proposal =
`proposals.Gaussian` with a diagonal covariance (random walk), `proposals.UniformLocal`, or
the reference's default `densities.Uniform` independence proposal.  Any other target (a user
`Density` subclass or a callable pdf taking (C, ndim) CUDA tensors) runs the generic path of
`ensemble.run_rwm_generic`: one round of launches per step, vectorised over the ensemble.
"""
import numpy as np

from ..density import Density, as_builtin
from ..densities import Uniform
from .base import MarkovUpdate
from . import ensemble


class MetropolisUpdate(MarkovUpdate):

    def __init__(self, ndim, target, adaptive=False, hasting=False):
        super().__init__(ndim, is_adaptive=adaptive, target=target)
        self.pdf = target.pdf if isinstance(target, Density) else target
        self.is_hasting = hasting

    def proposal_pdf(self, state, candidate):
        pass


class DefaultMetropolis(MetropolisUpdate):
    """Metropolis(-Hastings) update with a `Proposal` (metropolis.py:98-127).

    Native draws of the ensemble kernel (symmetric proposals over a built-in target): the accept uniform of a
    step is taken from the 24 bits the step's proposal draw leaves unused, u = (k + 1/2) 2^-24.  The realised
    acceptance probability is therefore the Metropolis ratio rounded to a multiple of 2^-24 (absolute error
    <= 3e-8 per step), and a move with ratio < 3e-8 is never accepted.  `HamiltonianUpdate` and the (MC)^3
    mixing kernel draw a full 52-bit accept uniform; replay tapes are used as given."""

    def __init__(self, ndim, target, proposal=None, adaptive=False):
        if proposal is None:
            proposal = Uniform(ndim)
        self._proposal = proposal
        super().__init__(ndim, target, adaptive=adaptive, hasting=not proposal.is_symmetric)

    def proposal_pdf(self, state, candidate):
        return self._proposal.proposal_pdf(state, candidate)

    def _proposal_params(self):
        from ..proposals import Gaussian as GaussianProposal, UniformLocal
        prop = self._proposal
        if isinstance(prop, UniformLocal):
            return 2, prop._kernel_params()
        if isinstance(prop, GaussianProposal):
            cov = np.asarray(prop.cov)
            if np.count_nonzero(cov - np.diag(np.diagonal(cov))):
                raise TypeError("the RWM kernel supports diagonal proposal covariances")
            return 0, np.sqrt(np.diagonal(cov))
        if isinstance(prop, Uniform):
            return 1, np.array([prop.low, prop.high], dtype=np.float64)
        raise TypeError("DefaultMetropolis on the device needs proposals.Gaussian, proposals.UniformLocal or "
                        "densities.Uniform, got "
                        + type(prop).__name__)

    def _run_ensemble(self, x0, sample_size, thin, store_chain, first_chain, replay):
        dens = as_builtin(self.target) or as_builtin(self.pdf)
        if dens is None:
            # user-defined target: per-step launches, the user's pdf runs on the (C, ndim) CUDA tensor
            if self.is_adaptive or replay is not None or not callable(self.pdf):
                raise TypeError("a user-defined target needs a callable pdf; adaptive updates and replay tapes "
                                "are available for the built-in densities only")
            from ..integration.multi_channel import MultiChannel
            if isinstance(self._proposal, MultiChannel):
                raise TypeError("importance-sampling Metropolis over a MultiChannel proposal needs a built-in target")
            kind, params = self._proposal_params()
            return ensemble.run_rwm_generic(self.pdf, x0, kind, params, sample_size, thin, store_chain, first_chain)
        if dens.ndim != self.ndim:
            raise ValueError('target must have dimension ' + str(self.ndim))
        if self.is_adaptive:
            raise TypeError("adaptive Metropolis mutates shared state per step: out of scope")
        from ..integration.multi_channel import MultiChannel
        if isinstance(self._proposal, MultiChannel):
            # importance-sampling Metropolis-Hastings: the (MC)^3 kernel with beta = 1
            from .mixing import run_mc3
            return run_mc3(dens, self._proposal, 1.0, None, x0, sample_size, thin, store_chain, first_chain, replay)
        kind, params = self._proposal_params()
        return ensemble.run_rwm(dens._block(), x0, kind, params, sample_size, thin, store_chain,
                                first_chain, None, replay)
