"""Local uniform proposal, `proposals.UniformLocal(ndim, delta, sample_range=(0, 1))`
(reference: hepmc/core/proposals/uniform_local.py:5-50).

A box of edge `delta` (scalar or per axis) around the state; near the border of
`sample_range` the box is pushed back inside, keeping its size.  As in the reference ONE
uniform number per step moves every coordinate (`np.random.rand()` there is a scalar) and the
proposal is treated as symmetric.  The chain kernels (`proposal_kind = 2` / `local_kind = 2`)
do the arithmetic; this object carries the parameters.
"""
import numpy as np

from ..density import Proposal


class UniformLocal(Proposal):

    def __init__(self, ndim, delta, sample_range=(0, 1)):
        Proposal.__init__(self, ndim, True)
        self.low, self.high = sample_range[0], sample_range[1]
        self.vol = np.prod(np.diff(sample_range, axis=0))
        self._delta = None
        self.delta = delta

    @property
    def delta(self):
        return self._delta

    @delta.setter
    def delta(self, value):
        if np.ndim(value) != 0 and len(value) != self.ndim:
            raise ValueError('delta must be a float or an array of length ndim.')
        if np.any((value - self.vol) > 0):
            raise ValueError('delta must be smaller than sample area.')
        self._delta = np.ones(self.ndim) * value

    def _kernel_params(self):
        """[delta_0 .. delta_{n-1}, low, high] as the C ABI expects them."""
        return np.concatenate((self._delta, [float(self.low), float(self.high)]))

    def corner(self, state):
        """Lower corner of the box the candidate is drawn from."""
        state = np.asarray(state, dtype=np.float64)
        return np.minimum(np.maximum(self.low, state - self._delta / 2), self.high - self._delta)

    def proposal(self, state):
        """One candidate for one state (host convenience; the kernels draw in registers)."""
        from ..integration import engine
        u = float(engine.uniform_rvs(1, 1)[0, 0])
        return self.corner(state) + u * self._delta

    def proposal_pdf(self, state, candidate):
        return 1 / np.prod(self._delta)
