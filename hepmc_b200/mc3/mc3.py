"""(MC)^3: multi-channel integration followed by mixed importance / local sampling
(mirror of hepmc/mc3/mc3.py:17-149).  Integration = MultiChannelMC (adapts the channel
weights); sampling = MixingMarkovUpdate([IS-Metropolis over the adapted channels, local
update]) -- both phases run in the CUDA kernels.
"""
import numpy as np

from ..core.densities import Gaussian
from ..core.markov import MixingMarkovUpdate, DefaultMetropolis
from ..core.proposals import UniformLocal
from ..core.integration import MultiChannelMC
from ..hamiltonian import HamiltonianUpdate


class BasicMC3(object):
    def __init__(self, target, channels, sample_local, beta=.5):
        self.ndim = channels.ndim
        self.target = target
        self.channels = channels
        self.mc_importance = MultiChannelMC(channels)
        self.sample_is = DefaultMetropolis(self.ndim, target, channels)
        self.sample_local = sample_local
        self.mixing_sampler = MixingMarkovUpdate(self.ndim, [self.sample_is, self.sample_local])
        self.beta = beta
        self.integration_sample = None

    def integrate(self, *eval_sizes, **kwargs):
        """Multi-channel integration / channel-weight optimisation (mc3.py:40-50)."""
        self.integration_sample = self.mc_importance(self.target, *eval_sizes, **kwargs)
        return self.integration_sample

    def sample(self, sample_size, initial=None, **kwargs):
        """MC3 sample (mc3.py:52-74).  Without `initial` a start with non-zero density is drawn
        from the channels."""
        if initial is None:
            for _ in range(1000):
                cand = np.asarray(self.channels.rvs(64))
                pdf = np.asarray(self.target.pdf(cand))
                ok = np.nonzero(pdf != 0)[0]
                if ok.size:
                    initial = cand[ok[0]]
                    break
            if initial is None:
                raise RuntimeError("Could not find a suitable initial value using the multi channel distribution.")
        return self.mixing_sampler.sample(sample_size, initial, **kwargs)

    @property
    def beta(self):
        return self.mixing_sampler.weights[0]

    @beta.setter
    def beta(self, beta):
        self.mixing_sampler.weights = np.array([beta, 1 - beta])

    def __call__(self, eval_sizes, sample_size, initial=None, **kwargs):
        self.integrate(*eval_sizes)
        return self.sample(sample_size, initial, **kwargs)


def _forward(owner, name):
    """Property forwarding attribute `name` to the sub-object `owner` (e.g. the local sampler)."""
    return property(lambda self: getattr(getattr(self, owner), name),
                    lambda self, value: setattr(getattr(self, owner), name, value))


class MC3Uniform(BasicMC3):
    """(MC)^3 with a local uniform-box Metropolis update of edge `delta` (mc3.py:100-114)."""

    delta = _forward("_local_proposal", "delta")

    def __init__(self, fn, channels, delta, beta=.5):
        self._local_proposal = UniformLocal(channels.ndim, delta)
        BasicMC3.__init__(self, fn, channels,
                          DefaultMetropolis(channels.ndim, fn, self._local_proposal), beta)


class MC3Hamilton(BasicMC3):
    """(MC)^3 with an HMC local update; `m` = momentum standard deviations (mc3.py:117-149)."""

    step_size = _forward("sample_local", "step_size")
    steps = _forward("sample_local", "steps")

    def __init__(self, target_density, channels, m, steps, step_size, beta=.5):
        self.p_dist = Gaussian(channels.ndim, scale=m)
        BasicMC3.__init__(self, target_density, channels,
                          HamiltonianUpdate(target_density, self.p_dist, steps, step_size), beta)

    @property
    def m(self):
        return np.sqrt(self.p_dist.cov)

    @m.setter
    def m(self, value):
        self.p_dist.cov = np.square(value)
