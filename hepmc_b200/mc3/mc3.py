"""(MC)^3: multi-channel integration followed by mixed importance / local sampling
(mirror of hepmc/mc3/mc3.py:17-149).  Integration = MultiChannelMC (adapts the channel
weights); sampling = MixingMarkovUpdate([IS-Metropolis over the adapted channels, local
update]) -- both phases run in the CUDA kernels.
"""
import numpy as np

from ..core.densities import Gaussian
from ..core.markov import MixingMarkovUpdate, DefaultMetropolis
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


class MC3Hamilton(BasicMC3):

    def __init__(self, target_density, channels, m, steps, step_size, beta=.5):
        self.p_dist = Gaussian(channels.ndim, scale=m)
        sample_local = HamiltonianUpdate(target_density, self.p_dist, steps, step_size)
        super().__init__(target_density, channels, sample_local, beta)

    @property
    def step_size(self):
        return self.sample_local.step_size

    @step_size.setter
    def step_size(self, step_size):
        self.sample_local.step_size = step_size

    @property
    def steps(self):
        return self.sample_local.steps

    @steps.setter
    def steps(self, steps):
        self.sample_local.steps = steps

    @property
    def m(self):
        return np.sqrt(self.p_dist.cov)

    @m.setter
    def m(self, value):
        self.p_dist.cov = value ** 2
