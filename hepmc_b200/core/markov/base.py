"""Markov chain driver (mirror of hepmc/core/markov/base.py:5-97).

`MarkovUpdate.sample(sample_size, initial, out_mask=None, log_every=5000)` keeps its
signature.  Updates whose transition is built from library pieces (DefaultMetropolis,
HamiltonianUpdate) run whole trajectories of whole ensembles inside one kernel; there is
no per-step Python loop and no CPU fallback.
"""
import numpy as np
import torch

from ..sampling import Sample
from ... import config


class MarkovSample(Sample):

    def __init__(self, **kwargs):
        self.accepted = 0
        self.chain_length = None
        super().__init__(**kwargs)
        self._sample_info.append(('accept_ratio', 'acceptance rate', '%s'))

    @property
    def accept_ratio(self):
        """accepted / sample_size (base.py:14-16); per chain for an ensemble."""
        length = self.chain_length if self.chain_length is not None else self.size
        return self.accepted / length


class MarkovUpdate(object):
    """Base of the update mechanisms."""

    def __init__(self, ndim, is_adaptive=False, target=None):
        self.ndim = ndim
        self.target = target
        self.is_adaptive = is_adaptive
        self.sample_info = None

    def init_adapt(self, initial_state):
        pass

    def init_state(self, state):
        return state

    def next_state(self, state, iteration):
        raise NotImplementedError("AbstractMarkovUpdate is abstract.")

    def _run_ensemble(self, x0, sample_size, thin, store_chain, first_chain, replay):
        """(chain | None, last, accepted, total) for the (C, ndim) starts x0."""
        raise TypeError(
            type(self).__name__ + " has no device kernel: hepmc_b200 runs DefaultMetropolis and "
            "HamiltonianUpdate over built-in densities; it has no host-side step loop (no CPU fallback).")

    def sample(self, sample_size, initial, out_mask=None, log_every=5000,
               thin=1, store_chain=True, first_chain=0, replay=None, total_chains=None):
        """Generate `sample_size` states per chain (the first is `initial`).

        :param initial: (ndim,) for one chain -> data (sample_size, ndim) like the reference;
            (C, ndim) for C independent chains -> data (C, sample_size, ndim).
        :param out_mask: slice/indices of the dimensions to return (base.py:89-90).
        :param log_every: accepted for compatibility; the kernel does not log.
        :param thin: keep every thin-th state (t = 0, thin, 2 thin, ...).
        :param store_chain: False -> keep only the final states (data has one state per chain).
        :param first_chain: global index of chain 0 (Philox counter; sharding over GPUs).
        :param replay: (z, u) tapes consumed instead of the native generator.
        :param total_chains: size of the WHOLE ensemble when `initial` is one shard of it (several
            calls or GPUs); selects the kernel layout of surrogate-gradient updates so that a chain
            does not depend on the sharding.  Default: the sum over the torch.distributed group,
            else the size of `initial`."""
        from .ensemble import prepare_initial
        x0, is_ensemble = prepare_initial(initial, self.ndim)
        self._total_chains = total_chains
        chain, last, accepted, total = self._run_ensemble(
            x0, int(sample_size), int(thin), bool(store_chain), int(first_chain), replay)
        data = chain if store_chain else last[:, None, :]
        if out_mask is not None:
            data = data[:, :, out_mask]
        sample = MarkovSample()
        sample.chain_length = int(sample_size)
        if is_ensemble:
            sample.data = config.deliver(data)
            sample.accepted = config.deliver(accepted)
        else:
            sample.data = config.deliver(data[0])
            sample.accepted = int(accepted[0].item())
        sample.last_state = config.deliver(last if is_ensemble else last[0])
        sample.total_accepted = total
        sample.target = self.target
        return sample
