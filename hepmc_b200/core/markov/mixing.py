"""Mixing of update mechanisms (mirror of hepmc/core/markov/base.py:139-179).

`MixingMarkovUpdate(ndim, updates, weights)`: at every step one of the updates is chosen with
the given probabilities.  Device kernel (`hepmc_mc3_chains`): the (MC)^3 combination the
package is built around -- updates = [importance-sampling Metropolis whose proposal is a
`MultiChannel`, local update], local update = `DefaultMetropolis` with a Gaussian proposal or
`proposals.UniformLocal`, or `HamiltonianUpdate` (exact or ELM-surrogate gradient), all over one
built-in target density.
"""
import ctypes

import numpy as np
import torch

from .base import MarkovUpdate
from .metropolis import DefaultMetropolis
from ..density import as_builtin, BuiltinDensity
from ... import _lib, rng
from . import ensemble


def run_mc3(dens, channels, beta, local, x0, sample_size, thin, store_chain, first_chain, replay, total_chains=None):
    """Launch hepmc_mc3_chains.  `local` is None (beta == 1), a DefaultMetropolis with a Gaussian or
    UniformLocal proposal, or a HamiltonianUpdate."""
    from ...hamiltonian.hmc import HamiltonianUpdate
    from ...surrogate.extreme_learning import SurrogateGradient
    from ..proposals import Gaussian as GaussianProposal, UniformLocal
    from ..densities import Gaussian as GaussianDensity
    dev = x0.device
    n_chains, ndim = x0.shape
    chain, last, accepted, total = ensemble.alloc_outputs(n_chains, ndim, sample_size, thin, store_chain, dev)
    local_kind, scale, step_size, steps, elm_blk, keep, prec = 0, None, 0.0, 0, None, None, 0
    counted = None
    if local is None:
        if beta < 1.0:
            raise TypeError("a local update is needed when beta < 1")
    elif isinstance(local, HamiltonianUpdate):
        if local.target_density is not dens and as_builtin(local.target_density) is not dens:
            raise TypeError("all mixed updates must share one built-in target density")
        if not isinstance(local.p_dist, GaussianDensity) or np.any(local.p_dist.mean != 0):
            raise TypeError("the momentum distribution must be a zero-mean densities.Gaussian")
        local_kind, scale = 1, local.p_dist._diag()
        step_size, steps = float(local.step_size), int(local.steps)
        grad = vars(dens).get("pot_gradient", None)
        counted = grad if hasattr(grad, "count") and hasattr(grad, "fn") else None   # util.count_calls wrapper
        grad = getattr(grad, "fn", grad)
        if getattr(grad, "__self__", None) is dens and getattr(grad, "__func__", None) is type(dens).pot_gradient:
            grad = None                             # the density's own gradient behind util.count_calls
        if grad is not None:
            if not isinstance(grad, SurrogateGradient):
                raise TypeError("a replaced pot_gradient must be a surrogate.SurrogateGradient")
            elm_blk, keep = grad.block()
        prec = {"f64": 0, "f32": 1}.get(local.precision, 0) if elm_blk is not None else 0
        if elm_blk is not None and local.precision in ("tc", "tc32"):
            import warnings
            warnings.warn("the (MC)^3 mixing kernel has no tensor-core gradient: precision=%r of the local "
                          "HamiltonianUpdate runs as 'f64' here" % local.precision, RuntimeWarning, stacklevel=3)
    elif isinstance(local, DefaultMetropolis) and isinstance(local._proposal, GaussianProposal):
        cov = np.asarray(local._proposal.cov)
        if np.count_nonzero(cov - np.diag(np.diagonal(cov))):
            raise TypeError("the local random walk needs a diagonal proposal covariance")
        local_kind, scale = 0, np.diagonal(cov)
    elif isinstance(local, DefaultMetropolis) and isinstance(local._proposal, UniformLocal):
        local_kind, scale = 2, local._proposal._kernel_params()
    else:
        raise TypeError("the local update must be HamiltonianUpdate or DefaultMetropolis with proposals.Gaussian / "
                        "proposals.UniformLocal, got " + type(local).__name__)
    sel = draw = u = None
    if replay is not None:
        sel, _ = _lib.to_device(replay[0], dtype=torch.int64)
        draw, _ = _lib.to_device(replay[1])
        u, _ = _lib.to_device(replay[2])
        if tuple(draw.shape) != (n_chains, sample_size - 1, ndim):
            raise ValueError("replay tapes must be (C, T-1), (C, T-1, ndim), (C, T-1)")
    table = channels._table()
    n_local = torch.zeros(1, dtype=torch.int64, device=dev) if counted is not None else None
    coop = ensemble.coop_flag(n_chains, elm_blk.nodes, total_chains) if elm_blk is not None else 0
    _lib.call("hepmc_mc3_chains", ctypes.byref(dens._block()), _lib.ptr(table), channels.count, float(beta),
              local_kind, _lib.dbl_array(scale) if scale is not None else None, step_size, steps,
              ctypes.byref(elm_blk) if elm_blk is not None else None, prec, coop, _lib.ptr(x0),
              n_chains, int(first_chain), int(sample_size), int(thin), rng.get_seed(), rng.next_stream(),
              _lib.ptr(sel), _lib.ptr(draw), _lib.ptr(u),
              _lib.ptr(chain), _lib.ptr(last), _lib.ptr(accepted), _lib.ptr(total), _lib.ptr(n_local), _lib.stream_ptr())
    if counted is not None:
        # util.count_calls(target, 'pot_gradient') (interfaces/mc3_hmc_el.py:35): the kernel evaluates the
        # gradient steps + 1 times per local HMC update (simulation.py:34-42, one row per call)
        counted.count += int(n_local.item()) * (steps + 1)
    del keep
    return chain, last, accepted, total


class MixingMarkovUpdate(MarkovUpdate):

    def __init__(self, ndim, updates, weights=None, masks=None, target=None):
        is_adaptive = any(update.is_adaptive for update in updates)
        if target is None:
            for update in updates:
                if update.target is not None:
                    target = update.target
                    break
        super().__init__(ndim, is_adaptive=is_adaptive, target=target)
        if masks:
            raise NotImplementedError("masked updates are host-side control flow (out of scope)")
        self.updates = updates
        self.updates_count = len(updates)
        if weights is None:
            weights = np.ones(self.updates_count) / self.updates_count
        self.weights = weights

    def _run_ensemble(self, x0, sample_size, thin, store_chain, first_chain, replay):
        from ..integration.multi_channel import MultiChannel
        if self.updates_count != 2:
            raise TypeError("the device kernel mixes exactly [importance-sampling Metropolis, local update]")
        is_upd, local = self.updates
        from ..densities import Uniform
        if isinstance(is_upd, DefaultMetropolis) and isinstance(is_upd._proposal, Uniform):
            # the reference's default independence proposal (interfaces/mc3_hmc_uni.py:14) is the
            # one-channel mixture: constant q, the Hastings ratio reduces to the plain one
            channels = MultiChannel([is_upd._proposal])
        elif isinstance(is_upd, DefaultMetropolis) and isinstance(is_upd._proposal, MultiChannel):
            channels = is_upd._proposal
        else:
            raise TypeError("updates[0] must be DefaultMetropolis(ndim, target[, MultiChannel(...)])")
        dens = as_builtin(is_upd.target) or as_builtin(is_upd.pdf)
        if dens is None:
            # user-defined target: per-step launches, the user's pdf runs on the (C, ndim) CUDA tensor
            if replay is not None or not callable(is_upd.pdf):
                raise TypeError("a user-defined target needs a callable pdf; replay tapes are available for the "
                                "built-in densities only")
            return ensemble.run_mixing_generic(is_upd.pdf, channels, float(self.weights[0]), local, x0, sample_size,
                                               thin, store_chain, first_chain)
        return run_mc3(dens, channels, float(self.weights[0]), local, x0, sample_size, thin, store_chain,
                       first_chain, replay, getattr(self, "_total_chains", None))
