"""Sampler recipes: `interfaces.<name>.get_sampler(target, ndim, initial, ...) -> (sampler, initial)`
with the argument lists of hepmc/interfaces/*.py (plain_uniform.py:5-8, plain_hmc.py:6-14,
mc3_gauss.py:6-26, mc3_hmc.py:7-29, mc3_hmc_uni.py:7-22, mc3_hmc_el.py:8-45).

Every recipe is assembled from three steps -- channels (multi-channel integration that adapts the
importance-sampling proposal), an optional surrogate (ELM trained on -log f of the integration
sample, attached as `target.pot_gradient`), and the local update -- and returns the updates
(all of them run in the chain kernels).  The NUTS / spherical-NUTS recipes of the reference
need tree-building control flow per chain and are out of scope (NotImplementedError).
"""
import types

import numpy as np

from ..core import densities, proposals
from ..core.markov import DefaultMetropolis, MixingMarkovUpdate
from ..core.integration import MultiChannel, MultiChannelMC
from ..hamiltonian import HamiltonianUpdate
from ..surrogate import extreme_learning


def _start(ndim, initial):
    return np.full(ndim, initial) if np.isscalar(initial) else initial


def _adapted_channels(target, ndim, centers, widths, nintegral):
    """Gaussian channels at `centers` with `widths`, weights adapted by one multi-channel
    iteration of `nintegral` points (mc3_*.py: `mc_importance(target, [], [nintegral], [])`)."""
    channels = MultiChannel([densities.Gaussian(ndim, mu=c, scale=w) for c, w in zip(centers, widths)])
    sample = MultiChannelMC(channels)(target, [], [nintegral], [])
    return channels, sample


def _train_surrogate(target, ndim, sample, el_size):
    """ELM on -log f over the points with f > 0 (mc3_hmc_el.py:23-37); the surrogate's gradient
    replaces target.pot_gradient."""
    values = np.asarray(sample.function_values.cpu() if hasattr(sample.function_values, "cpu")
                        else sample.function_values)
    data = np.asarray(sample.data.cpu() if hasattr(sample.data, "cpu") else sample.data)
    keep = values > 0
    basis = extreme_learning.GaussianBasis(ndim)
    params, bias, weights = basis.extreme_learning_train(data[keep], -np.log(values[keep]), el_size)
    return extreme_learning.attach_surrogate_gradient(target, basis, params, bias, weights)


def _hmc(target, ndim, mass, steps, step_size):
    return HamiltonianUpdate(target, densities.Gaussian(ndim, cov=mass), steps, step_size)


def _mixed(ndim, is_sampler, local, beta):
    return MixingMarkovUpdate(ndim, [is_sampler, local], [beta, 1 - beta])


def _plain_uniform(target, ndim, initial):
    return DefaultMetropolis(ndim, target), _start(ndim, initial)


def _plain_hmc(target, ndim, initial, mass=10, steps=10, step_size=.1):
    return _hmc(target, ndim, mass, steps, step_size), _start(ndim, initial)


def _mc3_gauss(target, ndim, initial, centers, widths, beta, nintegral=1000, local_width=.1):
    channels, _ = _adapted_channels(target, ndim, centers, widths, nintegral)
    local = DefaultMetropolis(ndim, target, proposals.Gaussian(ndim, scale=local_width))
    return _mixed(ndim, DefaultMetropolis(ndim, target, channels), local, beta), _start(ndim, initial)


def _mc3_hmc(target, ndim, initial, centers, widths, beta, nintegral=1000, mass=10, steps=10, step_size=.1):
    channels, _ = _adapted_channels(target, ndim, centers, widths, nintegral)
    return (_mixed(ndim, DefaultMetropolis(ndim, target, channels), _hmc(target, ndim, mass, steps, step_size), beta),
            _start(ndim, initial))


def _mc3_hmc_uni(target, ndim, initial, beta, mass=10, steps=10, step_size=.1):
    return (_mixed(ndim, DefaultMetropolis(ndim, target), _hmc(target, ndim, mass, steps, step_size), beta),
            _start(ndim, initial))


def _mc3_hmc_el(target, ndim, initial, centers, widths, beta, nintegral=1000, mass=10, steps=10, step_size=.1,
                el_size=100):
    channels, sample = _adapted_channels(target, ndim, centers, widths, nintegral)
    _train_surrogate(target, ndim, sample, el_size)
    return (_mixed(ndim, DefaultMetropolis(ndim, target, channels), _hmc(target, ndim, mass, steps, step_size), beta),
            _start(ndim, initial))


def _out_of_scope(name):
    def get_sampler(*args, **kwargs):
        raise NotImplementedError("interfaces.%s: NUTS-type updates are out of scope of hepmc_b200" % name)
    return get_sampler


_RECIPES = dict(plain_uniform=_plain_uniform, plain_hmc=_plain_hmc, mc3_gauss=_mc3_gauss, mc3_hmc=_mc3_hmc,
                mc3_hmc_uni=_mc3_hmc_uni, mc3_hmc_el=_mc3_hmc_el)
for _name in ("plain_nuts", "plain_spherical_nuts", "mc3_nuts", "mc3_spherical_nuts", "mc3_nuts_el",
              "mc3_spherical_nuts_el"):
    _RECIPES[_name] = _out_of_scope(_name)
for _name, _fn in _RECIPES.items():
    _fn.__name__ = "get_sampler"
    globals()[_name] = types.SimpleNamespace(get_sampler=_fn, __name__=__name__ + "." + _name)
__all__ = sorted(_RECIPES)
del _name, _fn
