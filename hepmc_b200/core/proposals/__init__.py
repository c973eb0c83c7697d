"""Proposal generators understood by the chain kernels.

kernel proposal kinds (hepmc_rwm_chains / hepmc_mc3_chains):
    0  Gaussian random walk, diagonal covariance   -> gaussian.Gaussian
    1  uniform independence proposal                -> densities.Uniform (DefaultMetropolis' default)
    2  uniform box around the state                 -> uniform_local.UniformLocal
"""
import importlib

_KINDS = {"Gaussian": ("gaussian", 0), "UniformLocal": ("uniform_local", 2)}
__all__ = sorted(_KINDS)

for _name, (_module, _kind) in _KINDS.items():
    globals()[_name] = getattr(importlib.import_module("." + _module, __name__), _name)
del _name, _module, _kind
