"""Proposal generators understood by the chain kernels."""
from .gaussian import Gaussian
from .uniform_local import UniformLocal

__all__ = ["Gaussian", "UniformLocal"]
