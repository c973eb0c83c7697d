"""Proposal generators understood by the chain kernels."""
from .gaussian import Gaussian

__all__ = ["Gaussian"]
