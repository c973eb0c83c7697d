from .gaussian import Gaussian

__all__ = ['Gaussian']
