"""Analytic target densities and sampling distributions with CUDA kernels behind them."""
from .banana import Banana
from .camel import Camel, UnconstrainedCamel
from .gaussian import Gaussian
from .rambo import Rambo, RamboOnDiet
from .uniform import Uniform

__all__ = [name for name in dir() if name[0].isupper()]
