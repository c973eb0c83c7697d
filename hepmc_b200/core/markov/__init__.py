from .base import MarkovUpdate, MarkovSample
from .metropolis import MetropolisUpdate, DefaultMetropolis

__all__ = ['MarkovUpdate', 'MarkovSample', 'MetropolisUpdate', 'DefaultMetropolis']
