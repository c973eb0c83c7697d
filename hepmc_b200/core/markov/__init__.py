from .base import MarkovUpdate, MarkovSample
from .metropolis import MetropolisUpdate, DefaultMetropolis
from .mixing import MixingMarkovUpdate

__all__ = ['MarkovUpdate', 'MarkovSample', 'MetropolisUpdate', 'DefaultMetropolis', 'MixingMarkovUpdate']
