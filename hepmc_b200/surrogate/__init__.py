"""Extreme-learning-machine surrogates (function bases, training, gradient wiring for HMC)."""
from . import extreme_learning
from .extreme_learning import (FunctionBasis, AdditiveBasis, RadialBasis, TrigBasis, GaussianBasis,
                               SurrogateGradient, attach_surrogate_gradient)
