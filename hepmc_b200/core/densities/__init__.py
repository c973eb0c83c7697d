from .gaussian import Gaussian
from .camel import Camel, UnconstrainedCamel
from .uniform import Uniform
from .banana import Banana
from .rambo import Rambo, RamboOnDiet

__all__ = ['Gaussian', 'Camel', 'UnconstrainedCamel', 'Uniform', 'Banana', 'Rambo', 'RamboOnDiet']
