from .rambo import Rambo, RamboOnDiet, PhaseSpaceMapping
from .mapping import MappedDensity

__all__ = ['Rambo', 'RamboOnDiet', 'MappedDensity', 'PhaseSpaceMapping']
