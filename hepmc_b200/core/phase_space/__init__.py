"""Phase-space mappings of the unit hypercube (RAMBO family) and densities seen through them."""
from .mapping import MappedDensity
from .rambo import PhaseSpaceMapping, Rambo, RamboOnDiet

__all__ = [name for name in dir() if name[0].isupper()]
