from .rambo import Rambo, PhaseSpaceMapping

__all__ = ['Rambo', 'PhaseSpaceMapping']
