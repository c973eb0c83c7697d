from .mc3 import BasicMC3, MC3Uniform, MC3Hamilton

__all__ = ['BasicMC3', 'MC3Uniform', 'MC3Hamilton']
