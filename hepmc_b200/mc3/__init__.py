from .mc3 import BasicMC3, MC3Uniform, MC3Hamilton

import numpy as np   # the reference's `from .mc3 import *` re-exports its module-level names, `np` included (tests/test_mc3.py uses it)

__all__ = ['BasicMC3', 'MC3Uniform', 'MC3Hamilton', 'np']
