"""hepmc_b200 -- the data-parallel hot path of hepmc (mathisgerdes/hep-monte-carlo) as
hand-written sm_100a CUDA behind the reference's own Python API.

    from hepmc_b200 import *          # same names as `from hepmc import *`

Densities, integrators, the RAMBO generator, Metropolis / HMC ensembles and the ELM
surrogate keep the reference signatures; their bodies call libhepmc_b200.so (C ABI in
include/hepmc_b200.h) on the current CUDA stream.  There is no CPU fallback.
"""
from . import config, rng, parallel
from .config import set_outputs
from .rng import seed

from .core import proposals
from .core import densities
from .core import util
from .core import phase_space

from .core.integration import *
from .core.markov import *

from .core.sampling import Sample, AcceptRejectSampler
from .core.density import Proposal, Density, Distribution
from . import hamiltonian
from . import surrogate
from . import mc3
from . import interfaces

__version__ = "0.1.0"
