from . import proposals
from . import densities
from . import util
from . import phase_space

from .integration import *
from .markov import *

from .sampling import Sample, AcceptRejectSampler
from .density import Proposal, Density, Distribution
