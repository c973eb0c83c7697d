from .hmc import HamiltonianUpdate
from .simulation import HamiltonLeapfrog

__all__ = ['HamiltonianUpdate', 'HamiltonLeapfrog']
