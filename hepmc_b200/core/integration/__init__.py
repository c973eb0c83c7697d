from .integration import PlainMC, IntegrationSample
from .importance import ImportanceMC
from .vegas import VegasMC, VegasSample
from .stratified_volume import GridVolumes

__all__ = ['PlainMC', 'ImportanceMC', 'VegasMC', 'GridVolumes', 'IntegrationSample', 'VegasSample']
