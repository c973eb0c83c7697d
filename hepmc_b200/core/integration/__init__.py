from .integration import PlainMC, IntegrationSample
from .importance import ImportanceMC, MultiChannelMC
from .vegas import VegasMC, VegasSample
from .stratified_volume import GridVolumes
from .stratified import StratifiedMC, StratifiedSample
from .multi_channel import MultiChannel, ChannelsSample

__all__ = ['PlainMC', 'ImportanceMC', 'MultiChannelMC', 'VegasMC', 'GridVolumes', 'MultiChannel',
           'IntegrationSample', 'VegasSample', 'ChannelsSample', 'StratifiedMC', 'StratifiedSample']
