from . import stat_tools
from .helper import (assure_2d, interpret_array, damped_update, hypercube_bounded,
                     Counted, count_calls, host_function)

__all__ = ['assure_2d', 'interpret_array', 'damped_update', 'hypercube_bounded',
           'Counted', 'count_calls', 'host_function', 'stat_tools']
