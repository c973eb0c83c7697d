from .extreme_learning import *
from . import extreme_learning
