from .network_utils import *
from .network_bodies import *
from .network_heads import *
