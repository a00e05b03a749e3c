from .network_utils import *
from .network_bodies import *
from .network_heads import *
from .fused import frame_scale
