from .replay import *
from .envs import *
