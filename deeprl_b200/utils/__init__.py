from .config import *
from .normalizer import *
from .misc import *
from .logger import *
from .schedule import *
from .torch_utils import *
from .sum_tree import *
import numpy as np
import torch
import pickle
