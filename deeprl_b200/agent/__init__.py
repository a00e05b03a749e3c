from .BaseAgent import *
from .DQN_agent import *
from .CategoricalDQN_agent import *
from .QuantileRegressionDQN_agent import *
from .A2C_agent import *
from .PPO_agent import *
from .NStepDQN_agent import *
