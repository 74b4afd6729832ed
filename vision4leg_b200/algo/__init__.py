from .on_policy import PPO, A2C, OnRLAlgo   # noqa: F401
from .rl_algo import RLAlgo                 # noqa: F401
