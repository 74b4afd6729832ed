from torchrl import _extend as _ext
_ext(__path__, "algo", "on_policy")
from vision4leg_b200.algo.on_policy import PPO, A2C, OnRLAlgo   # noqa: E402,F401
