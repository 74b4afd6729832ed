from vision4leg_b200.algo.on_policy.on_rl_algo import *  # noqa: F401,F403
from vision4leg_b200.algo.on_policy import on_rl_algo as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith('__')})
