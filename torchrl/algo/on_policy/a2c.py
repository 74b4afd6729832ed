from vision4leg_b200.algo.on_policy.a2c import *  # noqa: F401,F403
from vision4leg_b200.algo.on_policy import a2c as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith('__')})
