from vision4leg_b200.networks.init import *  # noqa: F401,F403
from vision4leg_b200.networks import init as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith('__')})
