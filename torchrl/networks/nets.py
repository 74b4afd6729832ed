from vision4leg_b200.networks.nets import *  # noqa: F401,F403
from vision4leg_b200.networks import nets as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith('__')})
