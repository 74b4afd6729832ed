from vision4leg_b200.networks import *          # noqa: F401,F403
from vision4leg_b200.networks import base, nets, init   # noqa: F401
