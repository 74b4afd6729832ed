from vision4leg_b200.policies import *          # noqa: F401,F403
