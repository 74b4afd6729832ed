from vision4leg_b200.policies.continuous_policy import *   # noqa: F401,F403
