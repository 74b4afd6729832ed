from .continuous_policy import *      # noqa: F401,F403
from .continuous_policy import (GaussianContPolicyBase, GaussianContPolicyBasicBias,  # noqa: F401
                                GaussianContPolicyImpalaEncoderProj, GaussianContPolicyNatureEncoderProj,
                                GaussianContPolicyTransformer,
                                GaussianContPolicyLocoTransformer, LOG_SIG_MAX, LOG_SIG_MIN)
from .distribution import TanhNormal  # noqa: F401
