from vision4leg_b200.algo.utils import *   # noqa: F401,F403
from vision4leg_b200.algo.utils import (soft_update_from_to, copy_model_params_from_to,  # noqa: F401
                                        update_linear_schedule, linear_lr)
