"""torchrl.algo: PPO / A2C come from vision4leg_b200; the other algorithm names the starters
import (VMPO, PPOAux, TRPO, Reinforce, SAC, ...) resolve from the reference checkout when one is
on the path, and raise a clear ImportError otherwise."""
from torchrl import _extend as _ext
_ext(__path__, "algo")

from vision4leg_b200.algo import PPO, A2C, OnRLAlgo, RLAlgo   # noqa: E402,F401
from . import on_policy                                        # noqa: E402,F401

__all__ = ["PPO", "A2C"]


def __getattr__(name):
  import importlib
  for mod in ("torchrl.algo.on_policy.%s" % {"VMPO": "v_mpo", "PPOAux": "ppo_aux", "TRPO": "trpo",
                                              "Reinforce": "reinforce"}.get(name, "_none_"),
              "torchrl.algo.off_policy"):
    try:
      m = importlib.import_module(mod)
    except Exception:
      continue
    if hasattr(m, name):
      return getattr(m, name)
  raise AttributeError(
    "torchrl.algo.%s is outside the accelerated hot path and needs the reference checkout on "
    "sys.path (set V4L_REFERENCE_ROOT)" % name)
