"""Generate tests/golden/obs_normalizer.npz from the LIVE reference Normalizer (build container only).

  PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_obs

torchrl/env/base_wrapper.py is loaded on its own (importlib, with a stub `gym` that provides the three wrapper base
classes it subclasses); the file is not modified.  Inputs are regenerated from the seed by the tests; the fixture
holds the reference's outputs only, plus `obs_normalizer_ref.pkl`: the reference object pickled the way
RLAlgo.snapshot does (the `_obs_normalizer_{epoch}.pkl` wire format).  TEST INFRASTRUCTURE — not imported by the product.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = os.environ.get("V4L_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SEED, STEPS, N, S = 77, 6, 5, 37


def inputs():
  rng = np.random.RandomState(SEED)
  scale = rng.uniform(0.1, 30.0, size=S)
  shift = rng.uniform(-5.0, 5.0, size=S)
  # float64 arrays, as the reference's environments produce them (env_utils.flatten_observations concatenates
  # float64 sensor readings with the float32 image), holding float32-representable values so that the device
  # path's float32 rows carry the same numbers
  return [(rng.randn(N, S) * scale + shift).astype(np.float32).astype(np.float64) for _ in range(STEPS)]


def main():
  sys.dont_write_bytecode = True
  gym = types.ModuleType("gym")
  for name in ("Wrapper", "RewardWrapper", "ObservationWrapper"):
    setattr(gym, name, type(name, (), {}))
  sys.modules["gym"] = gym
  # under its real module path, so that the pickle written below names the class the way the reference's own
  # snapshot does (rl_algo.py:84-90)
  spec = importlib.util.spec_from_file_location("torchrl.env.base_wrapper", os.path.join(REF, "torchrl/env/base_wrapper.py"))
  mod = importlib.util.module_from_spec(spec)
  sys.modules["torchrl.env.base_wrapper"] = mod
  spec.loader.exec_module(mod)
  nz = mod.Normalizer((S,))
  outs = []
  for i, x in enumerate(inputs()):
    if i == STEPS - 1:
      nz.stop_update_estimate()            # the evaluation path: filter only
    nz.update_estimate(x)
    outs.append(nz.filt(x))
  np.savez_compressed(os.path.join(OUT, "obs_normalizer.npz"), filt=np.stack(outs), mean=nz._mean, var=nz._var,
                      count=np.float64(nz._count))
  import pickle
  with open(os.path.join(OUT, "obs_normalizer_ref.pkl"), "wb") as f:      # the checkpoint wire format
    pickle.dump(nz, f)
  print("wrote obs_normalizer.npz", np.stack(outs).shape, nz._count)


if __name__ == "__main__":
  main()
