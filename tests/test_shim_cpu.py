"""The `torchrl/` drop-in shim (the boundary is "whatever starter/ppo_*.py import"): with this
repo BEFORE a reference checkout on the path, the hot-path names resolve to vision4leg_b200 and everything
else falls through to the reference.  Needs the reference checkout (skipped without it: the GPU box has none);
`gym` is absent from this image, so the ten-line stub the oracle tooling uses stands in for it."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("V4L_REFERENCE_ROOT", "/root/reference")

SCRIPT = r'''
import sys, types
gym = types.ModuleType("gym"); spaces = types.ModuleType("gym.spaces")
class Box:                      # only isinstance(..., gym.spaces.Box) is used (reference rl_algo.py:36)
  def __init__(self, *a, **k): self.shape = k.get("shape", ())
spaces.Box = Box; gym.spaces = spaces
sys.modules["gym"] = gym; sys.modules["gym.spaces"] = spaces
import torchrl, torchrl.algo, torchrl.networks, torchrl.policies
from torchrl.replay_buffers.on_policy import OnPolicyReplayBuffer
ours = lambda o: o.__module__.startswith("vision4leg_b200")
assert ours(torchrl.algo.PPO) and ours(torchrl.algo.A2C), torchrl.algo.PPO.__module__
for n in ("LocoTransformer", "LocoTransformerEncoder", "NatureFuseEncoder", "ImpalaEncoderProjNet", "Net", "MLPBase",
          "NatureEncoder", "TransformerEncoder", "Transformer", "NatureEncoderProjNet"):
  assert ours(getattr(torchrl.networks, n)), n
for n in ("GaussianContPolicyLocoTransformer", "GaussianContPolicyImpalaEncoderProj", "GaussianContPolicyBasicBias"):
  assert ours(getattr(torchrl.policies, n)), n
assert ours(OnPolicyReplayBuffer)
# outside the hot path: served by the reference checkout
VMPO = torchrl.algo.VMPO
assert not ours(VMPO) and REF in sys.modules[VMPO.__module__].__file__, VMPO.__module__
import torchrl.algo.off_policy as off
assert REF in off.__file__
import torchrl.algo.utils as atu
assert hasattr(atu, "update_linear_schedule")
print("shim ok")
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "torchrl")), reason="no reference checkout")
def test_torchrl_shim_resolution():
  env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + REF, V4L_REFERENCE_ROOT=REF, PYTHONDONTWRITEBYTECODE="1")
  out = subprocess.run([sys.executable, "-c", SCRIPT.replace("REF", repr(REF))], env=env, capture_output=True, text=True,
                       timeout=300)
  assert out.returncode == 0 and "shim ok" in out.stdout, out.stdout + out.stderr
