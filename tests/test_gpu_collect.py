"""Collector-side inference on the device (reference torchrl/collector/on_policy.py:90-118):
`PPOUpdateEngine.act()` = pf.explore + vf for the E observations of one env step with ONE shared-encoder pass on
the tensor-core tier, writing the converted observation into the device-resident rollout planes so that the next
`update_per_epoch()` copies no observation at all."""
import numpy as np
import pytest
import torch

from benchutil import synth
from benchutil.harness import build_nets, load_np_sd, fill_buffer, make_ppo
from tests import _golden as g

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _agent(family, buf, B, frames):
  S, A = g.FAMILIES[family]
  pf, vf = build_nets(family, S, A)
  pf_np, vf_np = g.family_weights(family)
  load_np_sd(pf, pf_np); load_np_sd(vf, vf_np)
  pf, vf = pf.to(DEV), vf.to(DEV)
  agent, logger = make_ppo(pf, vf, buf, A, B, frames, 2, device=DEV)
  agent.precision = "f16"
  return agent, pf, vf


@pytest.mark.parametrize("family", ["loco", "nature"])
def test_act_matches_modules_and_feeds_the_update(family):
  S, A = g.FAMILIES[family]
  T, E, B = 8, 4, 16
  roll = synth.make_rollout(77, T, E, S, A, p_term=0.1)
  runs = {}
  for mode in ("host", "device"):
    buf = fill_buffer(roll, T, E)
    agent, pf, vf = _agent(family, buf, B, T * E)
    eng = agent.engine
    agent.current_epoch = 1
    np.random.seed(5)
    agent.update_per_epoch()                       # epoch 1: allocates the device planes (observations come from the host)
    h2d_host = eng.h2d_bytes
    if mode == "device":
      rng = np.random.default_rng(3)
      for t in range(T):
        eps = rng.standard_normal((E, A)).astype(np.float32)
        out = eng.act(roll["obs"][t], noise=eps, row=t)
        mean, value = eng.infer(roll["obs"][t])
        assert np.array_equal(out["mean"], mean.cpu().numpy()) and np.array_equal(out["value"], value.cpu().numpy())
        np.testing.assert_allclose(out["action"], out["mean"] + out["std"][None] * eps, rtol=1e-6, atol=1e-7)
        # against the exact-tier modules (what the reference collector evaluates): reduced-precision bound
        with torch.no_grad():
          x = torch.tensor(roll["obs"][t], device=DEV)
          ref_mean = pf.update(x, torch.zeros(E, A, device=DEV))["mean"].cpu().numpy()
          ref_val = vf(x).cpu().numpy()
        assert g.rel_err(out["mean"], ref_mean) < 1e-2 and g.rel_err(out["value"], ref_val) < 1e-2
    agent.current_epoch = 2
    np.random.seed(6)
    agent.update_per_epoch()
    torch.cuda.synchronize()
    runs[mode] = (eng.bucket.flat.clone(), eng.h2d_bytes, h2d_host)
  # the device-resident epoch copied no observation row ...
  assert runs["device"][1] < 0.05 * runs["device"][2], runs["device"][1:]
  assert runs["host"][1] == runs["host"][2]
  # ... and trained on exactly the same data: bit-identical parameters
  assert torch.equal(runs["host"][0], runs["device"][0])


def test_single_env_bootstrap_value_and_act():
  """E = 1 (one env column per rank in the 8-GPU strong-scaling sweep): a one-row slice x[:, S:] is already
  "contiguous" and keeps its 4*S-byte offset — the image must still reach the vectorised ingest kernel aligned."""
  S, A = g.FAMILIES["loco"]
  T, E, B = 8, 1, 4
  roll = synth.make_rollout(78, T, E, S, A, p_term=0.1)
  buf = fill_buffer(roll, T, E)
  agent, pf, vf = _agent("loco", buf, B, T * E)
  agent.current_epoch = 1
  np.random.seed(5)
  agent.update_per_epoch()
  out = agent.engine.act(roll["obs"][0], noise=np.zeros((E, A), np.float32))
  torch.cuda.synchronize()
  with torch.no_grad():
    ref = vf(torch.tensor(roll["obs"][0], device=DEV)).cpu().numpy()
  assert g.rel_err(out["value"], ref) < 1e-2
