"""Pin the CPU oracle (oracle/ppo_oracle.py) against outputs of the live reference modules
(tests/golden/*.npz, produced by oracle/make_golden.py). CPU only.

Tolerances: the oracle restates the same fp32 math with a different op order, so forward
outputs agree to ~1e-6; one optimiser step is compared at 2e-4 relative on the gradient /
parameter fingerprints (fp32 reassociation through ~30 layers of backward).
"""
import numpy as np
import pytest
import torch

from oracle import ppo_oracle as po
from tests import _golden as g


@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_gae_matches_reference(case):
  G = g.load("gae")
  roll, last_value, tlf = g.gae_case(case, G["gae_%s/cfg" % case])
  advs, rets = po.gae(roll["rewards"], roll["values"], roll["terminals"], roll["time_limits"],
                      last_value, 0.99, 0.95, tlf)
  # float64 both sides: exact up to reassociation
  np.testing.assert_allclose(advs, G["gae_%s/advs" % case], rtol=0, atol=1e-12)
  np.testing.assert_allclose(rets, G["gae_%s/rets" % case], rtol=0, atol=1e-12)
  advs, rets = po.discount_reward(roll["rewards"], roll["values"], roll["terminals"],
                                  roll["time_limits"], last_value, 0.99, tlf)
  np.testing.assert_allclose(advs, G["disc_%s/advs" % case], rtol=0, atol=1e-12)
  np.testing.assert_allclose(rets, G["disc_%s/rets" % case], rtol=0, atol=1e-12)


def _oracle(family, **kw):
  S, A = g.FAMILIES[family]
  pf_np, vf_np = g.family_weights(family)
  pf, vf = po.sd_to_torch(pf_np, vf_np)
  return po.PPOOracle(family, pf, vf, S, **kw)


@pytest.mark.parametrize("family", ["loco", "nature", "mlp", "vit", "nvo"])
def test_forward_matches_reference(family):
  G = g.load(family)
  orc = _oracle(family)
  obs, acts = g.fwd_inputs(family)
  obs_t, acts_t = torch.tensor(obs), torch.tensor(acts)
  mean = orc.policy(obs_t)
  value = orc.values(obs_t)
  lp, ent, _ = po.gaussian_update(mean, orc.pf["logstd"], acts_t)
  assert g.rel_err(mean.numpy(), G["fwd/mean"]) < 2e-5
  assert g.rel_err(value.numpy(), G["fwd/value"]) < 2e-5
  assert g.rel_err(lp.numpy(), G["fwd/log_prob"]) < 2e-5
  assert g.rel_err(ent.numpy(), G["fwd/ent"]) < 1e-6
  if "fwd/value_1d" in G:
    assert g.rel_err(orc.values(obs_t[:1]).numpy()[0], G["fwd/value_1d"]) < 2e-5


@pytest.mark.parametrize("family", ["loco", "nature", "mlp", "vit", "nvo"])
def test_one_update_matches_reference(family):
  G = g.load(family)
  orc = _oracle(family, batch_size=16, opt_epochs=1)
  info = orc.update(g.update_inputs(family))
  g.check_info(G, "upd/info", info, rtol=2e-4)
  # the shared encoder's .grad is overwritten by the actor backward in the reference, so the
  # critic gradient is only observable on critic-exclusive tensors (and through upd/vf below)
  g.check_summary(G, "upd/vgrad", [(k, v.numpy()) for k, v in orc._last["vgrads"].items()
                                   if k not in orc.pf], 5e-4)
  g.check_summary(G, "upd/pgrad", [(k, v.numpy()) for k, v in orc._last["pgrads"].items()], 5e-4)
  g.check_summary(G, "upd/pf", [(k, v.numpy()) for k, v in orc.pf.items()], 2e-4)
  g.check_summary(G, "upd/vf", [(k, v.numpy()) for k, v in orc.vf.items()], 2e-4)


@pytest.mark.parametrize("family", ["loco", "mlp"])
def test_clipped_value_loss_matches_reference(family):
  G = g.load(family)
  orc = _oracle(family, batch_size=16, opt_epochs=1, clipped_value_loss=True)
  info = orc.update(g.update_inputs(family))
  g.check_info(G, "updclip/info", info, rtol=2e-4)
  g.check_summary(G, "updclip/vf", [(k, v.numpy()) for k, v in orc.vf.items()], 2e-4)


@pytest.mark.parametrize("family", ["loco", "nature", "mlp", "vit", "nvo"])
def test_update_per_epoch_matches_reference(family):
  G = g.load(family)
  orc = _oracle(family, batch_size=16, opt_epochs=2)
  orc.current_epoch = 30
  roll = g.epoch_inputs(family)
  advs, rets, infos = orc.update_per_epoch(roll, G["epoch/perms"])
  assert g.rel_err(advs, G["epoch/advs"]) < 1e-5
  assert g.rel_err(rets, G["epoch/rets"]) < 1e-5
  assert len(infos) == int(G["epoch/n_infos"])
  for i, info in enumerate(infos):
    # ratio extremes amplify parameter noise by (a-mu)/sigma^2 ~ 64x: looser there
    g.check_info(G, "epoch/info%d" % i, info, rtol=2e-3, atol=1e-4)
  g.check_summary(G, "epoch/pf", [(k, v.numpy()) for k, v in orc.pf.items()], 5e-4)
  g.check_summary(G, "epoch/vf", [(k, v.numpy()) for k, v in orc.vf.items()], 5e-4)
  g.check_summary(G, "epoch/target", [(k, v.numpy()) for k, v in orc.target_pf.items()], 5e-4)
  lr = 1e-4 * (1 - 30 / 1500.0)
  np.testing.assert_allclose(G["epoch/lr"], [lr, lr], rtol=1e-12)


def test_a2c_update_matches_reference():
  """oracle A2C restatement vs the live reference's two A2C.update calls (oracle/make_golden_a2c.py)"""
  from oracle import make_golden_a2c as mk
  from oracle import synth
  G = g.load("a2c_mlp")
  pf_np, vf_np = synth.make_family_weights(1000, "mlp", mk.S, mk.A)
  pf, vf = po.sd_to_torch(pf_np, vf_np, shared_prefixes=())
  orc = po.A2COracle("mlp", pf, vf, mk.S)
  for i, b in enumerate(mk.batches()):
    info = orc.update(b)
    for k, v in info.items():
      want = float(G["info%d/%s" % (i, k)])
      assert abs(v - want) <= 1e-6 + 2e-5 * abs(want), (i, k, v, want)
  g.check_summary(G, "pf", [(k, v.numpy()) for k, v in pf.items()], 2e-5, "a2c oracle")
  g.check_summary(G, "vf", [(k, v.numpy()) for k, v in vf.items()], 2e-5, "a2c oracle")
