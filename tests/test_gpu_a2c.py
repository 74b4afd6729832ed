"""A2C.update on the device against the fixture produced by the live reference
(oracle/make_golden_a2c.py -> tests/golden/a2c_mlp.npz): two updates on the state-only family with separate actor
and critic trunks (the only configuration the reference's A2C can step, see the generator's header).
fp32 CUDA forward / backward kernels under torch.autograd + torch.optim.Adam: parameters within 2e-4 of the
reference's after two Adam steps (first steps are sign-like, lr 3e-4, so 1e-3 of the weight scale is one step),
losses within 1e-4."""
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_a2c_update_matches_reference_fixture():
  from oracle import make_golden_a2c as mk
  from oracle import synth
  from tests import _golden as g
  from benchutil.harness import Obj, Box, ListLogger, load_np_sd
  import vision4leg_b200.networks as networks
  import vision4leg_b200.policies as policies
  from vision4leg_b200.algo import A2C
  G = g.load("a2c_mlp")
  dev = torch.device("cuda", 0)
  net = {"append_hidden_shapes": [256, 256], "hidden_shapes": [256, 256], "base_type": networks.MLPBase}
  pf = policies.GaussianContPolicyBasicBias(input_shape=mk.S, output_shape=mk.A, **net)
  vf = networks.Net(input_shape=(mk.S,), output_shape=1, **net)
  pf_np, vf_np = synth.make_family_weights(1000, "mlp", mk.S, mk.A)
  load_np_sd(pf, pf_np); load_np_sd(vf, vf_np)
  env = Obj(); env.action_space = Box((mk.A,))
  collector = Obj(); collector.epoch_frames = mk.B
  agent = A2C(pf=pf, vf=vf, plr=3e-4, vlr=3e-4, entropy_coeff=0.001, env=env, replay_buffer=Obj(),
              collector=collector, logger=ListLogger(), discount=0.99, num_epochs=10, batch_size=mk.B, device=dev,
              save_interval=100, eval_interval=10, save_dir=tempfile.mkdtemp())
  for i, b in enumerate(mk.batches()):
    info = agent.update(b)
    for k, v in info.items():
      want = float(G["info%d/%s" % (i, k)])
      assert abs(float(v) - want) <= 1e-5 + 1e-4 * abs(want), (i, k, v, want)
  torch.cuda.synchronize(dev)
  w1 = g.check_summary(G, "pf", [(k, v.detach().cpu().numpy()) for k, v in pf.state_dict().items()], 2e-4, "a2c")
  w2 = g.check_summary(G, "vf", [(k, v.detach().cpu().numpy()) for k, v in vf.state_dict().items()], 2e-4, "a2c")
  print("A2C two updates vs reference fixture: worst parameter fingerprint error %.2e" % max(w1, w2))
