"""Generate tests/golden/a2c_mlp.npz from the LIVE reference A2C.update (build container only).

  PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_a2c

Two A2C.update calls (reference torchrl/algo/on_policy/a2c.py:42-107) on the state-only family with SEPARATE actor
and critic trunks — with a shared trunk the reference's step order raises inside torch (the actor's step modifies
weights the critic's graph saved), so that is the only configuration the reference can run.
TEST INFRASTRUCTURE — not imported by the product.
"""
import os
import tempfile

import numpy as np
import torch

from oracle import synth
from oracle import make_golden as mg

S, A, B = 84, 6, 16


def batches():
  roll = synth.make_rollout(5100, 4, 4, S, A, with_img=False)
  rng = np.random.default_rng(23)
  out = []
  for _ in range(2):
    out.append({"obs": roll["obs"].reshape(B, -1) + rng.standard_normal((B, S)).astype(np.float32) * 0.1,
                "acts": roll["acts"].reshape(B, -1), "advs": rng.standard_normal((B, 1)),
                "estimate_returns": rng.standard_normal((B, 1))})
  return out


def main():
  torch.set_num_threads(4)
  bs = batches()                       # before the reference takes over sys.path
  pf_np, vf_np = synth.make_family_weights(1000, "mlp", S, A)
  networks, policies, PPO, Buffer, Box = mg.import_reference()
  from torchrl.algo import A2C
  net = {"append_hidden_shapes": [256, 256], "hidden_shapes": [256, 256], "base_type": networks.MLPBase}
  pf = policies.GaussianContPolicyBasicBias(input_shape=S, output_shape=A, **net)
  vf = networks.Net(input_shape=(S,), output_shape=1, **net)
  mg.load_np_sd(pf, pf_np); mg.load_np_sd(vf, vf_np)
  env = mg._Obj(); env.action_space = Box(shape=(A,))
  collector = mg._Obj(); collector.epoch_frames = B
  logger = mg._Obj(); logger.add_update_info = lambda info: None
  agent = A2C(pf=pf, vf=vf, plr=3e-4, vlr=3e-4, entropy_coeff=0.001, env=env, replay_buffer=mg._Obj(),
              collector=collector, logger=logger, discount=0.99, num_epochs=10, batch_size=B, device="cpu",
              save_interval=100, eval_interval=10, save_dir=tempfile.mkdtemp())
  out = {}
  for i, b in enumerate(bs):
    info = agent.update(b)
    for k, v in info.items():
      out["info%d/%s" % (i, k)] = np.array(v)
  mg.summarize("pf", pf.state_dict().items(), out)
  mg.summarize("vf", vf.state_dict().items(), out)
  np.savez_compressed(os.path.join(mg.OUT, "a2c_mlp.npz"), **out)
  print("a2c_mlp.npz", len(out), "arrays")


if __name__ == "__main__":
  main()
