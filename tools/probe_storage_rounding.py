"""How far does a PPO epoch drift from the fp32 reference trajectory when ONLY the storage precision of
activations / weights changes?  (CPU experiment, no CUDA: tests-side tooling, imports the oracle.)

The oracle's LocoTransformer forward is re-stated with a straight-through rounding q(x) at exactly the
points where the tensor-core tier stores an fp16 tensor (weights as MMA operands, every activation that
leaves a kernel); all arithmetic, the whole backward pass, the losses and Adam stay fp32.  Running
PPOOracle.update_per_epoch with that forward and comparing with the plain fp32 oracle isolates what
16-bit STORAGE alone does to the trajectory (ReLU gates that flip for pre-activations within the rounding
error of 0, amplified by Adam's sign-like first steps) from anything specific to the CUDA kernels.

  python tools/probe_storage_rounding.py [f16|bf16|eps1e-6] [B] [minibatches] [opt_epochs]
"""
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ppo_oracle as po, synth   # noqa: E402
from tests import _golden as g   # noqa: E402

MODE = sys.argv[1] if len(sys.argv) > 1 else "f16"
DT = {"f16": torch.float16, "bf16": torch.bfloat16}.get(MODE)
EPS = float(MODE[3:]) if MODE.startswith("eps") else 0.0     # "eps1e-6": relative noise of that size instead


def q(x):
  if DT is None:      # fp32-rounding-sized perturbation (what a different summation order does)
    return x + (x * EPS * torch.randn_like(x)).detach()
  return x + (x.to(DT).float() - x).detach()


def lin(x, P, n, relu):
  y = F.linear(x, q(P[n + "weight"]), P[n + "bias"])
  return F.relu(y) if relu else y


def nature_cnn(P, prefix, img):
  x = q(F.relu(F.conv2d(q(img), q(P[prefix + "layers.0.weight"]), P[prefix + "layers.0.bias"], stride=4)))
  x = q(F.relu(F.conv2d(x, q(P[prefix + "layers.2.weight"]), P[prefix + "layers.2.bias"], stride=2)))
  return q(F.relu(F.conv2d(x, q(P[prefix + "layers.4.weight"]), P[prefix + "layers.4.bias"], stride=1)))


def layer(P, p, x):
  d = x.shape[-1]
  qkv = q(lin(x, P, p + "self_attn.in_proj_", False))
  qq, k, v = qkv.split(d, -1)
  pr = q(torch.softmax(torch.matmul(qq, k.transpose(-1, -2)) / math.sqrt(d), -1))
  o = q(torch.matmul(pr, v))
  o = F.linear(o, q(P[p + "self_attn.out_proj.weight"]), P[p + "self_attn.out_proj.bias"])
  h = q(po.layer_norm(x + o, P[p + "norm1.weight"], P[p + "norm1.bias"]))
  f = lin(q(lin(h, P, p + "linear1.", True)), P, p + "linear2.", False)
  return q(po.layer_norm(h + f, P[p + "norm2.weight"], P[p + "norm2.bias"]))


def loco_rounded(P, x, S, n_head=1):
  B = x.shape[0]
  state, img = q(x[:, :S]), x[:, S:].reshape(B, 4, 64, 64)
  feat = nature_cnn(P, "encoder.depth_visual_base.", img)
  up = q(F.conv2d(feat, q(P["encoder.depth_up_conv.weight"]), P["encoder.depth_up_conv.bias"]))
  s = q(lin(q(lin(state, P, "encoder.base.seq_fcs.0.", True)), P, "encoder.base.seq_fcs.2.", True))
  st = q(lin(s, P, "encoder.state_projector.projection.0.", True))
  tok = torch.cat([st[:, None], up.reshape(B, 64, 16).permute(0, 2, 1)], 1)
  for l in range(2):
    tok = layer(P, "visual_append_layers.%d." % l, tok)
  pooled = q(torch.cat([tok[:, 0], tok[:, 1:17].mean(1)], -1))
  h2 = q(lin(q(lin(pooled, P, "visual_seq_append_fcs.0.", True)), P, "visual_seq_append_fcs.2.", True))
  return lin(h2, P, "visual_seq_append_fcs.4.", False)


def main():
  B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
  n_mb = int(sys.argv[3]) if len(sys.argv) > 3 else 8
  opt_epochs = int(sys.argv[4]) if len(sys.argv) > 4 else 2
  S, A = g.FAMILIES["loco"]
  E = 8
  T = n_mb * B // E
  roll = synth.make_rollout(31, T, E, S, A, p_term=1.0 / 200)
  np.random.seed(77)
  perms = np.stack([np.random.permutation(T) for _ in range(opt_epochs)])
  pf_np, vf_np = g.family_weights("loco")
  runs = {}
  for name in ("fp32", MODE):
    opf, ovf = po.sd_to_torch(pf_np, vf_np)
    orc = po.PPOOracle("loco", opf, ovf, S, batch_size=B, opt_epochs=opt_epochs)
    orc.current_epoch = 7
    if name != "fp32":
      orc.fwd = loco_rounded
    _, _, infos = orc.update_per_epoch(roll, perms)
    held = torch.tensor(synth.make_obs(np.random.default_rng(5), 256, S))
    runs[name] = (infos, orc.policy(held).numpy(), orc.values(held).numpy())
  ref, got = runs["fp32"], runs[MODE]
  print("storage = %s, B = %d, %d minibatches x %d opt-epochs: worst relative deviation from the fp32 trajectory" %
        (MODE, B, n_mb, opt_epochs))
  for k in g.INFO_KEYS:
    w = max(abs(a[k] - b[k]) / (abs(b[k]) + 1e-4) for a, b in zip(got[0], ref[0]))
    print("  %-22s %.3e" % (k, w))
  print("  %-22s %.3e" % ("heldout/mean", g.rel_err(got[1], ref[1])))
  print("  %-22s %.3e" % ("heldout/value", g.rel_err(got[2], ref[2])))


if __name__ == "__main__":
  torch.set_num_threads(max(1, (os.cpu_count() or 2) - 1))
  main()
