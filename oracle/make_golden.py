"""Generate tests/golden/*.npz from the LIVE reference (run in the build container only).

  PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

/root/reference is imported unmodified (with a stub `gym` module, the only missing import on
the path, used once for an isinstance check at reference torchrl/algo/rl_algo.py:36). It does
not exist on the GPU box: the fixtures written here are what travels. Inputs and weights are
regenerated from seeds by oracle/synth.py, so the fixtures hold only the reference's OUTPUTS.

TEST INFRASTRUCTURE — not imported by the product.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

from oracle import synth   # imported BEFORE the repo root is dropped from sys.path

REF = os.environ.get("V4L_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

FAMILIES = {
  # family: (S, A) — (93,12) is the north-star shape, (84,6) the shipped-JSON shape
  "loco": (93, 12),
  "nature": (84, 6),
  "mlp": (84, 6),
  # vision-only variants (starter/ppo_locotransformer_vision_only.py, ppo_nature_cnn_vision_only.py)
  "vit": (0, 6),
  "nvo": (0, 6),
}


def import_reference():
  sys.dont_write_bytecode = True
  gym = types.ModuleType("gym")
  spaces = types.ModuleType("gym.spaces")

  class Box:
    def __init__(self, low=-1.0, high=1.0, shape=(1,)):
      self.shape = shape
  spaces.Box = Box
  gym.spaces = spaces
  sys.modules["gym"] = gym
  sys.modules["gym.spaces"] = spaces
  # make sure the repo's own torchrl shim does not shadow the reference here
  repo = os.path.dirname(os.path.dirname(OUT))
  sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != repo]
  for m in [m for m in sys.modules if m == "torchrl" or m.startswith("torchrl.")]:
    del sys.modules[m]
  sys.path.insert(0, REF)
  import torchrl.networks as networks
  import torchrl.policies as policies
  from torchrl.algo import PPO
  from torchrl.replay_buffers.on_policy import OnPolicyReplayBuffer
  assert networks.__file__.startswith(REF), networks.__file__
  return networks, policies, PPO, OnPolicyReplayBuffer, Box


def build_reference_nets(networks, policies, family, S, A):
  """Mirrors starter/ppo_locotransformer.py:79-100, ppo_nature_cnn.py:81-102, ppo_state.py:90-104."""
  net = {"transformer_params": [[1, 256], [1, 256]], "append_hidden_shapes": [256, 256],
         "base_type": networks.MLPBase}
  if family == "loco":
    enc = networks.LocoTransformerEncoder(in_channels=4, state_input_dim=S,
                                          hidden_shapes=[256, 256], visual_dim=256)
    pf = policies.GaussianContPolicyLocoTransformer(
      encoder=enc, state_input_shape=S, visual_input_shape=(4, 64, 64), output_shape=A, **net)
    vf = networks.LocoTransformer(
      encoder=enc, state_input_shape=S, visual_input_shape=(4, 64, 64), output_shape=1, **net)
  elif family == "nature":
    enc = networks.NatureFuseEncoder(in_channels=4, state_input_dim=S,
                                     hidden_shapes=[256, 256], visual_dim=256)
    pf = policies.GaussianContPolicyImpalaEncoderProj(
      encoder=enc, state_input_shape=S, visual_input_shape=(4, 64, 64), output_shape=A, **net)
    vf = networks.ImpalaEncoderProjNet(
      encoder=enc, state_input_shape=S, visual_input_shape=(4, 64, 64), output_shape=1, **net)
  elif family == "vit":
    enc = networks.TransformerEncoder(in_channels=4, hidden_shapes=[256, 256], visual_dim=256)
    pf = policies.GaussianContPolicyTransformer(
      encoder=enc, visual_input_shape=(4, 64, 64), output_shape=A, **net)
    vf = networks.Transformer(encoder=enc, visual_input_shape=(4, 64, 64), output_shape=1, **net)
  elif family == "nvo":
    enc = networks.NatureEncoder(in_channels=4, hidden_shapes=[256, 256], visual_dim=256)
    pf = policies.GaussianContPolicyNatureEncoderProj(
      encoder=enc, visual_input_shape=(4, 64, 64), output_shape=A, **net)
    vf = networks.NatureEncoderProjNet(encoder=enc, visual_input_shape=(4, 64, 64), output_shape=1, **net)
  else:
    net = {"append_hidden_shapes": [256, 256], "hidden_shapes": [256, 256],
           "base_type": networks.MLPBase}
    pf = policies.GaussianContPolicyBasicBias(input_shape=S, output_shape=A, **net)
    vf = networks.Net(input_shape=(S,), output_shape=1, **net)
    vf.base = pf.base
  return pf, vf


def load_np_sd(module, sd_np):
  sd = module.state_dict()
  assert set(sd.keys()) == set(sd_np.keys()), (sorted(set(sd) ^ set(sd_np)))
  for k in sd:
    assert tuple(sd[k].shape) == sd_np[k].shape, (k, sd[k].shape, sd_np[k].shape)
  module.load_state_dict({k: torch.tensor(v) for k, v in sd_np.items()})


def summarize(prefix, named, out):
  """Compact fingerprint of a set of tensors: float64 sum, abs-sum and a strided sample."""
  for k, t in named:
    a = t.detach().double().numpy().ravel()
    out["%s/%s/sum" % (prefix, k)] = np.array(a.sum())
    out["%s/%s/abs" % (prefix, k)] = np.array(np.abs(a).sum())
    step = max(1, a.size // 61)
    out["%s/%s/sample" % (prefix, k)] = a[::step][:64].astype(np.float32)


class _Obj:
  pass


def make_ppo(PPO, Box, pf, vf, buf, A, batch_size, epoch_frames, opt_epochs, **kw):
  env = _Obj(); env.action_space = Box(shape=(A,))
  collector = _Obj(); collector.epoch_frames = epoch_frames
  logger = _Obj(); logger.infos = []
  logger.add_update_info = lambda info: logger.infos.append(dict(info))
  agent = PPO(pf=pf, vf=vf, plr=1e-4, vlr=1e-4, clip_para=0.2, opt_epochs=opt_epochs, tau=0.95,
              shuffle=True, entropy_coeff=0.005, env=env, replay_buffer=buf, collector=collector,
              logger=logger, discount=0.99, num_epochs=1500, batch_size=batch_size, device="cpu",
              save_interval=100, eval_interval=10, save_dir=tempfile.mkdtemp(), gae=True, **kw)
  return agent, logger


def fill_buffer(Buffer, roll, T, E, time_limit_filter=True):
  buf = Buffer(env_nums=E, max_replay_buffer_size=T * E, time_limit_filter=time_limit_filter)
  for t in range(T):
    nxt = roll["obs"][t + 1] if t + 1 < T else roll["last_obs"]
    buf.add_sample({"obs": roll["obs"][t], "next_obs": nxt, "acts": roll["acts"][t],
                    "values": roll["values"][t], "rewards": roll["rewards"][t],
                    "terminals": roll["terminals"][t], "time_limits": roll["time_limits"][t]})
  return buf


def golden_gae(Buffer):
  out = {}
  for name, (T, E, p_term, p_tl, tlf) in {
      "a": (64, 4, 0.1, 0.0, True), "b": (97, 3, 0.05, 0.1, True), "c": (33, 8, 0.2, 0.1, False),
      "d": (1, 2, 0.5, 0.0, True)}.items():
    roll = synth.make_rollout(100 + T, T, E, 5, 2, with_img=False, p_term=p_term, time_limit_p=p_tl)
    buf = fill_buffer(Buffer, roll, T, E, tlf)
    rng = np.random.default_rng(7)
    last_value = rng.standard_normal((E, 1))
    buf.generalized_advantage_estimation(last_value, 0.99, 0.95)
    out["gae_%s/cfg" % name] = np.array([T, E, p_term, p_tl, float(tlf)])
    out["gae_%s/advs" % name] = buf._advs.copy()
    out["gae_%s/rets" % name] = buf._estimate_returns.copy()
    buf.discount_reward(last_value, 0.99)
    out["disc_%s/advs" % name] = buf._advs.copy()
    out["disc_%s/rets" % name] = buf._estimate_returns.copy()
  np.savez_compressed(os.path.join(OUT, "gae.npz"), **out)
  print("gae.npz", len(out))


def golden_family(mods, family):
  networks, policies, PPO, Buffer, Box = mods
  S, A = FAMILIES[family]
  with_img = family != "mlp"
  out = {"cfg": np.array([S, A])}
  pf_np, vf_np = synth.make_family_weights(1000, family, S, A)

  # ---- (1) forward + policy.update outputs
  pf, vf = build_reference_nets(networks, policies, family, S, A)
  load_np_sd(pf, pf_np); load_np_sd(vf, vf_np)
  roll = synth.make_rollout(2000, 4, 2, S, A, with_img=with_img)
  obs = torch.tensor(roll["obs"].reshape(8, -1))
  acts = torch.tensor(roll["acts"].reshape(8, -1))
  with torch.no_grad():
    upd = pf.update(obs, acts)
    out["fwd/mean"] = upd["mean"].numpy()
    out["fwd/log_prob"] = upd["log_prob"].numpy()
    out["fwd/ent"] = upd["ent"].numpy()
    out["fwd/value"] = vf(obs).numpy()
    out["fwd/eval_act"] = pf.eval_act(obs)
    try:
      out["fwd/value_1d"] = vf(obs[0]).numpy()    # 1-D input path
    except TypeError:
      # NatureEncoder.forward builds torch.Size([np.prod(())]) = [1.0] for un-batched input, which
      # numpy >= 2 / torch 2.11 reject: the reference itself cannot run this case for `nvo`
      pass

  # ---- (2) one PPO.update(batch) on B=16 (T=4 rows x E=4)
  T, E, Bm = 4, 4, 16
  roll = synth.make_rollout(3000, T, E, S, A, with_img=with_img, p_term=0.2)
  buf = fill_buffer(Buffer, roll, T, E)
  agent, logger = make_ppo(PPO, Box, pf, vf, buf, A, Bm, T * E, 1)
  agent.current_epoch = 0
  rng = np.random.default_rng(11)
  batch = {"obs": roll["obs"].reshape(Bm, -1), "acts": roll["acts"].reshape(Bm, -1),
           "advs": rng.standard_normal((Bm, 1)), "estimate_returns": rng.standard_normal((Bm, 1)),
           "values": roll["values"].reshape(Bm, -1)}
  info = agent.update(batch)
  for k, v in info.items():
    out["upd/info/" + k] = np.array(v)
  summarize("upd/pf", pf.state_dict().items(), out)
  summarize("upd/vf", vf.state_dict().items(), out)
  summarize("upd/pgrad", [(k, p.grad) for k, p in pf.named_parameters()], out)
  summarize("upd/vgrad", [(k, p.grad) for k, p in vf.named_parameters()], out)

  # ---- (2b) clipped value loss variant, fresh weights
  pf2, vf2 = build_reference_nets(networks, policies, family, S, A)
  load_np_sd(pf2, pf_np); load_np_sd(vf2, vf_np)
  agent2, _ = make_ppo(PPO, Box, pf2, vf2, buf, A, Bm, T * E, 1, clipped_value_loss=True)
  info2 = agent2.update(batch)
  for k, v in info2.items():
    out["updclip/info/" + k] = np.array(v)
  summarize("updclip/vf", vf2.state_dict().items(), out)

  # ---- (3) full update_per_epoch: T=8,E=4, batch 16 -> 2 minibatches x 2 opt epochs, epoch 30
  T, E, Bm, OE = 8, 4, 16, 2
  pf3, vf3 = build_reference_nets(networks, policies, family, S, A)
  load_np_sd(pf3, pf_np); load_np_sd(vf3, vf_np)
  roll = synth.make_rollout(4000, T, E, S, A, with_img=with_img, p_term=0.15, time_limit_p=0.1)
  buf = fill_buffer(Buffer, roll, T, E)
  agent3, logger3 = make_ppo(PPO, Box, pf3, vf3, buf, A, Bm, T * E, OE)
  agent3.current_epoch = 30
  np.random.seed(1234)
  perms = np.stack([np.random.permutation(T) for _ in range(OE)])
  np.random.seed(1234)
  agent3.update_per_epoch()
  out["epoch/perms"] = perms
  out["epoch/advs"] = buf._advs.copy()
  out["epoch/rets"] = buf._estimate_returns.copy()
  for i, info in enumerate(logger3.infos):
    for k, v in info.items():
      out["epoch/info%d/%s" % (i, k)] = np.array(v)
  out["epoch/n_infos"] = np.array(len(logger3.infos))
  summarize("epoch/pf", pf3.state_dict().items(), out)
  summarize("epoch/vf", vf3.state_dict().items(), out)
  summarize("epoch/target", agent3.target_pf.state_dict().items(), out)
  out["epoch/lr"] = np.array([agent3.pf_optimizer.param_groups[0]["lr"],
                              agent3.vf_optimizer.param_groups[0]["lr"]])

  # ---- (4) init fingerprint: torch.manual_seed(0) construction (init-parity is informational)
  torch.manual_seed(0)
  pf4, vf4 = build_reference_nets(networks, policies, family, S, A)
  summarize("init/pf", pf4.state_dict().items(), out)
  summarize("init/vf", vf4.state_dict().items(), out)
  out["init/pf_keys"] = np.array(list(pf4.state_dict().keys()))
  out["init/vf_keys"] = np.array(list(vf4.state_dict().keys()))
  out["init/pf_param_order"] = np.array([k for k, _ in pf4.named_parameters()])
  out["init/vf_param_order"] = np.array([k for k, _ in vf4.named_parameters()])

  np.savez_compressed(os.path.join(OUT, "%s.npz" % family), **out)
  print("%s.npz" % family, len(out), "arrays")


def main():
  os.makedirs(OUT, exist_ok=True)
  torch.set_num_threads(8)
  mods = import_reference()
  golden_gae(mods[3])
  for fam in FAMILIES:
    golden_family(mods, fam)


if __name__ == "__main__":
  main()
