"""Builders shared by the GPU tests, smoke() and bench.py: product modules + PPO wired the way
the reference starters wire them (starter/ppo_locotransformer.py:79-118 etc.)."""
import tempfile

import numpy as np
import torch

import vision4leg_b200.networks as networks
import vision4leg_b200.policies as policies
from vision4leg_b200.algo import PPO
from vision4leg_b200.replay_buffers import OnPolicyReplayBuffer


def build_nets(family, S, A):
  net = {"transformer_params": [[1, 256], [1, 256]], "append_hidden_shapes": [256, 256],
         "base_type": networks.MLPBase}
  if family == "loco":
    enc = networks.LocoTransformerEncoder(in_channels=4, state_input_dim=S, hidden_shapes=[256, 256],
                                          visual_dim=256)
    pf = policies.GaussianContPolicyLocoTransformer(
      encoder=enc, state_input_shape=S, visual_input_shape=(4, 64, 64), output_shape=A, **net)
    vf = networks.LocoTransformer(
      encoder=enc, state_input_shape=S, visual_input_shape=(4, 64, 64), output_shape=1, **net)
  elif family == "nature":
    enc = networks.NatureFuseEncoder(in_channels=4, state_input_dim=S, hidden_shapes=[256, 256],
                                     visual_dim=256)
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
  elif family == "mlp":
    net = {"append_hidden_shapes": [256, 256], "hidden_shapes": [256, 256], "base_type": networks.MLPBase}
    pf = policies.GaussianContPolicyBasicBias(input_shape=S, output_shape=A, **net)
    vf = networks.Net(input_shape=(S,), output_shape=1, **net)
    vf.base = pf.base
  else:
    raise ValueError(family)
  return pf, vf


def load_np_sd(module, sd_np):
  sd = module.state_dict()
  assert set(sd.keys()) == set(sd_np.keys()), sorted(set(sd) ^ set(sd_np))
  module.load_state_dict({k: torch.tensor(v) for k, v in sd_np.items()})


class Obj:
  pass


class Box:
  def __init__(self, shape):
    self.shape = shape


class ListLogger:
  def __init__(self):
    self.infos = []

  def add_update_info(self, info):
    self.infos.append(dict(info))


def fill_buffer(roll, T, E, time_limit_filter=True):
  buf = OnPolicyReplayBuffer(env_nums=E, max_replay_buffer_size=T * E, time_limit_filter=time_limit_filter)
  for t in range(T):
    nxt = roll["obs"][t + 1] if t + 1 < T else roll["last_obs"]
    buf.add_sample({"obs": roll["obs"][t], "next_obs": nxt, "acts": roll["acts"][t],
                    "values": roll["values"][t], "rewards": roll["rewards"][t],
                    "terminals": roll["terminals"][t], "time_limits": roll["time_limits"][t]})
  return buf


def make_ppo(pf, vf, buf, A, batch_size, epoch_frames, opt_epochs, device="cuda:0", **kw):
  env = Obj(); env.action_space = Box((A,))
  collector = Obj(); collector.epoch_frames = epoch_frames
  logger = ListLogger()
  args = dict(plr=1e-4, vlr=1e-4, clip_para=0.2, opt_epochs=opt_epochs, tau=0.95, shuffle=True,
              entropy_coeff=0.005, discount=0.99, num_epochs=1500, batch_size=batch_size,
              save_interval=100, eval_interval=10, gae=True)
  args.update(kw)
  agent = PPO(pf=pf, vf=vf, env=env, replay_buffer=buf, collector=collector, logger=logger,
              device=device, save_dir=tempfile.mkdtemp(), **args)
  return agent, logger
