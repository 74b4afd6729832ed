"""Deterministic synthetic rollouts and seeded weights for tests, smoke(), bench.py and the tools.

A neutral package (no import of oracle/, none of the product); the product (vision4leg_b200/) never imports it.

Everything here is generated from numpy's PCG64 stream (stable across numpy versions), so
the golden fixtures under tests/golden/ only need to store *outputs*: the same seed
regenerates the same inputs and weights on the GPU box.

Input distributions follow SURVEY.md §8(d):
  * proprio part  obs[:, :S] ~ clip(N(0,1), +-10)   (post-normaliser range, reference
    torchrl/env/base_wrapper.py:91-94)
  * depth part    4 stacked 64x64 frames (sqrt(log(d+1)) - 1.25)/0.425, d ~ U[0.3, 10]
    (reference vision4leg/envs/locomotion_gym_env_with_rich_information.py:623-654)
  * acts ~ 0.15 N(0,1) (near the policy mean), rewards ~ N(0,1), terminals ~ Bernoulli(1/500), time_limits = 0.
"""
import math

import numpy as np

IMG_ELEMS = 4 * 64 * 64


def make_obs(rng, n, S, dtype=np.float32):
  """[n, S + 16384] observation rows: proprio | CHW depth stack."""
  obs = np.empty((n, S + IMG_ELEMS), dtype=dtype)
  if S:
    obs[:, :S] = np.clip(rng.standard_normal((n, S)), -10, 10)
  d = rng.uniform(0.3, 10.0, size=(n, IMG_ELEMS))
  obs[:, S:] = (np.sqrt(np.log(d + 1.0)) - 1.25) / 0.425
  return obs


def make_rollout(seed, T, E, S, A, with_img=True, dtype=np.float32, p_term=1.0 / 500,
                 time_limit_p=0.0):
  """Synthetic rollout buffer contents, arrays shaped like the reference buffer
  (torchrl/replay_buffers/base.py:20-30): [T, E, dim]; time_limits [T, 1]."""
  rng = np.random.default_rng(seed)
  n = T * E
  if with_img:
    obs = make_obs(rng, n, S, dtype).reshape(T, E, -1)
    last_obs = make_obs(rng, E, S, dtype)
  else:
    obs = np.clip(rng.standard_normal((T, E, S)), -10, 10).astype(dtype)
    last_obs = np.clip(rng.standard_normal((E, S)), -10, 10).astype(dtype)
  out = {
    "obs": obs,
    "last_obs": last_obs,
    # actions near the (near-zero) policy mean at the policy's own scale (sigma = 0.125), i.e. the
    # regime PPO operates in; N(0,1) actions sit 8 sigma out and make exp(lp - lp') chaotic
    "acts": (0.15 * rng.standard_normal((T, E, A))).astype(dtype),
    "rewards": rng.standard_normal((T, E, 1)).astype(dtype),
    "values": rng.standard_normal((T, E, 1)).astype(dtype),
    "terminals": (rng.uniform(size=(T, E, 1)) < p_term).astype(dtype),
    "time_limits": (rng.uniform(size=(T, 1)) < time_limit_p).astype(dtype),
  }
  # reference on_rl_algo.py:24-28 masks the bootstrap value with terminals[T-1]
  out["last_terminals"] = out["terminals"][-1].copy()
  return out


def _fan_in(shape):
  return int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])


def make_weights(seed, spec):
  """spec: ordered list of (name, shape). Returns {name: float32 array}.

  Scales are init-like (1/sqrt(fan_in)) so activations stay O(1) through the net; the exact
  distribution is irrelevant for parity, only determinism matters.
  """
  rng = np.random.default_rng(seed)
  out = {}
  heads = sorted((n for n, _ in spec if "append_fcs." in n and n.endswith("weight")),
                 key=lambda n: int(n.split(".")[-2]))
  last_layer_names = set()
  if heads:
    last_layer_names = {heads[-1], heads[-1][:-6] + "bias"}
  for name, shape in spec:
    shape = tuple(int(s) for s in shape)
    if name.endswith("logstd"):
      w = math.log(0.125) + 0.1 * rng.standard_normal(shape)
    elif ".norm" in name and name.endswith("weight"):
      w = 1.0 + 0.1 * rng.standard_normal(shape)
    elif name.endswith("bias"):
      w = 0.05 * rng.standard_normal(shape)
    else:
      w = rng.standard_normal(shape) * (1.0 / math.sqrt(_fan_in(shape)))
    if name in last_layer_names:
      w = w * 0.05          # reference net_last_init_func is uniform +-3e-3 (init.py:18-19)
    out[name] = np.ascontiguousarray(w, dtype=np.float32)
  return out


# ---------------------------------------------------------------------------------------------
# state_dict key specs of the three policy families (probed from the
# reference modules; make_golden.py asserts they match the live reference state_dicts).
# ---------------------------------------------------------------------------------------------

def _nature_spec(prefix):
  return [
    (prefix + "layers.0.weight", (32, 4, 8, 8)), (prefix + "layers.0.bias", (32,)),
    (prefix + "layers.2.weight", (64, 32, 4, 4)), (prefix + "layers.2.bias", (64,)),
    (prefix + "layers.4.weight", (64, 64, 3, 3)), (prefix + "layers.4.bias", (64,)),
  ]


def _mlp_base_spec(prefix, S, hidden):
  spec, d = [], S
  for i, h in enumerate(hidden):
    spec += [(prefix + "seq_fcs.%d.weight" % (2 * i), (h, d)),
             (prefix + "seq_fcs.%d.bias" % (2 * i), (h,))]
    d = h
  return spec


def _head_spec(prefix, din, hidden, out):
  spec, d = [], din
  for i, h in enumerate(hidden):
    spec += [(prefix + "%d.weight" % (2 * i), (h, d)), (prefix + "%d.bias" % (2 * i), (h,))]
    d = h
  spec += [(prefix + "%d.weight" % (2 * len(hidden)), (out, d)),
           (prefix + "%d.bias" % (2 * len(hidden)), (out,))]
  return spec


def loco_encoder_spec(S, hidden=(256, 256), token_dim=64):
  return (_nature_spec("encoder.depth_visual_base.") +
          [("encoder.depth_up_conv.weight", (token_dim, 64, 1, 1)),
           ("encoder.depth_up_conv.bias", (token_dim,))] +
          _mlp_base_spec("encoder.base.", S, hidden) +
          [("encoder.state_projector.projection.0.weight", (token_dim, hidden[-1])),
           ("encoder.state_projector.projection.0.bias", (token_dim,))])


def loco_net_spec(out, n_layers=2, ff=256, hidden=(256, 256), token_dim=64):
  spec = []
  for l in range(n_layers):
    p = "visual_append_layers.%d." % l
    spec += [
      (p + "self_attn.in_proj_weight", (3 * token_dim, token_dim)),
      (p + "self_attn.in_proj_bias", (3 * token_dim,)),
      (p + "self_attn.out_proj.weight", (token_dim, token_dim)),
      (p + "self_attn.out_proj.bias", (token_dim,)),
      (p + "linear1.weight", (ff, token_dim)), (p + "linear1.bias", (ff,)),
      (p + "linear2.weight", (token_dim, ff)), (p + "linear2.bias", (token_dim,)),
      (p + "norm1.weight", (token_dim,)), (p + "norm1.bias", (token_dim,)),
      (p + "norm2.weight", (token_dim,)), (p + "norm2.bias", (token_dim,)),
    ]
  spec += _head_spec("visual_seq_append_fcs.", 2 * token_dim, hidden, out)
  return spec


def nature_encoder_spec(S, visual_dim=256, hidden=(256, 256)):
  return (_nature_spec("encoder.visual_base.") +
          [("encoder.visual_projector.projection.0.weight", (visual_dim, 1024)),
           ("encoder.visual_projector.projection.0.bias", (visual_dim,))] +
          _mlp_base_spec("encoder.base.", S, hidden))


def nature_net_spec(out, visual_dim=256, hidden=(256, 256), enc_hidden=256):
  return _head_spec("seq_append_fcs.", visual_dim + enc_hidden, hidden, out)


def mlp_base_spec(S, hidden=(256, 256)):
  return _mlp_base_spec("base.", S, hidden)


def mlp_net_spec(out, din=256, hidden=(256, 256)):
  return _head_spec("seq_append_fcs.", din, hidden, out)


def vit_encoder_spec(token_dim=64):
  return (_nature_spec("encoder.depth_visual_base.") +
          [("encoder.depth_up_conv.weight", (token_dim, 64, 1, 1)), ("encoder.depth_up_conv.bias", (token_dim,))])


def vit_net_spec(out, n_layers=2, ff=256, hidden=(256, 256), token_dim=64):
  spec = [kv for kv in loco_net_spec(out, n_layers, ff, hidden, token_dim) if not kv[0].startswith("visual_seq_append_fcs.")]
  return spec + _head_spec("visual_seq_append_fcs.", token_dim, hidden, out)


def make_family_weights(seed, family, S, A):
  """Returns (pf_sd, vf_sd) numpy state dicts with the *shared* tensors being the same
  arrays in both (encoder.* for loco/nature — reference starter/ppo_locotransformer.py:79-100;
  base.* for mlp — starter/ppo_state.py:104)."""
  if family == "loco":
    enc = make_weights(seed, loco_encoder_spec(S))
    pf = make_weights(seed + 1, [("logstd", (A,))] + loco_net_spec(A))
    vf = make_weights(seed + 2, loco_net_spec(1))
  elif family == "nature":
    enc = make_weights(seed, nature_encoder_spec(S))
    pf = make_weights(seed + 1, [("logstd", (A,))] + nature_net_spec(A))
    vf = make_weights(seed + 2, nature_net_spec(1))
  elif family == "vit":
    enc = make_weights(seed, vit_encoder_spec())
    pf = make_weights(seed + 1, [("logstd", (A,))] + vit_net_spec(A))
    vf = make_weights(seed + 2, vit_net_spec(1))
  elif family == "nvo":
    enc = make_weights(seed, _nature_spec("encoder."))
    pf = make_weights(seed + 1, [("logstd", (A,))] + _head_spec("seq_append_fcs.", 1024, (256, 256), A))
    vf = make_weights(seed + 2, _head_spec("seq_append_fcs.", 1024, (256, 256), 1))
  elif family == "mlp":
    enc = make_weights(seed, mlp_base_spec(S))
    pf = make_weights(seed + 1, [("logstd", (A,))] + mlp_net_spec(A))
    vf = make_weights(seed + 2, mlp_net_spec(1))
  else:
    raise ValueError(family)
  pf_sd = dict(pf); pf_sd.update(enc)
  vf_sd = dict(vf); vf_sd.update(enc)
  return pf_sd, vf_sd
