"""Encoder modules of the visuo-proprioceptive policies — PARAMETER CONTAINERS.

They own ordinary nn.Parameters under exactly the reference's state_dict keys and shapes
(reference torchrl/networks/base.py) and are constructed in the reference's
order so that seeded initialisation matches.  They do not compute: the arithmetic of a whole
network (encoder + transformer + head) is one CUDA layer plan (vision4leg_b200/engine.py)
driven by the owning net in nets.py.  Calling an encoder on its own therefore raises.
"""
import numpy as np
import torch
import torch.nn as nn

from . import init
from .._lib import V4LError


def _container_only(name):
  raise V4LError(
    "%s holds parameters only; run it through the owning network (networks.LocoTransformer, "
    "ImpalaEncoderProjNet, Net, ...) whose forward is the CUDA layer plan" % name)


class MLPBase(nn.Module):
  """Linear-ReLU stack; reference base.py:8-44 (keys seq_fcs.{0,2,..}.{weight,bias})."""

  def __init__(self, input_shape, hidden_shapes, activation_func=nn.ReLU, init_func=init.basic_init,
               add_ln=False, last_activation_func=None):
    super().__init__()
    if activation_func is not nn.ReLU or add_ln or last_activation_func not in (None, nn.ReLU):
      raise NotImplementedError("the CUDA path implements the shipped configuration: ReLU "
                                "activations, no LayerNorm in the MLPs")
    self.activation_func = activation_func
    self.last_activation_func = activation_func
    self.add_ln = add_ln
    width = int(np.prod(input_shape))
    self.input_dim = width
    self.output_shape = width
    layers = []
    for nxt in hidden_shapes:
      fc = nn.Linear(width, nxt)
      init_func(fc)
      layers += [fc, activation_func()]
      width = nxt
      self.output_shape = nxt
    self.seq_fcs = nn.Sequential(*layers)

  def forward(self, x):
    _container_only("MLPBase")


class Flatten(nn.Module):
  def forward(self, x):
    return x.view(x.size(0), -1)


def _relu_orthogonal(module):
  if isinstance(module, (nn.Linear, nn.Conv2d)):
    nn.init.orthogonal_(module.weight.data, nn.init.calculate_gain("relu"))
    nn.init.constant_(module.bias.data, 0)
  return module


class NatureEncoder(nn.Module):
  """conv 8x8/4 - 4x4/2 - 3x3/1 with ReLUs; reference base.py:304-342 (keys layers.{0,2,4})."""

  def __init__(self, in_channels, groups=1, flatten=True, **kwargs):
    super().__init__()
    if groups != 1:
      raise NotImplementedError("groups != 1 is not used by any shipped config")
    self.groups = groups
    self.in_channels = in_channels
    mods = [nn.Conv2d(in_channels, 32, kernel_size=8, stride=4), nn.ReLU(),
            nn.Conv2d(32, 64, kernel_size=4, stride=2), nn.ReLU(),
            nn.Conv2d(64, 64, kernel_size=3, stride=1), nn.ReLU()]
    if flatten:
      mods.append(Flatten())
    self.layers = nn.Sequential(*mods)
    self.output_dim = 1024
    self.apply(_relu_orthogonal)

  def forward(self, x, detach=False):
    _container_only("NatureEncoder")


def _projection_init(m):
  if isinstance(m, nn.Linear):
    nn.init.orthogonal_(m.weight.data)
    if hasattr(m.bias, "data"):
      m.bias.data.fill_(0.0)


class RLProjection(nn.Module):
  """Linear + ReLU, orthogonal init; reference base.py:209-230 (keys projection.0.*)."""

  def __init__(self, in_dim, out_dim, proj=True):
    super().__init__()
    if not proj:
      raise NotImplementedError("proj=False is not used by any shipped config")
    self.out_dim = out_dim
    self.output_dim = out_dim
    self.projection = nn.Sequential(nn.Linear(in_dim, out_dim), nn.ReLU())
    self.apply(_projection_init)

  def forward(self, x):
    _container_only("RLProjection")


class NatureFuseEncoder(nn.Module):
  """NatureCNN -> 1024->visual_dim projection, proprio MLP; reference base.py:345-385."""

  def __init__(self, in_channels, state_input_dim, visual_dim, hidden_shapes, proj=True, **kwargs):
    super().__init__()
    if in_channels != 4:
      raise NotImplementedError("the CUDA path implements the 4-frame depth stack (in_channels=4)")
    self.in_channels = in_channels
    self.visual_base = NatureEncoder(in_channels)
    self.visual_dim = visual_dim
    self.visual_projector = RLProjection(self.visual_base.output_dim, visual_dim)
    self.base = MLPBase(input_shape=state_input_dim, hidden_shapes=hidden_shapes, **kwargs)

  def forward(self, visual_x, state_x, detach=False):
    _container_only("NatureFuseEncoder")


class _TokenEncoderMixin:
  def _build_visual(self, in_channels, token_dim, two_by_two):
    if in_channels != 4:
      raise NotImplementedError("the CUDA path implements the depth-only token encoder "
                                "(in_channels=4); RGB-D variants are unused by the shipped configs")
    if two_by_two or token_dim != 64:
      raise NotImplementedError("two_by_two / token_dim != 64 are unused by the shipped configs")
    self.in_channels = in_channels
    self.token_dim = token_dim
    self.two_by_two = two_by_two
    self.depth_visual_base = NatureEncoder(4, flatten=False)
    self.depth_up_conv = nn.Conv2d(64, token_dim, 1)


class TransformerEncoder(nn.Module, _TokenEncoderMixin):
  """Vision-only token encoder (16 depth tokens); reference base.py:388-494."""

  def __init__(self, in_channels, token_dim=64, two_by_two=False, **kwargs):
    super().__init__()
    self._build_visual(in_channels, token_dim, two_by_two)
    self.visual_dim = token_dim
    self.per_modal_tokens = 16
    self.flatten_layer = Flatten()

  def forward(self, visual_x, detach=False, return_raw_visual_vecs=False):
    _container_only("TransformerEncoder")


class LocoTransformerEncoder(nn.Module, _TokenEncoderMixin):
  """16 depth tokens + 1 proprio token; reference base.py:497-626."""

  def __init__(self, in_channels, state_input_dim, hidden_shapes, token_dim=64, two_by_two=False,
               visual_dim=None, proj=True, **kwargs):
    super().__init__()
    self._build_visual(in_channels, token_dim, two_by_two)
    self.base = MLPBase(input_shape=state_input_dim, hidden_shapes=hidden_shapes, **kwargs)
    self.state_projector = RLProjection(self.base.output_shape, token_dim)
    self.visual_dim = token_dim        # the JSON's visual_dim is ignored here
    self.per_modal_tokens = 16
    self.flatten_layer = Flatten()

  def forward(self, visual_x, state_x, detach=False, return_raw_visual_vecs=False):
    _container_only("LocoTransformerEncoder")
