"""Policy / value networks whose forward (and backward) is the CUDA layer plan.

Same constructor signatures, attributes and state_dict keys as the reference classes
(torchrl/networks/nets.py:16-55, 133-262, 784-1038); unknown kwargs are swallowed exactly like
the reference's **kwargs.  `module(x)` accepts what the reference accepts — x [..., D] on the
module's CUDA device, any leading batch shape including none — and is
differentiable through torch.autograd (the backward pass calls the CUDA data/weight-gradient
kernels).  There is no CPU execution path: inputs or parameters on the CPU raise.
"""
import numpy as np
import torch
import torch.nn as nn

from . import init
from . import base
from .. import engine
from .._lib import V4LError


class _PlanFunction(torch.autograd.Function):
  """x -> plan.forward; grads of every parameter <- plan.backward."""

  @staticmethod
  def forward(ctx, net, plan, names, x, *params):
    P = dict(zip(names, params))
    out = torch.empty((x.shape[0], plan.out_dim), device=x.device, dtype=torch.float32)
    plan.forward(P, engine.Input.from_flat(x, net._state_dim, net._has_img), out)
    plan.version += 1
    ctx.plan, ctx.P, ctx.names, ctx.version = plan, P, names, plan.version
    ctx.shapes = [p.shape for p in params]
    return out

  @staticmethod
  def backward(ctx, d_out):
    plan = ctx.plan
    if plan.version != ctx.version:
      raise V4LError("backward() after another forward() of the same network at the same batch "
                     "size: the saved activations were overwritten")
    G = {n: torch.empty(s, device=d_out.device, dtype=torch.float32)
         for n, s in zip(ctx.names, ctx.shapes)}
    plan.backward(ctx.P, G, d_out.contiguous().float())
    return (None, None, None, None) + tuple(G[n] for n in ctx.names)


class _CudaNet:
  """Mixin: forward through an engine plan.  Subclasses set _family, _state_dim, _has_img."""

  _plan_kwargs = {}

  def _named_params(self):
    names, params = [], []
    for n, p in self.named_parameters():
      if n == "logstd":
        continue
      names.append(n)
      params.append(p)
    return names, params

  def _get_plan(self, device, B):
    plans = self.__dict__.setdefault("_plans", {})
    key = (device.index, B)
    plan = plans.get(key)
    if plan is None:
      if len(plans) >= 4:                 # collector/eval/update use a handful of batch sizes
        plans.pop(next(iter(plans))).release()
      plan = engine.make_plan(self._family, engine.ops_for(device), self._state_dim, self._out_dim,
                              **self._plan_kwargs)
      plan.version = 0
      plans[key] = plan
    return plan

  def _net_forward(self, x):
    if not isinstance(x, torch.Tensor):
      x = torch.as_tensor(x)
    if x.device.type != "cuda":
      raise V4LError("vision4leg_b200 networks run on CUDA only (input is on %s); there is no "
                     "CPU fallback" % x.device)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1]).contiguous().float()
    expect = self._state_dim + (engine.IMG_ELEMS if self._has_img else 0)
    if x2.shape[1] != expect:
      raise V4LError("observation width %d, expected %d" % (x2.shape[1], expect))
    names, params = self._named_params()
    if params[0].device != x2.device:
      raise V4LError("parameters on %s but input on %s" % (params[0].device, x2.device))
    plan = self._get_plan(x2.device, x2.shape[0])
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
      out = _PlanFunction.apply(self, plan, names, x2, *params)
    else:
      out = torch.empty((x2.shape[0], self._out_dim), device=x2.device, dtype=torch.float32)
      plan.forward(dict(zip(names, params)), engine.Input.from_flat(x2, self._state_dim, self._has_img), out)
      plan.version += 1
    return out.reshape(lead + (self._out_dim,))

  def __deepcopy__(self, memo):
    # plans hold device workspaces and ctypes handles: never copied (PPO deep-copies pf,
    # reference ppo.py:21)
    cls = self.__class__
    new = cls.__new__(cls)
    memo[id(self)] = new
    import copy
    for k, v in self.__dict__.items():
      if k == "_plans":
        continue
      new.__dict__[k] = copy.deepcopy(v, memo)
    return new


def _append_fcs(in_dim, hidden_shapes, out_dim, hidden_init, last_init, activation_func, add_ln):
  if activation_func is not nn.ReLU or add_ln:
    raise NotImplementedError("the CUDA path implements ReLU heads without LayerNorm "
                              "(every shipped config)")
  mods = []
  for nxt in hidden_shapes:
    fc = nn.Linear(in_dim, nxt)
    hidden_init(fc)
    mods += [fc, activation_func()]
    in_dim = nxt
  last = nn.Linear(in_dim, out_dim)
  last_init(last)
  mods.append(last)
  return nn.Sequential(*mods)


class ZeroNet(nn.Module):
  def forward(self, x):
    return torch.zeros(1)


class Net(_CudaNet, nn.Module):
  """base_type(**kwargs) + append MLP; reference nets.py:16-55 (starter/ppo_state.py)."""
  _family, _has_img = "mlp", False

  def __init__(self, output_shape, base_type, append_hidden_shapes=[],
               append_hidden_init_func=init.basic_init, net_last_init_func=init.uniform_init,
               activation_func=nn.ReLU, add_ln=False, **kwargs):
    super().__init__()
    self.base = base_type(activation_func=activation_func, add_ln=add_ln, **kwargs)
    if not isinstance(self.base, base.MLPBase):
      raise NotImplementedError("Net is accelerated for base_type=MLPBase (starter/ppo_state.py)")
    self.add_ln = add_ln
    self.activation_func = activation_func
    self.seq_append_fcs = _append_fcs(self.base.output_shape, append_hidden_shapes, output_shape,
                                      append_hidden_init_func, net_last_init_func, activation_func, add_ln)
    self._out_dim = output_shape

  @property
  def _state_dim(self):
    return self.base.input_dim

  def forward(self, x):
    return self._net_forward(x)


class ImpalaEncoderProjNet(_CudaNet, nn.Module):
  """NatureFuseEncoder -> cat(visual, state) -> MLP head; reference nets.py:194-262
  (the class name is historical; starter/ppo_nature_cnn.py passes a NatureFuseEncoder)."""
  _family, _has_img = "nature", True

  def __init__(self, encoder, output_shape, state_input_shape, visual_input_shape,
               append_hidden_shapes=[], append_hidden_init_func=init.basic_init,
               net_last_init_func=init.uniform_init, activation_func=nn.ReLU, add_ln=False,
               detach=False, **kwargs):
    super().__init__()
    if not isinstance(encoder, base.NatureFuseEncoder):
      raise NotImplementedError("ImpalaEncoderProjNet is accelerated for NatureFuseEncoder")
    if detach:
      raise NotImplementedError("detach=True is unused by the shipped configs")
    self.encoder = encoder
    self.add_ln, self.detach = add_ln, detach
    self.state_input_shape = state_input_shape
    self.visual_input_shape = visual_input_shape
    self.activation_func = activation_func
    self.seq_append_fcs = _append_fcs(encoder.base.output_shape + encoder.visual_dim,
                                      append_hidden_shapes, output_shape, append_hidden_init_func,
                                      net_last_init_func, activation_func, add_ln)
    self.normalizer = None
    self._out_dim = output_shape
    self._state_dim = int(state_input_shape)
    _check_visual(visual_input_shape)

  def forward(self, state):
    return self._net_forward(state)


class NatureEncoderProjNet(_CudaNet, nn.Module):
  """Vision-only: flattening NatureEncoder -> MLP head; reference nets.py:133-191
  (starter/ppo_nature_cnn_vision_only.py)."""
  _family, _has_img, _state_dim = "nvo", True, 0

  def __init__(self, encoder, output_shape, visual_input_shape, append_hidden_shapes=[],
               append_hidden_init_func=init.basic_init, net_last_init_func=init.uniform_init,
               activation_func=nn.ReLU, add_ln=False, detach=False, **kwargs):
    super().__init__()
    if not isinstance(encoder, base.NatureEncoder) or not any(isinstance(m, base.Flatten) for m in encoder.layers):
      raise NotImplementedError("NatureEncoderProjNet needs a flattening NatureEncoder")
    if detach:
      raise NotImplementedError("detach=True is unused by the shipped configs")
    self.encoder = encoder
    self.add_ln, self.detach = add_ln, detach
    self.visual_input_shape = visual_input_shape
    self.activation_func = activation_func
    self.seq_append_fcs = _append_fcs(encoder.output_dim, append_hidden_shapes, output_shape,
                                      append_hidden_init_func, net_last_init_func, activation_func, add_ln)
    self.normalizer = None
    self._out_dim = output_shape
    _check_visual(visual_input_shape)

  def forward(self, x):
    return self._net_forward(x)


def _check_visual(shape):
  if tuple(shape) != (4, 64, 64):
    raise NotImplementedError("visual_input_shape must be (4, 64, 64), got %r" % (tuple(shape),))


class _TransformerNet(_CudaNet, nn.Module):
  """Shared body of LocoTransformer / Transformer (reference nets.py:784-1038)."""

  def _build(self, encoder, output_shape, visual_input_shape, transformer_params,
             append_hidden_shapes, append_hidden_init_func, net_last_init_func, activation_func,
             add_ln, detach, state_detach, max_pool, token_norm, use_pytorch_encoder, head_in):
    if detach or state_detach or max_pool or token_norm or use_pytorch_encoder:
      raise NotImplementedError("detach/state_detach/max_pool/token_norm/use_pytorch_encoder are "
                                "unused by the shipped configs and not implemented on the CUDA path")
    self.encoder = encoder
    self.add_ln, self.detach, self.state_detach = add_ln, detach, state_detach
    self.visual_input_shape = visual_input_shape
    self.activation_func = activation_func
    self.max_pool, self.token_norm = max_pool, token_norm
    self.use_pytorch_encoder = use_pytorch_encoder
    d = encoder.visual_dim
    self.visual_append_layers = nn.ModuleList()
    for n_head, dim_feedforward in transformer_params:
      # parameter container: nn.TransformerEncoderLayer(d, n_head, ff, dropout=0) is what the
      # reference instantiates (nets.py:949-955); its forward is never called here
      self.visual_append_layers.append(nn.TransformerEncoderLayer(d, n_head, dim_feedforward, dropout=0))
    self._plan_kwargs = {"n_heads": tuple(int(h) for h, _ in transformer_params)}
    self.per_modal_tokens = encoder.per_modal_tokens
    self.second = False
    self.visual_seq_append_fcs = _append_fcs(head_in, append_hidden_shapes, output_shape,
                                             append_hidden_init_func, net_last_init_func,
                                             activation_func, add_ln)
    self.normalizer = None
    self._out_dim = output_shape
    _check_visual(visual_input_shape)

  def forward(self, x):
    return self._net_forward(x)


class LocoTransformer(_TransformerNet):
  """Reference nets.py:909-1038."""
  _family, _has_img = "loco", True

  def __init__(self, encoder, output_shape, state_input_shape, visual_input_shape,
               transformer_params=[], append_hidden_shapes=[],
               append_hidden_init_func=init.basic_init, net_last_init_func=init.uniform_init,
               activation_func=nn.ReLU, add_ln=False, detach=False, state_detach=False,
               max_pool=False, token_norm=False, use_pytorch_encoder=False, **kwargs):
    super().__init__()
    if not isinstance(encoder, base.LocoTransformerEncoder):
      raise NotImplementedError("LocoTransformer needs a LocoTransformerEncoder")
    self.state_input_shape = state_input_shape
    self._state_dim = int(state_input_shape)
    self._build(encoder, output_shape, visual_input_shape, transformer_params, append_hidden_shapes,
                append_hidden_init_func, net_last_init_func, activation_func, add_ln, detach,
                state_detach, max_pool, token_norm, use_pytorch_encoder, encoder.visual_dim * 2)


class Transformer(_TransformerNet):
  """Vision-only variant; reference nets.py:784-906."""
  _family, _has_img, _state_dim = "vit", True, 0

  def __init__(self, encoder, output_shape, visual_input_shape, transformer_params=[],
               append_hidden_shapes=[], append_hidden_init_func=init.basic_init,
               net_last_init_func=init.uniform_init, activation_func=nn.ReLU, add_ln=False,
               detach=False, state_detach=False, max_pool=False, token_norm=False,
               use_pytorch_encoder=False, **kwargs):
    super().__init__()
    if not isinstance(encoder, base.TransformerEncoder):
      raise NotImplementedError("Transformer needs a TransformerEncoder")
    self._build(encoder, output_shape, visual_input_shape, transformer_params, append_hidden_shapes,
                append_hidden_init_func, net_last_init_func, activation_func, add_ln, detach,
                state_detach, max_pool, token_norm, use_pytorch_encoder, encoder.visual_dim)
