"""State-independent-std Gaussian policies over the CUDA networks.

Same surface as the reference (torchrl/policies/continuous_policy.py:77-146, 239-290, 461-492):
`forward(x) -> (mean, std, log_std)`, `update(obs, acts) -> dict`, `explore(x, ...) -> dict`,
`eval_act(x) -> np.ndarray`; parameter `logstd` initialised to log(log_init) and clamped to
[LOG_SIG_MIN, LOG_SIG_MAX].  `mean` comes from the CUDA layer plan; the few [B, A] elementwise
ops of explore/update run as torch ops on the device (collector side — the PPO update itself
uses the fused loss kernel, see algo/ppo.py).
"""
import numpy as np
import torch
import torch.nn as nn
from torch.distributions import Normal

from .. import networks
from .distribution import TanhNormal

LOG_SIG_MAX = 2
LOG_SIG_MIN = -5


class GaussianContPolicyBase:
  def _gaussian_head(self, mean):
    logstd = torch.clamp(self.logstd, LOG_SIG_MIN, LOG_SIG_MAX)
    std = torch.exp(logstd).unsqueeze(0).expand_as(mean)
    return mean, std, logstd

  def _dist(self, mean, std):
    return TanhNormal(mean, std) if self.tanh_action else Normal(mean, std)

  def eval_act(self, x):
    with torch.no_grad():
      mean, _, _ = self.forward(x)
    if self.tanh_action:
      mean = torch.tanh(mean)
    return mean.squeeze(0).detach().cpu().numpy()

  def explore(self, x, return_log_probs=False, return_pre_tanh=False):
    mean, std, log_std = self.forward(x)
    dis = self._dist(mean, std)
    out = {"mean": mean, "log_std": log_std, "std": std, "ent": dis.entropy().sum(-1, keepdim=True)}
    if self.tanh_action:
      if return_log_probs:
        action, z = dis.rsample(return_pretanh_value=True)
        out["log_prob"] = dis.log_prob(action, pre_tanh_value=z).sum(dim=-1, keepdim=True)
        out["pre_tanh"] = z.squeeze(0)
      else:
        if return_pre_tanh:
          _, z = dis.rsample(return_pretanh_value=True)
          out["pre_tanh"] = z.squeeze(0)
        action = dis.rsample(return_pretanh_value=False)
    else:
      action = dis.sample()
      if return_log_probs:
        out["log_prob"] = dis.log_prob(action).sum(dim=-1, keepdim=True)
    out["action"] = action.squeeze(0)
    return out

  def update(self, obs, actions):
    mean, std, log_std = self.forward(obs)
    dis = self._dist(mean, std)
    return {"mean": mean, "dis": Normal(mean, std), "log_std": log_std, "std": std,
            "log_prob": dis.log_prob(actions).sum(-1, keepdim=True),
            "ent": dis.entropy().sum(-1, keepdim=True)}


def _gaussian_policy(net_cls, doc):
  class _Policy(net_cls, GaussianContPolicyBase):
    def __init__(self, output_shape, tanh_action=False, log_init=0.125, **kwargs):
      super().__init__(output_shape=output_shape, **kwargs)
      self.continuous = True
      self.logstd = nn.Parameter(torch.ones(output_shape) * np.log(log_init))
      self.tanh_action = tanh_action

    def forward(self, x):
      return self._gaussian_head(net_cls.forward(self, x))
  _Policy.__doc__ = doc
  return _Policy


GaussianContPolicyBasicBias = _gaussian_policy(
  networks.Net, "Gaussian policy over Net (reference continuous_policy.py:239-254)")
GaussianContPolicyImpalaEncoderProj = _gaussian_policy(
  networks.ImpalaEncoderProjNet, "reference continuous_policy.py:275-290")
GaussianContPolicyNatureEncoderProj = _gaussian_policy(
  networks.NatureEncoderProjNet, "reference continuous_policy.py:257-272")
GaussianContPolicyTransformer = _gaussian_policy(
  networks.Transformer, "reference continuous_policy.py:461-475")
GaussianContPolicyLocoTransformer = _gaussian_policy(
  networks.LocoTransformer, "reference continuous_policy.py:478-492")
for _n in ("GaussianContPolicyBasicBias", "GaussianContPolicyImpalaEncoderProj",
           "GaussianContPolicyNatureEncoderProj",
           "GaussianContPolicyTransformer", "GaussianContPolicyLocoTransformer"):
  globals()[_n].__name__ = globals()[_n].__qualname__ = _n
