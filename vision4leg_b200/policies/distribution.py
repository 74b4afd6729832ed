"""tanh-squashed Gaussian used when a policy is built with tanh_action=True (reference
torchrl/policies/distribution.py:5-80; every shipped config has "policy": {} -> off).
Collector-side only: the PPO CUDA path covers the unsquashed Gaussian of the shipped configs."""
import torch
from torch.distributions import Distribution, Normal


class TanhNormal(Distribution):
  arg_constraints = {}

  def __init__(self, normal_mean, normal_std, epsilon=1e-6):
    super().__init__(validate_args=False)
    self.normal_mean, self.normal_std = normal_mean, normal_std
    self.normal = Normal(normal_mean, normal_std)
    self.epsilon = epsilon

  def _squash(self, z, with_z):
    y = torch.tanh(z)
    return (y, z) if with_z else y

  def log_prob(self, value, pre_tanh_value=None):
    z = pre_tanh_value
    if z is None:
      z = 0.5 * torch.log((1 + value) / (1 - value))      # atanh
    return self.normal.log_prob(z) - torch.log(1 - value * value + self.epsilon)

  def sample(self, return_pretanh_value=False):
    return self._squash(self.normal.sample().detach(), return_pretanh_value)

  def rsample(self, return_pretanh_value=False):
    noise = torch.randn(self.normal_mean.size(), device=self.normal_mean.device)
    return self._squash(self.normal_mean + self.normal_std * noise, return_pretanh_value)

  def entropy(self):
    return self.normal.entropy()
