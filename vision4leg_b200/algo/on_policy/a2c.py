"""Advantage actor-critic (reference torchrl/algo/on_policy/a2c.py:8-114).

Base class of PPO (optimiser construction, snapshot list).  Its own `update` is not on the
hot path (README.md:88 — only PPO is used): it differentiates through the CUDA networks with
torch.autograd and steps with torch.optim, i.e. the network forward/backward kernels are ours,
the scalar loss glue is torch.
"""
import torch
import torch.nn as nn
import torch.optim as optim

from .on_rl_algo import OnRLAlgo


class A2C(OnRLAlgo):
  def __init__(self, pf, vf, plr=3e-4, vlr=3e-4, optimizer_class=optim.Adam, entropy_coeff=0.001,
               **kwargs):
    super().__init__(**kwargs)
    self.pf = pf
    self.vf = vf
    self.to(self.device)
    self.plr = plr
    self.vlr = vlr
    self.optimizer_class = optimizer_class
    self.pf_optimizer = optimizer_class(self.pf.parameters(), lr=self.plr, eps=1e-5)
    self.vf_optimizer = optimizer_class(self.vf.parameters(), lr=self.vlr, eps=1e-5)
    self.entropy_coeff = entropy_coeff
    self.vf_criterion = nn.MSELoss()

  def _to_device(self, batch, keys):
    return [torch.as_tensor(batch[k], dtype=torch.float32).to(self.device) for k in keys]

  def update(self, batch):
    self.training_update_num += 1
    obs, acts, advs, est_rets = self._to_device(batch, ["obs", "acts", "advs", "estimate_returns"])
    out = self.pf.update(obs, acts)
    log_probs, ent = out["log_prob"], out["ent"]
    advs = (advs - advs.mean()) / (advs.std() + 1e-5)
    assert log_probs.shape == advs.shape, \
      "log_prob shape: {}, adv shape: {}".format(log_probs.shape, advs.shape)
    policy_loss = (-log_probs * advs).mean() - self.entropy_coeff * ent.mean()
    values = self.vf(obs)
    vf_loss = self.vf_criterion(values, est_rets)
    # Both backward passes run BEFORE either optimiser step: with a shared encoder the actor's step
    # changes weights the critic's saved activations were computed with (the reference's order,
    # a2c.py:66-76, makes torch raise an in-place-modification error in that case; with separate
    # networks the two orders give identical results).  Gradients of shared parameters are kept apart.
    shared = {id(p) for p in self.pf.parameters()} & {id(p) for p in self.vf.parameters()}
    self.pf_optimizer.zero_grad()
    policy_loss.backward()
    pf_grads = {id(p): p.grad.clone() for p in self.pf.parameters() if id(p) in shared and p.grad is not None}
    self.vf_optimizer.zero_grad()
    vf_loss.backward()
    vf_grads = {id(p): p.grad.clone() for p in self.vf.parameters() if id(p) in shared and p.grad is not None}
    for p in self.pf.parameters():
      if id(p) in pf_grads:
        p.grad.copy_(pf_grads[id(p)])
    torch.nn.utils.clip_grad_norm_(self.pf.parameters(), 0.5)
    self.pf_optimizer.step()
    for p in self.vf.parameters():
      if id(p) in vf_grads:
        p.grad.copy_(vf_grads[id(p)])
    torch.nn.utils.clip_grad_norm_(self.vf.parameters(), 0.5)
    self.vf_optimizer.step()

    info = {"Training/policy_loss": policy_loss.item(), "Training/vf_loss": vf_loss.item()}
    for tag, t in (("v_pred", values),) + ((("std", out["std"]),) if "std" in out else ()):
      info[tag + "/mean"] = t.mean().item()
      info[tag + "/std"] = t.std().item()
      info[tag + "/max"] = t.max().item()
      info[tag + "/min"] = t.min().item()
    info["ent"] = ent.mean().item()
    info["log_prob"] = log_probs.mean().item()
    return info

  @property
  def snapshot_networks(self):
    return [("pf", self.pf), ("vf", self.vf)]
