"""CPU oracle for the PPO-update hot path (TEST INFRASTRUCTURE — never imported by the product).

A functional restatement, in plain torch-CPU / numpy, of the reference's PPO update path.
It follows the reference file:line cited at each function; it shares no code with it (the
reference is a set of nn.Module classes, this is a flat functional form over state_dicts).

Parity status: the reference ships NO tests or golden vectors for this path,
so the oracle is pinned against outputs of the *live reference modules* imported in the build
container by oracle/make_golden.py (tests/golden/*.npz, checked in tests/test_oracle_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
import this module.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LOG_SIG_MAX = 2.0   # reference torchrl/policies/continuous_policy.py:8-9
LOG_SIG_MIN = -5.0


# =============================================================================================
# GAE / discounted return  (reference torchrl/replay_buffers/on_policy.py:17-71)
# =============================================================================================

def gae(rewards, values, terminals, time_limits, last_value, gamma, tau, time_limit_filter=True):
  """float64 reverse recurrence, vectorised over the env axis.
  rewards/values/terminals [T,E,1]; time_limits [T,1] or [T,E,1]; last_value [E,1].
  Returns (advs, estimate_returns) float64 [T,E,1]."""
  r = np.asarray(rewards, np.float64)
  v = np.asarray(values, np.float64)
  d = np.asarray(terminals, np.float64)
  tl = np.asarray(time_limits, np.float64)
  T = r.shape[0]
  advs = np.zeros_like(r)
  rets = np.zeros_like(r)
  v_next = np.asarray(last_value, np.float64)
  A = np.zeros_like(v_next)
  for t in range(T - 1, -1, -1):
    nt = 1.0 - d[t]
    delta = r[t] + nt * gamma * v_next - v[t]
    A = delta + nt * gamma * tau * A
    if time_limit_filter:
      A = A * (1.0 - tl[t])
    advs[t] = A
    rets[t] = A + v[t]
    v_next = v[t]
  return advs, rets


def discount_reward(rewards, values, terminals, time_limits, last_value, gamma,
                    time_limit_filter=True):
  """reference on_policy.py:47-71 (used when gae=False)."""
  r = np.asarray(rewards, np.float64)
  v = np.asarray(values, np.float64)
  d = np.asarray(terminals, np.float64)
  tl = np.asarray(time_limits, np.float64)
  T = r.shape[0]
  advs = np.zeros_like(r)
  rets = np.zeros_like(r)
  R = np.asarray(last_value, np.float64)
  for t in range(T - 1, -1, -1):
    if time_limit_filter:
      R = (r[t] + (1.0 - d[t]) * gamma * R * (1.0 - tl[t])) + tl[t] * v[t]
    else:
      R = r[t] + (1.0 - d[t]) * gamma * R
    advs[t] = R - v[t]
    rets[t] = R
  return advs, rets


# =============================================================================================
# Networks (functional; P is a dict name -> torch tensor using the reference state_dict keys)
# =============================================================================================

def nature_cnn(P, prefix, img):
  """NatureEncoder, flatten=False (reference torchrl/networks/base.py:304-342).
  img [B,4,64,64] -> [B,64,4,4]"""
  x = F.relu(F.conv2d(img, P[prefix + "layers.0.weight"], P[prefix + "layers.0.bias"], stride=4))
  x = F.relu(F.conv2d(x, P[prefix + "layers.2.weight"], P[prefix + "layers.2.bias"], stride=2))
  x = F.relu(F.conv2d(x, P[prefix + "layers.4.weight"], P[prefix + "layers.4.bias"], stride=1))
  return x


def mlp_base(P, prefix, x):
  """MLPBase: Linear-ReLU stack, last activation ReLU as well (base.py:8-44)."""
  i = 0
  while (prefix + "seq_fcs.%d.weight" % i) in P:
    x = F.relu(F.linear(x, P[prefix + "seq_fcs.%d.weight" % i], P[prefix + "seq_fcs.%d.bias" % i]))
    i += 2
  return x


def head(P, prefix, x):
  """append_fcs: Linear-ReLU ... Linear (nets.py:36-49, 973-992)."""
  idx = sorted(int(k[len(prefix):].split(".")[0]) for k in P
               if k.startswith(prefix) and k.endswith(".weight"))
  for j, i in enumerate(idx):
    x = F.linear(x, P[prefix + "%d.weight" % i], P[prefix + "%d.bias" % i])
    if j + 1 < len(idx):
      x = F.relu(x)
  return x


def layer_norm(x, w, b, eps=1e-5):
  mu = x.mean(-1, keepdim=True)
  var = ((x - mu) ** 2).mean(-1, keepdim=True)
  return (x - mu) / torch.sqrt(var + eps) * w + b


def transformer_layer(P, prefix, x, n_head=1):
  """nn.TransformerEncoderLayer(d, n_head, ff, dropout=0): post-norm, ReLU, eps 1e-5
  (instantiated at reference nets.py:949-955). x [B,T,d] batch-first here
  (the reference is seq-first; the math is per sample so only the layout differs)."""
  B, T, d = x.shape
  hd = d // n_head
  qkv = F.linear(x, P[prefix + "self_attn.in_proj_weight"], P[prefix + "self_attn.in_proj_bias"])
  q, k, v = qkv.split(d, dim=-1)
  q = q.reshape(B, T, n_head, hd).transpose(1, 2)
  k = k.reshape(B, T, n_head, hd).transpose(1, 2)
  v = v.reshape(B, T, n_head, hd).transpose(1, 2)
  s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(hd)
  p = torch.softmax(s, dim=-1)
  o = torch.matmul(p, v).transpose(1, 2).reshape(B, T, d)
  o = F.linear(o, P[prefix + "self_attn.out_proj.weight"], P[prefix + "self_attn.out_proj.bias"])
  h = layer_norm(x + o, P[prefix + "norm1.weight"], P[prefix + "norm1.bias"])
  f = F.linear(F.relu(F.linear(h, P[prefix + "linear1.weight"], P[prefix + "linear1.bias"])),
               P[prefix + "linear2.weight"], P[prefix + "linear2.bias"])
  return layer_norm(h + f, P[prefix + "norm2.weight"], P[prefix + "norm2.bias"])


def loco_forward(P, x, S, n_head=1, return_tokens=False):
  """LocoTransformer.forward + LocoTransformerEncoder.forward
  (reference nets.py:996-1038, base.py:550-626). x [B, S+16384] -> [B, out]."""
  B = x.shape[0]
  state = x[:, :S]
  img = x[:, S:].reshape(B, 4, 64, 64)
  feat = nature_cnn(P, "encoder.depth_visual_base.", img)
  up = F.conv2d(feat, P["encoder.depth_up_conv.weight"], P["encoder.depth_up_conv.bias"])
  vis_tok = up.reshape(B, up.shape[1], 16).permute(0, 2, 1)          # [B,16,64], token = h*4+w
  s = mlp_base(P, "encoder.base.", state)
  s_tok = F.relu(F.linear(s, P["encoder.state_projector.projection.0.weight"],
                          P["encoder.state_projector.projection.0.bias"]))
  tok = torch.cat([s_tok[:, None, :], vis_tok], dim=1)               # [B,17,64]
  l = 0
  while ("visual_append_layers.%d.linear1.weight" % l) in P:
    tok = transformer_layer(P, "visual_append_layers.%d." % l, tok, n_head)
    l += 1
  pooled = torch.cat([tok[:, 0], tok[:, 1:17].mean(1)], dim=-1)      # [B,128]
  out = head(P, "visual_seq_append_fcs.", pooled)
  if return_tokens:
    return out, tok
  return out


def nature_forward(P, x, S):
  """ImpalaEncoderProjNet.forward + NatureFuseEncoder.forward
  (reference nets.py:247-262, base.py:371-385)."""
  B = x.shape[0]
  state = x[:, :S]
  img = x[:, S:].reshape(B, 4, 64, 64)
  feat = nature_cnn(P, "encoder.visual_base.", img).reshape(B, 1024)  # (c,h,w) flatten
  vis = F.relu(F.linear(feat, P["encoder.visual_projector.projection.0.weight"],
                        P["encoder.visual_projector.projection.0.bias"]))
  s = mlp_base(P, "encoder.base.", state)
  return head(P, "seq_append_fcs.", torch.cat([vis, s], dim=-1))


def vit_forward(P, x, S=0, n_head=1):
  """Transformer.forward + TransformerEncoder.forward, vision only (reference nets.py:868-906,
  base.py:428-494): 16 depth tokens, mean-pooled."""
  B = x.shape[0]
  img = x.reshape(B, 4, 64, 64)
  feat = nature_cnn(P, "encoder.depth_visual_base.", img)
  up = F.conv2d(feat, P["encoder.depth_up_conv.weight"], P["encoder.depth_up_conv.bias"])
  tok = up.reshape(B, up.shape[1], 16).permute(0, 2, 1)
  l = 0
  while ("visual_append_layers.%d.linear1.weight" % l) in P:
    tok = transformer_layer(P, "visual_append_layers.%d." % l, tok, n_head)
    l += 1
  return head(P, "visual_seq_append_fcs.", tok.mean(1))


def nvo_forward(P, x, S=0):
  """NatureEncoderProjNet.forward with a flattening NatureEncoder (reference nets.py:177-191,
  base.py:334-342)."""
  B = x.shape[0]
  feat = nature_cnn(P, "encoder.", x.reshape(B, 4, 64, 64)).reshape(B, 1024)
  return head(P, "seq_append_fcs.", feat)


def mlp_forward(P, x, S=None):
  """Net.forward with base_type=MLPBase (reference nets.py:51-55)."""
  return head(P, "seq_append_fcs.", mlp_base(P, "base.", x))


FORWARD = {"loco": loco_forward, "nature": nature_forward, "mlp": mlp_forward, "vit": vit_forward,
           "nvo": nvo_forward}


def gaussian_update(mean, logstd_param, acts):
  """GaussianContPolicyBase.update with tanh_action=False
  (reference continuous_policy.py:127-146, forward :486-492; torch.distributions.Normal).
  Returns log_prob [B,1], ent [B,1], clamped logstd [A]."""
  logstd = torch.clamp(logstd_param, LOG_SIG_MIN, LOG_SIG_MAX)
  std = torch.exp(logstd)
  var = std * std
  lp = -((acts - mean) ** 2) / (2 * var) - torch.log(std) - math.log(math.sqrt(2 * math.pi))
  ent = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(std)).expand_as(mean)
  return lp.sum(-1, keepdim=True), ent.sum(-1, keepdim=True), logstd


# =============================================================================================
# Optimiser pieces
# =============================================================================================

def clip_grad_norm(grads, max_norm=0.5):
  """torch.nn.utils.clip_grad_norm_ semantics (used at reference ppo.py:73-74,118-119)."""
  total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
  coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
  for g in grads:
    g.mul_(coef)
  return total


class Adam:
  """torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-5, no weight decay) restated
  (instantiated at reference a2c.py:30-40)."""

  def __init__(self, params, lr, eps=1e-5, betas=(0.9, 0.999)):
    self.params = params          # list of tensors (shared tensors allowed across optimisers)
    self.lr = lr
    self.eps = eps
    self.b1, self.b2 = betas
    self.t = 0
    self.m = [torch.zeros_like(p) for p in params]
    self.v = [torch.zeros_like(p) for p in params]

  def step(self, grads):
    self.t += 1
    bc1 = 1 - self.b1 ** self.t
    bc2 = 1 - self.b2 ** self.t
    for p, g, m, v in zip(self.params, grads, self.m, self.v):
      m.mul_(self.b1).add_(g, alpha=1 - self.b1)
      v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
      denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
      p.addcdiv_(m, denom, value=-self.lr / bc1)


# =============================================================================================
# PPO
# =============================================================================================

class PPOOracle:
  """Restates PPO.update / update_per_epoch (reference torchrl/algo/on_policy/ppo.py:28-153,
  a2c.py:13-44, on_rl_algo.py:23-34, algo/utils.py:23-32).

  pf_sd / vf_sd: dict name -> torch tensor; tensors with the same key prefix `shared_prefix`
  are THE SAME tensor objects in both (shared encoder stepped by both optimisers).
  """

  def __init__(self, family, pf_sd, vf_sd, S, plr=1e-4, vlr=1e-4, clip_para=0.2, opt_epochs=3,
               entropy_coeff=0.005, tau=0.95, discount=0.99, num_epochs=1500, batch_size=1024,
               clipped_value_loss=False, gae=True, n_head=1, dtype=torch.float32):
    self.family = family
    self.S = S
    self.fwd = FORWARD[family]
    self.n_head = n_head
    self.pf = pf_sd
    self.vf = vf_sd
    self.target_pf = {k: v.clone() for k, v in pf_sd.items()}   # ppo.py:21 deepcopy
    self.plr, self.vlr = plr, vlr
    self.clip_para = clip_para
    self.opt_epochs = opt_epochs
    self.entropy_coeff = entropy_coeff
    self.tau, self.discount = tau, discount
    self.num_epochs, self.batch_size = num_epochs, batch_size
    self.clipped_value_loss = clipped_value_loss
    self.use_gae = gae
    self.dtype = dtype
    self.pf_keys = list(pf_sd.keys())
    self.vf_keys = list(vf_sd.keys())
    self.pf_opt = Adam([pf_sd[k] for k in self.pf_keys], plr)
    self.vf_opt = Adam([vf_sd[k] for k in self.vf_keys], vlr)
    self.current_epoch = 0

  def _forward(self, P, x):
    if self.family in ("loco", "vit"):
      return self.fwd(P, x, self.S, self.n_head)
    return self.fwd(P, x, self.S)

  def values(self, obs):
    with torch.no_grad():
      return self._forward(self.vf, obs)

  def policy(self, obs):
    with torch.no_grad():
      return self._forward(self.pf, obs)

  def update(self, batch):
    """One minibatch: reference ppo.py:125-153 (+ :94-123 critic, :42-92 actor)."""
    info = {}
    obs = torch.as_tensor(batch["obs"], dtype=self.dtype)
    acts = torch.as_tensor(batch["acts"], dtype=self.dtype)
    advs = torch.as_tensor(batch["advs"], dtype=self.dtype)
    old_values = torch.as_tensor(batch["values"], dtype=self.dtype)
    est_rets = torch.as_tensor(batch["estimate_returns"], dtype=self.dtype)

    info["advs/mean"] = advs.mean().item()
    info["advs/std"] = advs.std().item()
    info["advs/max"] = advs.max().item()
    info["advs/min"] = advs.min().item()
    advs = (advs - advs.mean()) / (advs.std() + 1e-5)

    # ---- critic (ppo.py:94-123)
    vparams = [self.vf[k] for k in self.vf_keys]
    for p in vparams:
      p.requires_grad_(True)
    values = self._forward(self.vf, obs)
    if self.clipped_value_loss:
      vc = old_values + (values - old_values).clamp(-self.clip_para, self.clip_para)
      vf_loss = 0.5 * torch.max((values - est_rets).pow(2), (vc - est_rets).pow(2)).mean()
    else:
      vf_loss = ((values - est_rets) ** 2).mean()
    vgrads = list(torch.autograd.grad(vf_loss, vparams))
    for p in vparams:
      p.requires_grad_(False)
    vnorm = clip_grad_norm(vgrads, 0.5)
    self.vf_opt.step(vgrads)
    info["Training/vf_loss"] = vf_loss.item()
    info["grad_norm/vf"] = vnorm.item()

    # ---- actor (ppo.py:42-92); sees the encoder already stepped by the critic
    pparams = [self.pf[k] for k in self.pf_keys]
    for p in pparams:
      p.requires_grad_(True)
    mean = self._forward(self.pf, obs)
    lp, ent, logstd = gaussian_update(mean, self.pf["logstd"], acts)
    with torch.no_grad():
      tmean = self._forward(self.target_pf, obs)
      tlp, _, _ = gaussian_update(tmean, self.target_pf["logstd"], acts)
    ratio = torch.exp(lp - tlp)
    s1 = ratio * advs
    s2 = torch.clamp(ratio, 1.0 - self.clip_para, 1.0 + self.clip_para) * advs
    ploss = -torch.mean(torch.min(s2, s1)) - self.entropy_coeff * ent.mean()
    pgrads = list(torch.autograd.grad(ploss, pparams))
    for p in pparams:
      p.requires_grad_(False)
    pnorm = clip_grad_norm(pgrads, 0.5)
    self.pf_opt.step(pgrads)

    lp_d, ratio_d, logstd_d = lp.detach(), ratio.detach(), logstd.detach()
    info["Training/policy_loss"] = ploss.item()
    info["logprob/mean"] = lp_d.mean().item()
    info["logprob/std"] = lp_d.std().item()
    info["logprob/max"] = lp_d.max().item()
    info["logprob/min"] = lp_d.min().item()
    info["log_std/mean"] = logstd_d.mean().item()
    info["log_std/std"] = logstd_d.std().item()
    info["log_std/max"] = logstd_d.max().item()
    info["log_std/min"] = logstd_d.min().item()
    info["ratio/max"] = ratio_d.max().item()
    info["ratio/min"] = ratio_d.min().item()
    info["grad_norm/pf"] = pnorm.item()
    self._last = {"values": values.detach(), "mean": mean.detach(), "log_prob": lp_d,
                  "vgrads": dict(zip(self.vf_keys, vgrads)),
                  "pgrads": dict(zip(self.pf_keys, pgrads))}
    return info

  def process_epoch_samples(self, roll):
    """on_rl_algo.py:23-34: bootstrap value of next_obs[T-1], masked by terminals[T-1]."""
    last_ob = torch.as_tensor(roll["last_obs"], dtype=self.dtype)
    last_value = self.values(last_ob).double().numpy() * (1 - np.asarray(roll["last_terminals"], np.float64))
    if self.use_gae:
      return gae(roll["rewards"], roll["values"], roll["terminals"], roll["time_limits"],
                 last_value, self.discount, self.tau, True)
    return discount_reward(roll["rewards"], roll["values"], roll["terminals"],
                           roll["time_limits"], last_value, self.discount, True)

  def update_per_epoch(self, roll, perms):
    """ppo.py:28-40. `perms`: one time-row permutation per opt-epoch (the reference draws
    np.random.permutation(T) at on_policy.py:79; passing them in makes the run reproducible).
    Minibatch = batch_size//E time rows x all E envs (on_policy.py:73-92)."""
    advs, rets = self.process_epoch_samples(roll)
    lr_scale = 1.0 - self.current_epoch / float(self.num_epochs)     # utils.py:28-32
    self.pf_opt.lr = self.plr * lr_scale
    self.vf_opt.lr = self.vlr * lr_scale
    for k in self.pf_keys:                                           # utils.py:23-25
      self.target_pf[k].copy_(self.pf[k])
    T, E = roll["rewards"].shape[:2]
    rows = self.batch_size // E
    infos = []
    for ep in range(self.opt_epochs):
      perm = perms[ep]
      for pos in range(0, T, rows):
        idx = perm[pos:pos + rows]
        n = len(idx) * E
        batch = {
          "obs": roll["obs"][idx].reshape(n, -1),
          "acts": roll["acts"][idx].reshape(n, -1),
          "advs": advs[idx].reshape(n, 1),
          "estimate_returns": rets[idx].reshape(n, 1),
          "values": roll["values"][idx].reshape(n, 1),
        }
        infos.append(self.update(batch))
    return advs, rets, infos


class A2COracle:
  """Restates A2C.update (reference torchrl/algo/on_policy/a2c.py:42-107) for SEPARATE actor / critic networks
  (with shared tensors the reference's step order raises inside torch, so only this case has a reference answer).
  Pinned by tests/golden/a2c_mlp.npz (oracle/make_golden_a2c.py)."""

  def __init__(self, family, pf_sd, vf_sd, S, plr=3e-4, vlr=3e-4, entropy_coeff=0.001):
    self.fwd, self.S = FORWARD[family], S
    self.pf, self.vf = pf_sd, vf_sd
    self.pf_keys, self.vf_keys = list(pf_sd.keys()), list(vf_sd.keys())
    self.pf_opt = Adam([pf_sd[k] for k in self.pf_keys], plr)
    self.vf_opt = Adam([vf_sd[k] for k in self.vf_keys], vlr)
    self.entropy_coeff = entropy_coeff

  def update(self, batch):
    obs, acts, advs, rets = (torch.as_tensor(batch[k], dtype=torch.float32)
                             for k in ("obs", "acts", "advs", "estimate_returns"))
    pparams = [self.pf[k] for k in self.pf_keys]
    vparams = [self.vf[k] for k in self.vf_keys]
    for p in pparams + vparams:
      p.requires_grad_(True)
    lp, ent, logstd = gaussian_update(self.fwd(self.pf, obs, self.S), self.pf["logstd"], acts)
    advs = (advs - advs.mean()) / (advs.std() + 1e-5)
    ploss = (-lp * advs).mean() - self.entropy_coeff * ent.mean()
    values = self.fwd(self.vf, obs, self.S)
    vloss = ((values - rets) ** 2).mean()
    pgrads = list(torch.autograd.grad(ploss, pparams))
    vgrads = list(torch.autograd.grad(vloss, vparams))
    for p in pparams + vparams:
      p.requires_grad_(False)
    clip_grad_norm(pgrads, 0.5)
    self.pf_opt.step(pgrads)
    clip_grad_norm(vgrads, 0.5)
    self.vf_opt.step(vgrads)
    return {"Training/policy_loss": ploss.item(), "Training/vf_loss": vloss.item(), "ent": ent.mean().item(),
            "log_prob": lp.mean().item(), "v_pred/mean": values.mean().item(), "v_pred/std": values.std().item()}


def sd_to_torch(pf_np, vf_np, shared_prefixes=("encoder.", "base."), dtype=torch.float32):
  """numpy state dicts -> torch, keeping shared tensors shared (one tensor object)."""
  pf = {k: torch.tensor(v, dtype=dtype) for k, v in pf_np.items()}
  vf = {}
  for k, v in vf_np.items():
    if k.startswith(shared_prefixes) and k in pf:
      vf[k] = pf[k]
    else:
      vf[k] = torch.tensor(v, dtype=dtype)
  return pf, vf
