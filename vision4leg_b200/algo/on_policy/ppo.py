"""PPO with the update phase on the B200 (reference torchrl/algo/on_policy/ppo.py:10-161).

Same constructor, attributes and logged keys as the reference class.  `update_per_epoch()`
streams the epoch's rollout from the (pinned) replay buffer to the device once, runs the GAE
scan there and then replays one captured CUDA graph per minibatch — critic step, then actor
step against the frozen target policy, each a fused clip+Adam over a flat bucket — and reads
the 18 logged statistics of every minibatch back in one copy at the end of the epoch.
`update(batch)` (one minibatch from numpy arrays) is kept for API compatibility and runs the
same kernels eagerly.
"""
import copy

import numpy as np
import torch

from .a2c import A2C
from .ppo_engine import PPOUpdateEngine
from .. import utils as atu


class PPO(A2C):
  def __init__(self, pf, clip_para=0.2, opt_epochs=10, clipped_value_loss=False, **kwargs):
    self.target_pf = copy.deepcopy(pf)
    super().__init__(pf=pf, **kwargs)
    self.clip_para = clip_para
    self.opt_epochs = opt_epochs
    self.clipped_value_loss = clipped_value_loss
    self.sample_key = ["obs", "acts", "advs", "estimate_returns", "values"]
    self.process_group = None       # set to a torch.distributed group for data-parallel updates
    self.use_cuda_graph = True
    self.precision = "fp32"         # "fp32": exact CUDA-core tier; "f16" (or "fp16"): tcgen05 tensor-core tier
    self.half_image_staging = True  # f16 tier: stream the buffer's fp16 copy of the depth stack (same values)
    self._engine = None

  @property
  def engine(self):
    if self._engine is None:
      if self.precision == "fp16":
        self.precision = "f16"
      self._engine = PPOUpdateEngine(self.pf, self.vf, self.target_pf, self.device, self.clip_para,
                                     self.entropy_coeff, self.clipped_value_loss,
                                     use_cuda_graph=self.use_cuda_graph,
                                     process_group=self.process_group, precision=self.precision)
    return self._engine

  def _schedule(self):
    atu.update_linear_schedule(self.pf_optimizer, self.current_epoch, self.num_epochs, self.plr)
    atu.update_linear_schedule(self.vf_optimizer, self.current_epoch, self.num_epochs, self.vlr)
    self.engine.set_lr(self.pf_optimizer.param_groups[0]["lr"], self.vf_optimizer.param_groups[0]["lr"])

  def _draw_perms(self, T):
    # one np.random.permutation(T) per opt-epoch, drawn in the order the reference's
    # one_iteration generators would draw them (on_policy.py:77-79)
    if self.shuffle:
      return np.stack([np.random.permutation(T) for _ in range(self.opt_epochs)])
    return np.stack([np.arange(T) for _ in range(self.opt_epochs)])

  def update_per_epoch(self):
    eng, buf = self.engine, self.replay_buffer
    if self.precision == "f16" and self.half_image_staging and eng.has_img and hasattr(buf, "enable_half_image_staging"):
      # the tensor-core tier consumes the depth stack in fp16: let the buffer keep a pinned fp16 copy
      # (maintained by add_sample; rows already stored are converted once, here) and stream that
      buf.enable_half_image_staging(eng.S)
    eng.load_rollout(buf, stream_obs=True)
    sample = buf.last_sample(["next_obs", "terminals"])
    eng.compute_advantages(sample["next_obs"], sample["terminals"], self.discount, self.tau,
                           buf.time_limit_filter, self.gae)
    self._schedule()
    eng.sync_target()
    infos = eng.run_epoch(self._draw_perms(buf._max_replay_buffer_size), self.batch_size)
    self.training_update_num += len(infos)
    for info in infos:
      self.logger.add_update_info(info)
    # keep the buffer's public arrays coherent with what the update used
    r = eng._roll
    buf._advs = r["advs"].cpu().numpy().reshape(r["T"], r["E"], 1)
    buf._estimate_returns = r["rets"].cpu().numpy().reshape(r["T"], r["E"], 1)

  def update(self, batch):
    """One minibatch from host arrays (reference ppo.py:125-153)."""
    eng = self.engine
    n = np.shape(batch["obs"])[0]
    roll = {"obs": np.asarray(batch["obs"], np.float32).reshape(n, 1, -1),
            "acts": np.asarray(batch["acts"], np.float32).reshape(n, 1, -1),
            "values": np.asarray(batch["values"], np.float32).reshape(n, 1, 1),
            "rewards": np.zeros((n, 1, 1), np.float32), "terminals": np.zeros((n, 1, 1), np.float32)}
    r = eng.load_rollout_arrays(roll)
    r["advs"].copy_(torch.as_tensor(np.asarray(batch["advs"], np.float32).reshape(n)))
    r["rets"].copy_(torch.as_tensor(np.asarray(batch["estimate_returns"], np.float32).reshape(n)))
    if not hasattr(self, "current_epoch"):
      self.current_epoch = 0
    eng.set_lr(self.pf_optimizer.param_groups[0]["lr"], self.vf_optimizer.param_groups[0]["lr"])
    graphs, eng.use_cuda_graph = eng.use_cuda_graph, False
    try:
      infos = eng.run_epoch(np.arange(n)[None, :], n)
    finally:
      eng.use_cuda_graph = graphs
    self.training_update_num += 1
    return infos[0]

  @property
  def networks(self):
    return [self.pf, self.vf, self.target_pf]
