"""On-policy rollout buffer with the reference's API over pinned-host fp32 storage and a CUDA
GAE scan (reference torchrl/replay_buffers/on_policy.py:10-92, base.py:10-55).

Differences that do not change results (DESIGN.md §Rollout store):
  * arrays are float32 in page-locked memory (the reference's are float64 np.zeros):
    every consumer casts to fp32 before use (ppo.py:136-140), and pinned memory is what lets the
    update stream the rows to the device asynchronously;
  * `next_obs` keeps only the row the algorithm ever reads — index max-1 (on_policy.py:13-14);
  * `generalized_advantage_estimation` / `discount_reward` run the reverse segmented scan on
    the GPU in float64 registers and write `_advs` / `_estimate_returns` back as numpy arrays.
"""
import numpy as np
import torch

from .. import engine
from .._lib import V4LError


class BaseReplayBuffer:
  def __init__(self, max_replay_buffer_size, env_nums=1, time_limit_filter=False):
    self.env_nums = env_nums
    self._max_replay_buffer_size = max_replay_buffer_size // self.env_nums
    self._top = 0
    self._size = 0
    self.time_limit_filter = time_limit_filter
    self._host = {}            # key -> pinned torch tensor backing the numpy view `_key`
    self._half_S = None        # proprio width once half-precision image staging is enabled
    self.device = None         # set by the algorithm (PPO) or defaults to the current device

  # ---- storage --------------------------------------------------------------------------------
  def _alloc(self, key, shape):
    rows = 1 if key == "next_obs" else self._max_replay_buffer_size
    t = torch.zeros((rows,) + tuple(shape), dtype=torch.float32)
    if torch.cuda.is_available():
      t = t.pin_memory()
    self._host[key] = t
    setattr(self, "_" + key, t.numpy())

  def add_sample(self, sample_dict, **kwargs):
    for key, val in sample_dict.items():
      if key not in self._host:
        self._alloc(key, np.shape(val))
      if key == "next_obs":
        if self._top == self._max_replay_buffer_size - 1:
          self._next_obs[0, ...] = val
      else:
        getattr(self, "_" + key)[self._top, ...] = val
        if key == "obs" and self._half_S is not None:
          self._write_half_staging(self._top)
    self._advance()

  # ---- half-precision staging of the depth stack ---------------------------------------------------
  def enable_half_image_staging(self, state_dim):
    """Keep, next to the fp32 observation rows, a pinned fp16 copy of their image part and a packed
    fp32 copy of their proprio part, maintained by add_sample.  The tensor-core tier of the PPO update
    rounds the depth stack to fp16 anyway (round-to-nearest on the device == numpy's cast here), so
    streaming this copy instead of the fp32 rows halves the host->device bytes of every update without
    changing a single value it computes.  Rows already stored are converted once."""
    S = int(state_dim)
    if self._half_S == S:
      return
    self._half_S = S
    if "obs" in self._host:
      self._alloc_half()
      for t in range(self._size if self._size < self._max_replay_buffer_size else self._max_replay_buffer_size):
        self._write_half_staging(t)

  def _alloc_half(self):
    T, E, D = self._obs.shape
    S = self._half_S
    pin = lambda t: t.pin_memory() if torch.cuda.is_available() else t
    self._host["obs_img16"] = pin(torch.zeros((T, E, D - S), dtype=torch.float16))
    self._host["obs_state"] = pin(torch.zeros((T, E, max(S, 1)), dtype=torch.float32))
    self._obs_img16, self._obs_state = self._host["obs_img16"].numpy(), self._host["obs_state"].numpy()

  def _write_half_staging(self, t):
    if "obs_img16" not in self._host:
      self._alloc_half()
    S = self._half_S
    self._obs_img16[t] = self._obs[t, :, S:]
    if S:
      self._obs_state[t, :, :S] = self._obs[t, :, :S]

  def terminate_episode(self):
    pass

  def _advance(self):
    self._top = (self._top + 1) % self._max_replay_buffer_size
    if self._size < self._max_replay_buffer_size:
      self._size += 1

  def num_steps_can_sample(self):
    return self._size

  def random_batch(self, batch_size, sample_key):
    assert batch_size % self.env_nums == 0, "batch size should be dividable by env_nums"
    rows = batch_size // self.env_nums
    idx = np.random.randint(0, self.num_steps_can_sample(), rows)
    return self._gather(idx, sample_key)

  def _gather(self, idx, sample_key):
    out = {}
    for key in sample_key:
      arr = getattr(self, "_" + key)[idx]
      out[key] = arr.reshape((arr.shape[0] * arr.shape[1],) + arr.shape[2:])
    return out


class OnPolicyReplayBufferBase:
  def last_sample(self, sample_key):
    out = {}
    for key in sample_key:
      arr = getattr(self, "_" + key)
      out[key] = arr[0] if key == "next_obs" else arr[self._max_replay_buffer_size - 1]
    return out

  # ---- GAE on the GPU -------------------------------------------------------------------------
  def _scan(self, last_value, gamma, tau, mode):
    dev = self.device
    if dev is None:
      if not torch.cuda.is_available():
        raise V4LError("OnPolicyReplayBuffer: GAE runs on the GPU and no CUDA device is available")
      dev = torch.device("cuda", torch.cuda.current_device())
    ops = engine.ops_for(dev)
    T, E = self._rewards.shape[0], self.env_nums
    up = lambda k: self._host[k].reshape(T, -1).to(ops.device, non_blocking=True)
    r, v, d = up("rewards"), up("values"), up("terminals")
    tl = up("time_limits") if "time_limits" in self._host else None
    tl_st, tl_se = (tl.shape[1], 1) if (tl is not None and tl.shape[1] == E and E > 1) else (1, 0)
    lv = torch.as_tensor(np.asarray(last_value, np.float32).reshape(E)).to(ops.device)
    advs = torch.empty((T, E), device=ops.device, dtype=torch.float32)
    rets = torch.empty((T, E), device=ops.device, dtype=torch.float32)
    use_tl = bool(self.time_limit_filter and tl is not None)
    ops.gae(r, v, d, tl, tl_st, tl_se, lv, advs, rets, T, E, float(gamma), float(tau), use_tl, mode)
    self._advs_dev, self._rets_dev = advs, rets
    self._advs = advs.cpu().numpy().reshape(T, E, 1)
    self._estimate_returns = rets.cpu().numpy().reshape(T, E, 1)

  def generalized_advantage_estimation(self, last_value, gamma, tau):
    """A_t = delta_t + (1-d_t) gamma tau A_{t+1}, optionally zeroed at time limits
    (reference on_policy.py:17-45)."""
    self._scan(last_value, gamma, tau, 0)

  def discount_reward(self, last_value, gamma):
    """reference on_policy.py:47-71"""
    self._scan(last_value, gamma, 1.0, 1)

  def one_iteration(self, batch_size, sample_key, shuffle):
    """Time-row minibatches as numpy copies (reference on_policy.py:73-92).  PPO's CUDA path
    does not use this (it gathers rows on the device) — it exists for API compatibility."""
    assert batch_size % self.env_nums == 0, "batch size should be dividable by env_nums"
    rows = batch_size // self.env_nums
    n = self._max_replay_buffer_size
    indices = np.random.permutation(n) if shuffle else np.arange(n)
    for pos in range(0, n, rows):
      yield self._gather(indices[pos:pos + rows], sample_key)


class OnPolicyReplayBuffer(OnPolicyReplayBufferBase, BaseReplayBuffer):
  pass
