"""CPU oracle for the observation pipeline — TEST INFRASTRUCTURE, never imported by the product.

Restates, in numpy, what the reference does to an observation between the simulator and the policy:
  * depth frame -> feature:  vision4leg/envs/locomotion_gym_env_with_rich_information.py:620-633
  * k-frame stacking through a deque indexed by frame_idx:  :312-336 (reset), :549-554 (per step), :635-650
  * running-mean normaliser:  torchrl/env/base_wrapper.py:44-61 (merge), :64-90 (Normalizer), :119-122 (filt)

Parity status: the normaliser is pinned against the LIVE reference classes (oracle/make_golden_obs.py ->
tests/golden/obs_normalizer.npz, checked by tests/test_oracle_golden.py).  The depth path lives inside the
pybullet environment, which cannot be imported here (pybullet absent): it is restated from the source lines
above and is "parity unpinned" beyond the arithmetic identities the tests check.
"""
import collections

import numpy as np

NEAR, FAR = 0.01, 1000          # reference :623-624


def depth_feature(zbuf):
  """reference :620-633 (no blinding spots): z-buffer value -> sqrt(log(clip(metric depth, 0.3, 10) + 1)), float32."""
  depth = np.asarray(zbuf, np.float32)
  depth = FAR * NEAR / (FAR - (FAR - NEAR) * depth)
  depth = np.clip(depth, a_min=0.3, a_max=10)
  return np.sqrt(np.log(depth + 1)).astype(np.float32)


def fixed_frame_idx(frame_extract):
  """reference :317-323 (fixed_delay_observation)."""
  return [frame_extract - 1, 2 * frame_extract - 1, 3 * frame_extract - 1, 4 * frame_extract - 1]


def random_frame_idx(rng, frame_extract):
  """reference :325-331."""
  r = rng.randint(0, frame_extract, 4)
  return [int(r[0]), int(r[1]) + frame_extract, int(r[2]) + 2 * frame_extract, int(r[3]) + 3 * frame_extract]


def step_frame_idx(rng, frame_idx, frame_extract):
  """reference :549-554 (reset_frame_idx_each_step)."""
  return [rng.randint(1, frame_extract)] + [frame_idx[i] + frame_extract for i in range(3)]


class DepthStack:
  """One environment's depth_frames deque (reference :635-650): newest frame at index 0; a reset fills every slot
  with the first frame; the observation concatenates depth_frames[idx] for idx in frame_idx, then (x-1.25)/0.425."""

  def __init__(self, num_stored_frames, depth_norm=True):
    self.n = num_stored_frames
    self.depth_norm = depth_norm
    self.frames = collections.deque(maxlen=num_stored_frames)

  def push(self, zbuf, reset=False):
    f = depth_feature(zbuf)[np.newaxis, ...]
    for _ in range(self.n if reset else 1):
      self.frames.appendleft(f)

  def observe(self, frame_idx):
    out = np.concatenate([self.frames[i] for i in frame_idx], axis=0).reshape(-1)
    if self.depth_norm:
      out = (out - 1.25) / 0.425
    return out.astype(np.float32)


def merge_mean_var_count(mean, var, count, batch_mean, batch_var, batch_count):
  """reference torchrl/env/base_wrapper.py:44-61."""
  delta = batch_mean - mean
  tot = count + batch_count
  new_mean = mean + delta * batch_count / tot
  M2 = var * count + batch_var * batch_count + np.square(delta) * count * batch_count / tot
  return new_mean, M2 / tot, tot


class Normalizer:
  """reference base_wrapper.py:64-90: mean 0, var 1, count 1e-4 at start; clip 10."""

  def __init__(self, shape, clip=10.):
    self.mean = np.zeros(shape, np.float64)
    self.var = np.ones(shape, np.float64)
    self.count = 1e-4
    self.clip = clip

  def update(self, x):
    x = np.asarray(x, np.float64)      # the environments hand float64 rows to NormObs (flatten_observations)
    if x.ndim == 1:
      x = x[None]
    self.mean, self.var, self.count = merge_mean_var_count(
      self.mean, self.var, self.count, np.mean(x, axis=0), np.var(x, axis=0), x.shape[0])

  def filt(self, x):
    x = np.asarray(x, np.float64)
    return np.clip((x - self.mean) / (np.sqrt(self.var) + 1e-4), -self.clip, self.clip)
