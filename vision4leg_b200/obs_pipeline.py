"""Observation pipeline on the device: what the reference does to an observation between the
simulator and the policy, for E environments at once, without a host round trip.

  DepthFrameStack  depth-buffer frames -> sqrt(log(clip(depth)+1)) -> per-env history -> the 4-frame observation
                   (reference vision4leg/envs/locomotion_gym_env_with_rich_information.py:312-336,549-554,620-650)
  Normalizer       the running-mean observation normaliser of NormObs over a vectorised env
                   (reference torchrl/env/base_wrapper.py:44-122), same attribute / method names

Both call the C-ABI kernels of csrc/obs_ops.cu; there is no CPU path.
"""
import copyreg
import importlib
import pickle
import sys
import types

import numpy as np
import torch

from .engine import Ops

REFERENCE_NORMALIZER = ("torchrl.env.base_wrapper", "Normalizer")


def _reference_normalizer_class():
  """The class a `_obs_normalizer_{epoch}.pkl` names (reference torchrl/algo/rl_algo.py:84-90 pickles
  env._obs_normalizer; the viewers unpickle it, starter/*_viewer.py).  The real class when a reference checkout is
  importable; otherwise a stand-in registered under the same module path, so that pickle writes the same GLOBAL and
  the untouched reference can load the file."""
  mod_name, cls_name = REFERENCE_NORMALIZER
  try:
    return getattr(importlib.import_module(mod_name), cls_name)
  except Exception:      # no checkout on the path, or its imports (gym) are missing here
    mod = sys.modules.get(mod_name)
    if mod is None or not hasattr(mod, cls_name):
      mod = types.ModuleType(mod_name)
      mod.__doc__ = "stand-in written by vision4leg_b200.obs_pipeline (wire format of the normaliser pickle only)"
      cls = type(cls_name, (), {"__module__": mod_name, "__qualname__": cls_name})
      setattr(mod, cls_name, cls)
      sys.modules[mod_name] = mod
    return getattr(mod, cls_name)


def reference_normalizer_object(shape, mean, var, count, clip=10., should_estimate=True):
  """an instance of the reference's Normalizer class carrying these statistics (attribute names and order of
  reference torchrl/env/base_wrapper.py:64-71)"""
  cls = _reference_normalizer_class()
  obj = cls.__new__(cls)
  obj.__dict__.update(dict(shape=shape, _mean=np.asarray(mean, np.float64).reshape(shape),
                           _var=np.asarray(var, np.float64).reshape(shape), _count=float(count), clip=clip,
                           should_estimate=bool(should_estimate)))
  return obj


def merge_mean_var_count(mean, var, count, batch_mean, batch_var, batch_count):
  """the reference's update_mean_var_count (torchrl/env/base_wrapper.py:44-61) on float64 tensors"""
  delta = batch_mean - mean
  tot = count + batch_count
  new_mean = mean + delta * (batch_count / tot)
  m2 = var * count + batch_var * batch_count + delta * delta * (count * batch_count / tot)
  return new_mean, m2 / tot, tot


def gather_batch_stats(batch_mean, batch_var, n, process_group):
  """Data-parallel NormObs: every rank holds the statistics of ITS env columns' rows; the batch the reference would
  have seen is their union.  One all-gather of (mean, variance, count) per rank, combined in rank order with the
  same pairwise formula (exact for population variances), so every rank ends with the same float64 numbers as one
  process over all the rows (up to the rounding of the combination, ~1e-15)."""
  import torch.distributed as dist
  world = dist.get_world_size(process_group)
  S = batch_mean.numel()
  mine = torch.cat([batch_mean.reshape(-1), batch_var.reshape(-1),
                    torch.tensor([float(n)], dtype=torch.float64, device=batch_mean.device)])
  allr = [torch.empty_like(mine) for _ in range(world)]
  dist.all_gather(allr, mine, group=process_group)
  m, v, c = allr[0][:S], allr[0][S:2 * S], float(allr[0][2 * S])
  for r in range(1, world):
    cr = float(allr[r][2 * S])
    if cr > 0:
      if c > 0:
        m, v, c = merge_mean_var_count(m, v, c, allr[r][:S], allr[r][S:2 * S], cr)
      else:
        m, v, c = allr[r][:S], allr[r][S:2 * S], cr
  return m, v, c


class _WireUnpickler(pickle.Unpickler):
  """reads a normaliser pickle without needing the reference (or gym) to be importable"""

  def find_class(self, module, name):
    if (module, name) == REFERENCE_NORMALIZER:
      return _reference_normalizer_class()
    return super().find_class(module, name)


def load_reference_normalizer(f):
  """-> dict(shape, _mean, _var, _count, clip, should_estimate) from an `_obs_normalizer_*.pkl` stream"""
  return dict(_WireUnpickler(f).load().__dict__)


def fixed_frame_idx(frame_extract):
  """reference :317-323 (fixed_delay_observation)"""
  return [frame_extract - 1, 2 * frame_extract - 1, 3 * frame_extract - 1, 4 * frame_extract - 1]


def random_frame_idx(rng, frame_extract):
  """reference :325-331 (reset with random delays): one index per block of frame_extract frames"""
  r = rng.randint(0, frame_extract, 4)
  return [int(r[k]) + k * frame_extract for k in range(4)]


def step_frame_idx(rng, frame_idx, frame_extract):
  """reference :549-554 (reset_frame_idx_each_step): a fresh delay in [1, frame_extract) for the newest channel, the
  older channels follow the PREVIOUS indices by one block"""
  return [int(rng.randint(1, frame_extract))] + [int(frame_idx[i]) + frame_extract for i in range(3)]


class DepthFrameStack:
  """The reference keeps a deque of processed frames per env (newest at index 0) and builds the observation from
  depth_frames[frame_idx[k]], k = 0..3.  Here the history is a ring [E, n_slots, 64, 64] fp32 in HBM: deque index
  i is ring slot (head - i) mod n_slots, so a push is one slot write and the observation is a 4-slot gather."""

  def __init__(self, n_envs, num_stored_frames, frame_idx, depth_norm=True, device=None, ops=None):
    self.device = torch.device(device if device is not None else "cuda")
    self.ops = ops if ops is not None else Ops(self.device)
    self.E, self.n = int(n_envs), int(num_stored_frames)
    self.depth_norm = bool(depth_norm)
    self.ring = torch.zeros(self.E, self.n, 64, 64, device=self.device)
    self.head = 0
    self.slots = torch.zeros(self.E, 4, dtype=torch.int32, device=self.device)
    self._reset = torch.zeros(self.E, dtype=torch.uint8, device=self.device)
    self.set_frame_idx(frame_idx)
    self.s2d = torch.empty(self.E, 16, 16, 64, dtype=torch.float16, device=self.device)

  def set_frame_idx(self, frame_idx):
    """frame_idx: 4 deque indices, shared ([4]) or per env ([E, 4]) (reference :315-336, re-drawn per step at
    :549-554 when reset_frame_idx_each_step)"""
    fi = np.asarray(frame_idx, np.int64)
    fi = np.broadcast_to(fi, (self.E, 4)) if fi.ndim == 1 else fi
    if fi.shape != (self.E, 4) or fi.min() < 0 or fi.max() >= self.n:
      raise ValueError("frame_idx must hold 4 indices in [0, num_stored_frames) per env")
    self.frame_idx = torch.as_tensor(np.ascontiguousarray(fi), dtype=torch.int32).to(self.device)

  def push(self, zbuf, reset=None):
    """zbuf: [E, 64, 64] fp32 depth-buffer values on the device; reset: optional [E] bool (episode starts: the new
    frame fills that env's whole history, reference :635-637)"""
    if zbuf.shape != (self.E, 64, 64) or zbuf.dtype != torch.float32 or not zbuf.is_cuda:
      raise ValueError("zbuf must be a CUDA float32 tensor [E, 64, 64]")
    zbuf = zbuf.contiguous()
    self.head = (self.head + 1) % self.n
    rs = None
    if reset is not None:
      self._reset.copy_(torch.as_tensor(reset).to(self.device, torch.uint8))
      rs = self._reset
    self.ops.depth_frame(zbuf, self.ring, rs, self.E, self.n, self.head)

  def observe(self, out_chw=None, out_s2d=True):
    """the stacked observation: fp16 space-to-depth image [E,16,16,64] (what the tensor-core tier's conv1 reads)
    and / or fp32 CHW pixels written into out_chw [E, >=16384] (a view of the observation rows' image part)"""
    torch.remainder(self.head - self.frame_idx, self.n, out=self.slots)
    stride = 0
    if out_chw is not None:
      if out_chw.dtype != torch.float32 or out_chw.shape[0] != self.E or out_chw.stride(-1) != 1:
        raise ValueError("out_chw must be float32 [E, 16384] with unit inner stride")
      stride = out_chw.stride(0)
    self.ops.stack_frames(self.ring, self.slots, self.E, self.n, self.depth_norm,
                          self.s2d if out_s2d else None, out_chw, stride)
    return self.s2d if out_s2d else out_chw


class Normalizer:
  """reference torchrl/env/base_wrapper.py:64-105 for a vectorised env (observations [n, S]); mean and variance live
  on the device in float64, the count on the host (it is 1e-4 + the rows merged so far)."""

  def __init__(self, shape, clip=10., device=None, ops=None):
    self.device = torch.device(device if device is not None else "cuda")
    self.ops = ops if ops is not None else Ops(self.device)
    self.shape = shape
    self.S = int(np.prod(shape))
    self._mean_d = torch.zeros(self.S, dtype=torch.float64, device=self.device)
    self._var_d = torch.ones(self.S, dtype=torch.float64, device=self.device)
    self._count = 1e-4
    self.clip = clip
    self.should_estimate = True

  @property
  def _mean(self):
    return self._mean_d.cpu().numpy().reshape(self.shape)

  @property
  def _var(self):
    return self._var_d.cpu().numpy().reshape(self.shape)

  def stop_update_estimate(self):
    self.should_estimate = False

  def _rows(self, data):
    if data.dtype != torch.float32 or not data.is_cuda or data.shape[-1] != self.S or data.dim() != 2:
      raise ValueError("observations must be a CUDA float32 tensor [n, %d]" % self.S)
    return data.contiguous()

  def update_estimate(self, data, process_group=None):
    """process_group: data-parallel training — the statistics are those of the union of the ranks' rows, identical on
    every rank (the reference is single-process; this is what one process over all the env columns would compute)"""
    if not self.should_estimate:
      return
    x = self._rows(data)
    if process_group is None or torch.distributed.get_world_size(process_group) == 1:
      self.ops.normalizer(x, x.shape[0], self.S, self._mean_d, self._var_d, self._count, 1, self.clip, None)
      self._count += x.shape[0]
      return
    bm, bv = torch.empty_like(self._mean_d), torch.empty_like(self._var_d)
    self.ops.normalizer(x, x.shape[0], self.S, bm, bv, 1.0, 2, self.clip, None)          # this rank's batch statistics
    bm, bv, n = gather_batch_stats(bm, bv, x.shape[0], process_group)
    if n > 0:
      m, v, c = merge_mean_var_count(self._mean_d, self._var_d, self._count, bm, bv, n)
      self._mean_d.copy_(m); self._var_d.copy_(v)
      self._count = c

  def filt(self, raw, out=None):
    x = self._rows(raw)
    out = torch.empty_like(x) if out is None else out
    self.ops.normalizer(x, x.shape[0], self.S, self._mean_d, self._var_d, self._count, False, self.clip, out)
    return out

  filt_torch = filt

  def observation(self, observation, training=True, out=None, process_group=None):
    """NormObs.observation (reference :119-122): update (when training) then filter, one launch"""
    x = self._rows(observation)
    out = torch.empty_like(x) if out is None else out
    upd = bool(training and self.should_estimate)
    if upd and process_group is not None and torch.distributed.get_world_size(process_group) > 1:
      self.update_estimate(x, process_group)
      return self.filt(x, out)
    self.ops.normalizer(x, x.shape[0], self.S, self._mean_d, self._var_d, self._count, upd, self.clip, out)
    if upd:
      self._count += x.shape[0]
    return out

  def inverse_torch(self, raw):
    return raw * torch.sqrt(self._var_d).to(raw.dtype) + self._mean_d.to(raw.dtype)

  # ---- checkpoint wire format: RLAlgo.snapshot pickles env._obs_normalizer; what lands in the file
  # is an instance of the REFERENCE's class (numpy statistics), loadable by the untouched viewers
  def __reduce__(self):
    # copyreg._reconstructor(cls, object, None) + BUILD(state): stdlib names and the reference's class only, so the
    # file loads where this package is not installed.  Unpickling therefore yields the reference-format object
    # (host statistics); Normalizer(shape).load_reference(obj) puts it back on a device.
    ref = self.to_reference()
    return (copyreg._reconstructor, (type(ref), object, None), dict(ref.__dict__))

  def to_reference(self):
    return reference_normalizer_object(self.shape, self._mean, self._var, self._count, self.clip, self.should_estimate)

  def load_reference(self, state):
    """state: a reference Normalizer instance, or the dict of load_reference_normalizer"""
    st = state if isinstance(state, dict) else state.__dict__
    if int(np.prod(st["shape"])) != self.S:
      raise ValueError("normaliser of shape %s loaded into one of shape %s" % (st["shape"], self.shape))
    self._mean_d.copy_(torch.as_tensor(np.asarray(st["_mean"], np.float64).reshape(-1)))
    self._var_d.copy_(torch.as_tensor(np.asarray(st["_var"], np.float64).reshape(-1)))
    self._count = float(st["_count"])
    self.clip = st["clip"]
    self.should_estimate = bool(st["should_estimate"])
    return self

