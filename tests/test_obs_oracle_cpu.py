"""Observation-pipeline oracle (oracle/obs_oracle.py): the normaliser against the fixture produced by the live
reference class (oracle/make_golden_obs.py), the depth path against the arithmetic identities of the source lines
it restates (the pybullet environment cannot be imported here)."""
import os

import numpy as np

from oracle import obs_oracle as oo
from oracle import make_golden_obs as mk

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "obs_normalizer.npz")


def test_normalizer_matches_reference_fixture():
  g = np.load(GOLD)
  nz = oo.Normalizer((mk.S,))
  xs = mk.inputs()
  for i, x in enumerate(xs):
    if i < mk.STEPS - 1:
      nz.update(x)
    assert np.array_equal(nz.filt(x), g["filt"][i]), i        # same float64 arithmetic: bit-exact
  assert np.array_equal(nz.mean, g["mean"]) and np.array_equal(nz.var, g["var"])
  assert nz.count == float(g["count"])


def test_depth_feature_range_and_monotone():
  z = np.linspace(0.0, 1.0, 4096, dtype=np.float32).reshape(64, 64)
  f = oo.depth_feature(z)
  assert f.dtype == np.float32
  lo, hi = np.sqrt(np.log(np.float32(1.3))), np.sqrt(np.log(np.float32(11.0)))
  assert abs(f.min() - lo) < 1e-6 and abs(f.max() - hi) < 1e-6     # clip [0.3, 10]
  assert np.all(np.diff(f.reshape(-1)) >= 0)
  # a depth-buffer value that maps to 2 m
  zz = np.float32((oo.FAR - oo.FAR * oo.NEAR / 2.0) / (oo.FAR - oo.NEAR))
  assert abs(float(oo.depth_feature(np.full((1, 1), zz, np.float32))[0, 0]) - np.sqrt(np.log(3.0))) < 1e-3


def test_depth_stack_is_a_deque_of_processed_frames():
  rng = np.random.RandomState(0)
  st = oo.DepthStack(16, depth_norm=True)
  frames = [rng.uniform(0.9, 1.0, (64, 64)).astype(np.float32) for _ in range(20)]
  st.push(frames[0], reset=True)
  idx = oo.fixed_frame_idx(4)
  assert idx == [3, 7, 11, 15]
  o = st.observe(idx).reshape(4, 64, 64)
  for c in range(4):                              # after a reset every slot holds the first frame
    assert np.array_equal(o[c], ((oo.depth_feature(frames[0]) - 1.25) / 0.425).astype(np.float32))
  for f in frames[1:]:
    st.push(f)
  o = st.observe(idx).reshape(4, 64, 64)
  for c, i in enumerate(idx):                     # deque index i = the frame pushed i steps ago
    assert np.array_equal(o[c], ((oo.depth_feature(frames[19 - i]) - 1.25) / 0.425).astype(np.float32))
  r = oo.random_frame_idx(np.random.RandomState(3), 4)
  assert all(4 * k <= r[k] < 4 * (k + 1) for k in range(4))


def test_frame_index_helpers_of_the_product_match_the_oracle():
  """host logic of vision4leg_b200.obs_pipeline (no device needed): same draws from the same generator state"""
  from vision4leg_b200 import obs_pipeline as op
  assert op.fixed_frame_idx(4) == oo.fixed_frame_idx(4)
  assert op.random_frame_idx(np.random.RandomState(5), 4) == oo.random_frame_idx(np.random.RandomState(5), 4)
  a, b = np.random.RandomState(6), np.random.RandomState(6)
  fa = fb = oo.fixed_frame_idx(4)
  for _ in range(5):
    fa, fb = op.step_frame_idx(a, fa, 4), oo.step_frame_idx(b, fb, 4)
    assert fa == [int(v) for v in fb] and 1 <= fa[0] < 4
