"""Observation pipeline on the device (csrc/obs_ops.cu through the C-ABI) against oracle/obs_oracle.py.

Tolerances: the depth feature is fp32 log / sqrt on both sides (CUDA logf / sqrtf are within 1-2 ulp of libm):
2e-6 absolute on values in [0.51, 1.55]; the fp16 space-to-depth copy is the fp32 value rounded once (exact
against a torch .half() of the oracle value up to that 2e-6 -> 1 fp16 ulp = 1e-3).  The normaliser computes in
float64 like the reference: statistics within 1e-12 relative, outputs equal after the float32 store (1e-6)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _s2d_to_chw(s2d):
  """[E,16,16,64] (Y, X, (py, px, c)) -> [E,4,64,64]"""
  E = s2d.shape[0]
  t = s2d.float().reshape(E, 16, 16, 4, 4, 4)          # Y X py px c
  return t.permute(0, 5, 1, 3, 2, 4).reshape(E, 4, 64, 64)


@pytest.mark.parametrize("depth_norm,per_env_idx", [(True, False), (False, True)])
def test_depth_stack_matches_oracle(depth_norm, per_env_idx):
  from oracle import obs_oracle as oo
  from vision4leg_b200.obs_pipeline import DepthFrameStack
  E, n, steps = 5, 16, 23
  rng = np.random.RandomState(11)
  if per_env_idx:
    fidx = np.stack([oo.random_frame_idx(rng, 4) for _ in range(E)])
  else:
    fidx = np.asarray(oo.fixed_frame_idx(4))
  dev = torch.device("cuda", 0)
  gpu = DepthFrameStack(E, n, fidx, depth_norm=depth_norm, device=dev)
  cpu = [oo.DepthStack(n, depth_norm=depth_norm) for _ in range(E)]
  obs = torch.zeros(E, 93 + 16384, device=dev)         # the reference's observation row: proprio then pixels
  worst = 0.0
  for t in range(steps):
    # depth-buffer values: mostly near 1 (the far field), some close-ups, exact 0 and 1
    z = (1.0 - 10.0 ** rng.uniform(-6, -1, (E, 64, 64))).astype(np.float32)
    z[:, 0, 0], z[:, 0, 1] = 0.0, 1.0
    reset = (rng.rand(E) < 0.15) if t else np.ones(E, bool)
    gpu.push(torch.from_numpy(z).to(dev), reset=torch.from_numpy(reset))
    for e in range(E):
      cpu[e].push(z[e], reset=bool(reset[e]))
    # obs[:, 93:] starts 372 bytes into a row (not 16-byte aligned: scalar stores); every third step an aligned
    # buffer of its own (float4 stores)
    chw = torch.empty(E, 16384, device=dev) if t % 3 == 0 else obs[:, 93:]
    s2d = gpu.observe(out_chw=chw)
    want = np.stack([cpu[e].observe(list(fidx[e]) if per_env_idx else list(fidx)) for e in range(E)])
    got = chw.cpu().numpy()
    worst = max(worst, float(np.abs(got - want).max()))
    assert np.abs(got - want).max() < 2e-6 * (1 / 0.425 if depth_norm else 1) + 1e-7
    back = _s2d_to_chw(s2d).reshape(E, -1).cpu().numpy()
    assert np.array_equal(back, chw.half().float().cpu().numpy())      # the fp16 copy is the same value rounded once
  print("depth stack worst |gpu - oracle| = %.2e" % worst)
  assert float(obs[:, :93].abs().max()) == 0.0         # the proprio columns were not touched


def test_normalizer_matches_reference_fixture():
  import os
  from oracle import make_golden_obs as mk
  from vision4leg_b200.obs_pipeline import Normalizer
  g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "obs_normalizer.npz"))
  dev = torch.device("cuda", 0)
  nz = Normalizer((mk.S,), device=dev)
  for i, x in enumerate(mk.inputs()):
    if i == mk.STEPS - 1:
      nz.stop_update_estimate()
    xt = torch.from_numpy(x.astype(np.float32)).to(dev)          # lossless: the fixture's inputs are fp32 values
    if i % 2:                                             # both call shapes of the reference: update + filt, observation()
      nz.update_estimate(xt)
      out = nz.filt(xt)
    else:
      out = nz.observation(xt, training=True)
    assert np.allclose(out.cpu().numpy(), g["filt"][i], rtol=1e-6, atol=1e-6), i
  assert np.allclose(nz._mean, g["mean"], rtol=1e-12, atol=1e-13)
  assert np.allclose(nz._var, g["var"], rtol=1e-12, atol=1e-13)
  assert abs(nz._count - float(g["count"])) < 1e-9


def test_normalizer_whole_observation_rows():
  """the shipped configs normalise the WHOLE row (proprio + 16384 pixels, get_env.py:78-80): many columns, few rows,
  clipping active, constant columns (variance 0 -> the +1e-4 in the denominator)"""
  from oracle import obs_oracle as oo
  from vision4leg_b200.obs_pipeline import Normalizer
  S, E = 93 + 16384, 4
  rng = np.random.RandomState(5)
  dev = torch.device("cuda", 0)
  nz, ref = Normalizer((S,), device=dev), oo.Normalizer((S,))
  for t in range(3):
    x = (rng.randn(E, S) * rng.uniform(0.01, 100, S)).astype(np.float32)
    x[:, 7] = 3.0
    x[0, 11] = 1e6
    out = nz.observation(torch.from_numpy(x).to(dev))
    ref.update(x)
    want = ref.filt(x)
    assert np.allclose(out.cpu().numpy(), want, rtol=1e-6, atol=1e-6)
  assert np.allclose(nz._var, ref.var, rtol=1e-11, atol=1e-300)
  x[1, 20], x[2, 21] = 1e9, -1e9                          # the evaluation path: filter only, clipping active
  out = nz.filt(torch.from_numpy(x).to(dev))
  assert np.allclose(out.cpu().numpy(), ref.filt(x), rtol=1e-6, atol=1e-6)
  assert out[1, 20].item() == 10.0 and out[2, 21].item() == -10.0


def test_normalizer_checkpoint_roundtrip():
  """pickle.dump(env._obs_normalizer) (RLAlgo.snapshot, reference rl_algo.py:84-90) writes the reference's wire
  format — no name of this package in the stream — and the statistics come back onto a device bit-exactly"""
  import io
  import pickle
  from vision4leg_b200 import obs_pipeline as op
  dev = torch.device("cuda", 0)
  S = 93
  rng = np.random.RandomState(2)
  nz = op.Normalizer((S,), device=dev)
  for _ in range(3):
    nz.update_estimate(torch.from_numpy((rng.randn(8, S) * 3 + 1).astype(np.float32)).to(dev))
  raw = pickle.dumps(nz)
  assert b"vision4leg_b200" not in raw
  st = op.load_reference_normalizer(io.BytesIO(raw))
  assert np.array_equal(st["_mean"], nz._mean) and np.array_equal(st["_var"], nz._var) and st["_count"] == nz._count
  nz2 = op.Normalizer((S,), device=dev).load_reference(st)
  x = torch.from_numpy(rng.randn(5, S).astype(np.float32)).to(dev)
  assert torch.equal(nz2.filt(x), nz.filt(x))
  nz2.update_estimate(x); nz.update_estimate(x)
  assert np.array_equal(nz2._var, nz._var) and nz2._count == nz._count


def test_normalizer_batch_statistics_mode_and_rank_merge():
  """update = 2 writes the batch mean / population variance only; two 'ranks' (halves of one batch) combined with the
  data-parallel formulas give the statistics of one update over the whole batch"""
  from vision4leg_b200 import obs_pipeline as op
  dev = torch.device("cuda", 0)
  S = 16477
  rng = np.random.RandomState(9)
  x = (rng.randn(12, S) * rng.uniform(0.1, 50, S) + 3).astype(np.float32)
  xt = torch.from_numpy(x).to(dev)
  nz = op.Normalizer((S,), device=dev)
  parts = []
  for sl in (slice(0, 5), slice(5, 12)):
    bm, bv = torch.empty(S, dtype=torch.float64, device=dev), torch.empty(S, dtype=torch.float64, device=dev)
    nz.ops.normalizer(xt[sl].contiguous(), sl.stop - sl.start, S, bm, bv, 1.0, 2, 10.0, None)
    x64 = x[sl].astype(np.float64)
    assert np.allclose(bm.cpu().numpy(), x64.mean(0), rtol=1e-13, atol=1e-13)
    assert np.allclose(bv.cpu().numpy(), x64.var(0), rtol=1e-12, atol=1e-13)
    parts.append((bm, bv, sl.stop - sl.start))
  gm, gv, gn = op.merge_mean_var_count(parts[0][0], parts[0][1], parts[0][2], parts[1][0], parts[1][1], parts[1][2])
  m, v, c = op.merge_mean_var_count(nz._mean_d, nz._var_d, nz._count, gm, gv, gn)
  nz.update_estimate(xt)                                   # one process over the union
  assert np.allclose(m.cpu().numpy(), nz._mean, rtol=1e-12, atol=1e-13)
  assert np.allclose(v.cpu().numpy(), nz._var, rtol=1e-12, atol=1e-13) and abs(c - nz._count) < 1e-9
  assert float(torch.zeros(1, device=dev).sum()) == 0.0    # no sticky error from the launches above
