"""Time the observation-pipeline kernels (csrc/obs_ops.cu) with CUDA events and report them against the HBM roofline.

  python -m tools.bench_obs [--envs 1024] [--iters 200]

Algorithmic bytes per env and step:
  depth_frame   read 16 KB z-buffer + write 16 KB ring slot                               = 32 KB
  stack_frames  read 4 x 16 KB ring slots + write 32 KB fp16 image (+ 64 KB fp32 CHW row)  = 96 KB (160 KB with CHW)
  normalizer    read the row 3x (mean, variance, filter; 2 of them L2 hits) + write it     = 2 x 4 (S) bytes counted
Inputs are larger than L2 at the default size (1024 envs x 16 slots x 16 KB = 256 MB ring)."""
import argparse
import json
import os

import torch

from vision4leg_b200.obs_pipeline import DepthFrameStack, Normalizer


def timed(fn, iters):
  for _ in range(5):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(iters):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) * 1e3 / iters           # us


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--envs", type=int, default=1024)
  ap.add_argument("--iters", type=int, default=200)
  args = ap.parse_args()
  dev = torch.device("cuda", 0)
  E, S = args.envs, 93 + 16384
  peaks = {}
  pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
  if os.path.exists(pk):
    peaks = json.load(open(pk))
  peak = float(peaks.get("hbm_gbs", 6500.0))
  st = DepthFrameStack(E, 16, [3, 7, 11, 15], depth_norm=True, device=dev)
  z = torch.rand(E, 64, 64, device=dev)
  obs = torch.zeros(E, S, device=dev)
  nz = Normalizer((S,), device=dev)
  out = torch.empty_like(obs)
  rows = []
  t = timed(lambda: st.ops.depth_frame(z, st.ring, None, E, 16, 3), args.iters)
  rows.append(("depth_frame", t, E * 32768))
  st.push(z, reset=torch.ones(E, dtype=torch.bool))
  t = timed(lambda: st.observe(), args.iters)
  rows.append(("stack_frames (fp16 s2d)", t, E * (65536 + 32768)))
  t = timed(lambda: st.observe(out_chw=obs[:, 93:]), args.iters)
  rows.append(("stack_frames (+fp32 CHW row, unaligned)", t, E * (65536 + 32768 + 65536)))
  Es = min(E, 64)                                          # the normaliser sees one row per env and step
  t = timed(lambda: nz.ops.normalizer(obs[:Es], Es, S, nz._mean_d, nz._var_d, 1.0, True, 10.0, out[:Es]), args.iters)
  rows.append(("normalizer update+filt (%d rows)" % Es, t, Es * S * 8))
  for name, us, nbytes in rows:
    gbps = nbytes / us * 1e-3
    print("%-44s %8.1f us  %8.1f GB/s  frac of %.0f GB/s = %.3f" % (name, us, gbps, peak, gbps / peak))


if __name__ == "__main__":
  main()
