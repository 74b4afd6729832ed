"""Per-kernel time per SAMPLE at several minibatch sizes (torch.profiler / CUPTI over one epoch of graph replays):
finds the kernel whose cost per sample is not flat in the minibatch size.

  python -m tools.trace_sizes --batches 16384 32768 65536 [--minibatches 5] [--envs 8]
"""
import argparse
import collections
import re

import numpy as np
import torch


def run(B, n_mb, E, model="loco", S=93, A=12):
  from benchutil import synth
  from benchutil.harness import build_nets, load_np_sd, make_ppo
  dev = torch.device("cuda", 0)
  T = B // E * n_mb
  N = T * E
  pf, vf = build_nets(model, S, A)
  pf_np, vf_np = synth.make_family_weights(1000, model, S, A)
  load_np_sd(pf, pf_np); load_np_sd(vf, vf_np)
  pf, vf = pf.to(dev), vf.to(dev)
  agent, _ = make_ppo(pf, vf, None, A, B, N, 1, device=dev)
  agent.precision = "f16"
  eng = agent.engine
  r = eng._alloc_rollout(T, E)
  gen = torch.Generator(device=dev); gen.manual_seed(1)
  for n0 in range(0, N, 1 << 15):
    m = min(1 << 15, N - n0)
    d = torch.empty((m, 16, 16, 64), device=dev).uniform_(0.3, 10.0, generator=gen)
    r["imgs"][n0:n0 + m] = ((torch.sqrt(torch.log(d + 1.0)) - 1.25) / 0.425).to(torch.float16)
  del d
  r["state"].normal_(generator=gen).clamp_(-10, 10)
  r["acts"].normal_(generator=gen).mul_(0.15)
  r["rewards"].normal_(generator=gen); r["values"].normal_(generator=gen)
  r["terminals"].zero_(); r["time_limits"] = None
  rng = np.random.default_rng(5)
  last_obs = rng.standard_normal((E, S + 16384)).astype(np.float32)
  eng.compute_advantages(last_obs, np.zeros((E, 1), np.float32), 0.99, 0.95, True, True)
  eng.set_lr(1e-4, 1e-4); eng.sync_target()
  perm = np.arange(T)[None]
  for _ in range(2):
    eng.run_epoch(perm, B)
  torch.cuda.synchronize()
  from torch.profiler import profile, ProfilerActivity
  with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.run_epoch(perm, B)
    torch.cuda.synchronize()
  tot, cnt = collections.Counter(), collections.Counter()
  t_lo, t_hi = None, None
  for e in prof.events():
    if e.device_type is None or "cuda" not in str(e.device_type).lower():
      continue
    name = e.name
    if "memcpy" in name.lower() or "memset" in name.lower():
      continue
    m = re.search(r"(\w+_kernel)", name)
    short = m.group(1) if m else name[:40]
    tot[short] += e.time_range.end - e.time_range.start
    cnt[short] += 1
    t_lo = e.time_range.start if t_lo is None else min(t_lo, e.time_range.start)
    t_hi = e.time_range.end if t_hi is None else max(t_hi, e.time_range.end)
  del agent, eng, r
  torch.cuda.empty_cache()
  return tot, cnt, (t_hi - t_lo), N


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--batches", type=int, nargs="+", default=[16384, 32768, 65536])
  ap.add_argument("--minibatches", type=int, default=5)
  ap.add_argument("--envs", type=int, default=8)
  args = ap.parse_args()
  res = {B: run(B, args.minibatches, args.envs) for B in args.batches}
  names = sorted({n for B in res for n in res[B][0]}, key=lambda n: -res[args.batches[-1]][0].get(n, 0))
  print("ns of kernel time per sample (sum over launches of one epoch / samples); wall = first start .. last end")
  print("%-32s" % "kernel" + "".join("%12d" % B for B in args.batches) + "   launches/minibatch")
  for n in names:
    print("%-32s" % n + "".join("%12.1f" % (res[B][0].get(n, 0) * 1e3 / res[B][3]) for B in args.batches)
          + "   %d" % (res[args.batches[-1]][1][n] // args.minibatches))
  print("%-32s" % "SUM of kernels" + "".join("%12.1f" % (sum(res[B][0].values()) * 1e3 / res[B][3]) for B in args.batches))
  print("%-32s" % "wall" + "".join("%12.1f" % (res[B][2] * 1e3 / res[B][3]) for B in args.batches))
  print("%-32s" % "M samples/s (wall)" + "".join("%12.3f" % (res[B][3] / res[B][2]) for B in args.batches))


if __name__ == "__main__":
  main()
