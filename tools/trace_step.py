"""Kernel timeline of CUDA-graph replays of the PPO minibatch step (torch.profiler / CUPTI): prints, for one
replay in steady state, every kernel with stream, start offset, duration and the idle gap before it on its
stream, plus busy/idle totals of the main stream."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--model", default="loco")
  ap.add_argument("--batch", type=int, default=1024)
  ap.add_argument("--minibatches", type=int, default=8)
  args = ap.parse_args()
  from benchutil import synth
  from benchutil.harness import build_nets, load_np_sd, make_ppo
  E, S, A = 8, 93, 12
  T = args.batch // E * args.minibatches
  pf, vf = build_nets(args.model, S, A)
  pf_np, vf_np = synth.make_family_weights(1000, args.model, S, A)
  load_np_sd(pf, pf_np); load_np_sd(vf, vf_np)
  pf, vf = pf.cuda(), vf.cuda()
  agent, _ = make_ppo(pf, vf, None, A, args.batch, T * E, 1, device="cuda:0")
  agent.precision = "f16"
  eng = agent.engine
  roll = synth.make_rollout(3, T, E, S, A)
  eng.load_rollout_arrays(roll)
  eng.compute_advantages(roll["last_obs"], roll["last_terminals"], 0.99, 0.95, True, True)
  eng.set_lr(1e-4, 1e-4)
  eng.sync_target()
  for _ in range(2):                       # warm-up: eager pass, capture, replays
    eng.run_epoch(np.arange(T)[None], args.batch)
  torch.cuda.synchronize()
  from torch.profiler import profile, ProfilerActivity
  with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.run_epoch(np.arange(T)[None], args.batch)
    torch.cuda.synchronize()
  evs = [e for e in prof.events() if e.device_type is not None and "cuda" in str(e.device_type).lower()]
  ks = sorted(((e.time_range.start, e.time_range.end, e.name, getattr(e, "device_resource_id", getattr(e, "stream", -1)))
               for e in evs if "memcpy" not in e.name.lower() and "memset" not in e.name.lower()), key=lambda x: x[0])
  print("kernels recorded:", len(ks))
  starts = [i for i, k in enumerate(ks) if "mb_begin" in k[2] or "slot_advance" in k[2]]
  if len(starts) < 4:
    print("could not find minibatch boundaries", len(starts)); return
  a, b = starts[3], starts[4]
  t0 = ks[a][0]
  main_stream = ks[a][3]
  last_end = {}
  busy = 0.0
  print("minibatch duration (us): %.1f   launches: %d" % (ks[b][0] - t0, b - a))
  for s_, e_, n_, st in ks[a:b]:
    gap = s_ - last_end.get(st, s_)
    last_end[st] = e_
    if st == main_stream:
      busy += e_ - s_
    import re as _re
    _m = _re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*(?:<[^()]*>)?\s*\(", n_)
    _m = [x for x in _m if x not in ("void", "namespace")]
    short = (_m[0] if _m else n_)[:28]
    print("%8.1f  s%-4s %-28s dur %6.1f  gap %6.1f" % (s_ - t0, st, short, e_ - s_, gap))
  print("main stream busy %.1f us of %.1f" % (busy, ks[b][0] - t0))


if __name__ == "__main__":
  main()
