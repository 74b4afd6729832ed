"""Runs a few PPO minibatch updates (no CUDA graph) for ncu: 
  ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py --model loco --minibatches 3
The first minibatch is the warm-up (skip its launches with -s)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--model", default="loco")
  ap.add_argument("--minibatches", type=int, default=3)
  ap.add_argument("--batch", type=int, default=1024)
  ap.add_argument("--S", type=int, default=93)
  ap.add_argument("--A", type=int, default=12)
  ap.add_argument("--precision", default="f16")
  args = ap.parse_args()
  from benchutil import synth
  from benchutil.harness import build_nets, load_np_sd, make_ppo
  E = 8
  T = args.batch // E * args.minibatches
  pf, vf = build_nets(args.model, args.S, args.A)
  pf_np, vf_np = synth.make_family_weights(1000, args.model, args.S, args.A)
  load_np_sd(pf, pf_np); load_np_sd(vf, vf_np)
  pf, vf = pf.cuda(), vf.cuda()
  agent, _ = make_ppo(pf, vf, None, args.A, args.batch, T * E, 1, device="cuda:0")
  agent.use_cuda_graph = False
  agent.precision = args.precision
  eng = agent.engine
  roll = synth.make_rollout(3, T, E, args.S, args.A)
  eng.load_rollout_arrays(roll)
  eng.compute_advantages(roll["last_obs"], roll["last_terminals"], 0.99, 0.95, True, True)
  eng.set_lr(1e-4, 1e-4)
  eng.sync_target()
  torch.cuda.synchronize()
  l0 = eng.ops.launches
  infos = eng.run_epoch(np.arange(T)[None], args.batch)
  torch.cuda.synchronize()
  print("minibatches", len(infos), "launches", eng.ops.launches - l0, "vf_loss", infos[-1]["Training/vf_loss"])


if __name__ == "__main__":
  main()
