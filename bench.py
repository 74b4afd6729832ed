#!/usr/bin/env python
"""PPO-update throughput benchmark (BASELINE.json metric: PPO-update samples/sec on
64x64x4 depth + proprio rollouts).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--model loco|nature] [--impl reference]

One "step" = one PPO.update_per_epoch() over a synthetic rollout of T x E transitions:
GAE scan + opt_epochs x (T*E/B) minibatch updates (critic fwd/bwd/clip/Adam, actor fwd/bwd +
frozen-target fwd/clip/Adam) = opt_epochs*T*E sample-updates.
  value : rollout already resident in HBM when the timed region starts (CUDA events, max over
          ranks);
  e2e   : the same step through the public call PPO.update_per_epoch() with the rollout in
          pinned HOST memory — the H2D copy of the whole rollout and the D2H read of the logged
          statistics are inside the timed region.
N > 1 (torchrun, one rank per GPU): every rank owns its own T x E shard (weak scaling), the
minibatch is the union over ranks, gradients are all-reduced once per optimiser step (NCCL) and
the advantage statistics are global.
--impl reference: the reference's CPU torch path (oracle port, see oracle/ppo_oracle.py) timed
on this box's host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE = {"loco": 70.31e6, "nature": 50.81e6}       # BASELINE.md §2 (S=93, A=12)
OBS_BYTES = lambda S: (S + 16384) * 4


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=5)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--model", default="loco", choices=["loco", "nature"])
  ap.add_argument("--T", type=int, default=2048)
  ap.add_argument("--E", type=int, default=8)
  ap.add_argument("--batch", type=int, default=1024)
  ap.add_argument("--opt-epochs", type=int, default=3)
  ap.add_argument("--S", type=int, default=93)
  ap.add_argument("--A", type=int, default=12)
  ap.add_argument("--no-graph", action="store_true")
  ap.add_argument("--precision", default="f16", choices=["fp32", "f16"],
                  help="f16: tcgen05 tensor-core tier (fp16 operands, fp32 accumulate); fp32: exact CUDA-core tier")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--cpu-seconds", type=float, default=12.0)
  return ap.parse_args()


def peaks():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    p = json.load(open(path))
    return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
            "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "src": "measured"}
  return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
  """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
  Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
       "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
       "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, index=0):
    self.index, self.rows, self.stamps, self.proc = index, [], [], None

  def start(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits", "-lms", "20"],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.stamps.append(time.perf_counter())
      self.rows.append([c.strip() for c in line.split(",")])

  def stop(self, window=None):
    """window = (t0, t1) in time.perf_counter() seconds: keep the samples taken inside the timed region
    (the sampler is started one warm-up step early so that nvidia-smi is already running by then)."""
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    time.sleep(0.05)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    if window is not None:
      keep = [r for t, r in zip(self.stamps, self.rows) if window[0] <= t <= window[1] + 0.02]
      if keep:
        self.rows = keep
    sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
    mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
    reasons = set()
    for r in self.rows:
      if len(r) < 9:
        continue
      for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                        ("sw_power_cap", 8)):
        if r[col].lower().startswith("active"):
          reasons.add(name)
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
# CPU side: the reference's torch path (oracle port) on a bounded sample
# -------------------------------------------------------------------------------------------------
def make_oracle(model, S, A, batch):
  from oracle import ppo_oracle as po, synth
  pf_np, vf_np = synth.make_family_weights(1000, model, S, A)
  pf, vf = po.sd_to_torch(pf_np, vf_np)
  return po.PPOOracle(model, pf, vf, S, batch_size=batch, opt_epochs=1)


def cpu_minibatch(model, S, A, batch, seed):
  from oracle import synth
  rng = np.random.default_rng(seed)
  roll = synth.make_rollout(seed, batch // 8, 8, S, A)
  return {"obs": roll["obs"].reshape(batch, -1), "acts": roll["acts"].reshape(batch, -1),
          "advs": rng.standard_normal((batch, 1)).astype(np.float32),
          "estimate_returns": rng.standard_normal((batch, 1)).astype(np.float32),
          "values": roll["values"].reshape(batch, 1)}


def host_cores():
  """Usable host cores: min(affinity mask, cgroup CPU quota)."""
  cores = os.cpu_count() or 1
  try:
    cores = min(cores, len(os.sched_getaffinity(0)))
  except Exception:
    pass
  try:
    quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
    if quota != "max":
      cores = max(1, min(cores, int(float(quota) / float(period))))
  except Exception:
    pass
  return cores


def pick_threads():
  """torch-CPU threads for the reference arm: the fastest of a few candidates on a small
  conv fwd+bwd probe (oversubscribing a shared 128-core host is 50x slower than 16 threads)."""
  import torch.nn.functional as F
  cores = host_cores()
  cands = sorted({c for c in (4, 8, 16, 32, 64, cores) if c <= cores} | {min(cores, 8)})
  x = torch.randn(128, 4, 64, 64)
  w = torch.randn(32, 4, 8, 8, requires_grad=True)
  best, best_t = cands[0], float("inf")
  for c in cands:
    torch.set_num_threads(c)
    for rep in range(3):
      t0 = time.perf_counter()
      F.conv2d(x, w, stride=4).sum().backward()
      dt = time.perf_counter() - t0
      if rep and dt < best_t:
        best, best_t = c, dt
  torch.set_num_threads(best)
  return best, cores


def cpu_baseline(args, budget_s):
  """samples/s of reference-equivalent PPO.update on the host cores (tensors pre-converted)."""
  cores, avail = pick_threads()
  orc = make_oracle(args.model, args.S, args.A, args.batch)
  mb = cpu_minibatch(args.model, args.S, args.A, args.batch, 5)
  mb = {k: torch.as_tensor(v, dtype=torch.float32) for k, v in mb.items()}
  orc.update(mb)                                          # warm-up
  n, t0 = 0, time.perf_counter()
  while True:
    orc.update(mb)
    n += 1
    dt = time.perf_counter() - t0
    if dt >= budget_s or n >= 32:
      break
  return {"value": n * args.batch / dt, "unit": "samples/s", "cores": cores, "kind": "port",
          "sample": "%d PPO.update minibatches of %d (%s, S=%d, A=%d) after 1 warm-up, %.1f s, %d torch "
                    "threads (best of a probe; %d usable cores); oracle/ppo_oracle.py = torch-CPU "
                    "restatement of the reference path" %
                    (n, args.batch, args.model, args.S, args.A, dt, cores, avail)}


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  cores, avail = pick_threads()
  orc = make_oracle(args.model, args.S, args.A, args.batch)
  mb = cpu_minibatch(args.model, args.S, args.A, args.batch, 5)
  mb = {k: torch.as_tensor(v, dtype=torch.float32) for k, v in mb.items()}
  per_step = 2                                           # minibatches per step (bounded sample)
  for _ in range(args.warmup):
    orc.update(mb)
  t0 = time.perf_counter()
  for _ in range(args.steps * per_step):
    orc.update(mb)
  dt = time.perf_counter() - t0
  value = args.steps * per_step * args.batch / dt
  sample = "%d steps x %d PPO.update minibatches of %d, %d torch threads (%d usable cores)" % (
    args.steps, per_step, args.batch, cores, avail)
  print(json.dumps({
    "impl": "reference", "metric": "ppo_update_samples_per_sec", "value": value, "unit": "samples/s",
    "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
    "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
    "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, 1),
    "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
    "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }))


def workload_config(args, world):
  return {"workload": "ppo_%s update_per_epoch: T=%d x E=%d transitions/rank, minibatch %d/rank, "
                      "opt_epochs=%d, S=%d proprio + 4x64x64 depth, A=%d (BASELINE configs[%d])" %
                      ("locotransformer" if args.model == "loco" else "nature_cnn", args.T, args.E,
                       args.batch, args.opt_epochs, args.S, args.A, 2 if args.model == "loco" else 1),
          "global_batch": args.batch * world, "parallelism": "dp%d" % world,
          "l2": "inputs larger than L2 (%.2f GB fp32 rollout/rank on the host, %.2f GB resident as fp16 in the f16 tier; "
                "rows visited in a fresh permutation every opt-epoch)" %
                (args.T * args.E * OBS_BYTES(args.S) / 1e9, args.T * args.E * (args.S * 4 + 32768) / 1e9)}


# -------------------------------------------------------------------------------------------------
# GPU side
# -------------------------------------------------------------------------------------------------
def main():
  args = parse()
  if args.impl == "reference":
    run_reference(args)
    return
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device (there is no CPU path for the product arm); "
                     "use --impl reference for the CPU baseline")
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  pg = None
  if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD

  from oracle import synth
  from tests._harness import build_nets, load_np_sd, fill_buffer, make_ppo
  from vision4leg_b200.replay_buffers import OnPolicyReplayBuffer

  S, A, T, E = args.S, args.A, args.T, args.E
  pf, vf = build_nets(args.model, S, A)
  pf_np, vf_np = synth.make_family_weights(1000, args.model, S, A)      # same weights on all ranks
  load_np_sd(pf, pf_np); load_np_sd(vf, vf_np)
  pf, vf = pf.to(dev), vf.to(dev)

  # synthetic rollout in the pinned host buffer (generated in chunks to bound host memory)
  buf = OnPolicyReplayBuffer(env_nums=E, max_replay_buffer_size=T * E, time_limit_filter=True)
  chunk = 256
  for t0 in range(0, T, chunk):
    n = min(chunk, T - t0)
    roll = synth.make_rollout(1000 * rank + t0, n, E, S, A)
    for t in range(n):
      buf.add_sample({"obs": roll["obs"][t], "next_obs": roll["last_obs"], "acts": roll["acts"][t],
                      "values": roll["values"][t], "rewards": roll["rewards"][t],
                      "terminals": roll["terminals"][t], "time_limits": roll["time_limits"][t]})
  agent, logger = make_ppo(pf, vf, buf, A, args.batch, T * E, args.opt_epochs, device=dev)
  agent.process_group = pg
  agent.use_cuda_graph = not args.no_graph
  agent.precision = args.precision
  eng = agent.engine
  samples_per_step = args.opt_epochs * T * E * world

  def barrier():
    if world > 1:
      import torch.distributed as dist
      dist.barrier()
    torch.cuda.synchronize(dev)

  def max_over_ranks(ms):
    if world > 1:
      import torch.distributed as dist
      t = torch.tensor([ms], device=dev, dtype=torch.float64)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      return float(t)
    return ms

  # ---- device-resident step: rollout already in HBM
  np.random.seed(0)
  eng.load_rollout(buf)
  last = buf.last_sample(["next_obs", "terminals"])

  def resident_step(epoch):
    agent.current_epoch = epoch
    eng.compute_advantages(last["next_obs"], last["terminals"], agent.discount, agent.tau, True, True)
    agent._schedule()
    eng.sync_target()
    return eng.run_epoch(agent._draw_perms(T), args.batch)

  sampler = ClockSampler(local)
  for w in range(args.warmup):
    if rank == 0 and w == args.warmup - 1:
      sampler.start()                       # nvidia-smi needs ~100 ms to come up: start one warm-up step early
    resident_step(w)
  if rank == 0 and args.warmup == 0:
    sampler.start()
  barrier()
  launches0 = eng.ops.launches
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t_w0 = time.perf_counter()
  ev0.record()
  for k in range(args.steps):
    resident_step(args.warmup + k)
  ev1.record()
  barrier()
  t_w1 = time.perf_counter()
  ms = max_over_ranks(ev0.elapsed_time(ev1))
  launches = eng.ops.launches - launches0
  clocks = sampler.stop((t_w0, t_w1)) if rank == 0 else None
  value = samples_per_step * args.steps / (ms / 1e3)

  # ---- end-to-end step through the public API with HOST buffers
  for w in range(2):
    agent.current_epoch = w
    agent.update_per_epoch()
  barrier()
  t0 = time.perf_counter()
  ev0.record()
  for k in range(args.steps):
    agent.current_epoch = 10 + k
    agent.update_per_epoch()
  ev1.record()
  barrier()
  e2e_ms = max_over_ranks(max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3))
  e2e_value = samples_per_step * args.steps / (e2e_ms / 1e3)

  if world > 1:
    # no collective is issued past this point; ranks leave without tearing NCCL down (destroying a
    # communicator that captured CUDA graphs still reference can block) — hard exit after flushing
    import torch.distributed as dist
    dist.barrier()
    torch.cuda.synchronize(dev)
    if rank != 0:
      sys.stdout.flush()
      os._exit(0)
  pk = peaks()
  roof = dominant_kernel_roofline(args, eng, pk)
  line = {
    "metric": "ppo_update_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
    "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
    "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
    "dtype": "f16" if agent.precision == "f16" else "f32", "data": "synthetic",
    "config": workload_config(args, world),
    "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(eng.h2d_bytes),
            "d2h_bytes_per_step": int(eng.d2h_bytes + 2 * T * E * 4), "ms_per_step": e2e_ms / args.steps},
    "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
    "step_roofline": {"bound": "tensor", "achieved": value / world * FLOP_PER_SAMPLE[args.model] / 1e12,
                      "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                      "frac": value / world * FLOP_PER_SAMPLE[args.model] / 1e12 / pk["bf16_tflops_sustained"],
                      "hbm_frac": value / world * OBS_BYTES(S) / 1e9 / pk["hbm_gbs"],
                      "note": "whole step per GPU (BASELINE.md algorithmic FLOPs) vs %s bf16 sustained peak" % pk["src"]},
  }
  if not args.no_cpu_baseline and world == 1:
    line["cpu_baseline"] = cpu_baseline(args, args.cpu_seconds)
  print(json.dumps(line))
  sys.stdout.flush()
  if world > 1:
    os._exit(0)


def dominant_kernel_roofline(args, eng, pk):
  """Times the dominant kernel of the step alone with CUDA events on the launching stream:
  the conv1 weight-gradient GEMM (M = B*225 im2col rows, N = 32, K = 256; profiles/ has the ncu
  launch list it was picked from)."""
  from vision4leg_b200 import engine as E
  ops, B = eng.ops, args.batch
  if eng.precision == "f16":
    if args.model == "loco":
      roof = tc_block_roofline(args, eng, pk)
      try:
        roof["second_kernel"] = tc_conv1_roofline(args, eng, pk)
      except Exception as ex:          # never let the secondary entry take the bench line down
        roof["second_kernel"] = {"error": str(ex)[:200]}
      return roof
    return tc_conv1_roofline(args, eng, pk)
  plan = eng.plan_pf
  trunk = plan.trunk
  da1 = torch.randn(B, 225, 32, device=ops.device)
  w = torch.empty(32, 256, device=ops.device)
  b = torch.empty(32, device=ops.device)
  idx = eng._bufs(B)["cur_idx"]
  inp = eng._input(B, idx)
  img_map = E.RM(225, inp.img_stride, 0, inp.img_base, idx=idx, pos_off=trunk.pos1)
  run = lambda: ops.linear_wgrad(da1, E.RM.dense(32), inp.img, img_map, trunk.k1, w, b, B * 225, 32, 256)
  for _ in range(3):
    run()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  reps = 20
  e0.record()
  for _ in range(reps):
    run()
  e1.record()
  torch.cuda.synchronize()
  sec = e0.elapsed_time(e1) / 1e3 / reps
  flops = 2.0 * B * 225 * 32 * 257
  ach = flops / sec / 1e12
  return {"kernel": "wgrad_kernel+wgrad_reduce_kernel (conv1 dW, fp32 CUDA-core)", "bound": "tensor",
          "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
          "traffic": None, "us_per_launch": sec * 1e6, "peak_src": pk["src"],
          "algorithmic_flops_per_launch": flops,
          "hbm_bytes_per_launch_algorithmic": B * 65536 + B * 225 * 32 * 4}


def tc_block_roofline(args, eng, pk):
  """Tensor-core tier, LocoTransformer: the fused encoder-layer forward kernel (tc_block_fwd_kernel:
  six chained tcgen05 contractions per 7-sample tile) is the largest single-kernel share of the step
  (6 launches per minibatch, profiles/r1_f16_launches_summary.md).  Timed alone with CUDA events.
  Algorithmic work per sample (17 tokens, d=64, FFN 256): 4 projections 2*17*64*(192+64+256+256) FLOP +
  attention 2*2*17*17*64 FLOP = 1.745 MFLOP; algorithmic HBM bytes per sample = x in + y out + everything
  the backward needs (qkv, o, h, f1, the two normalised rows: fp16; P and the two (mean, rstd): fp32)
  = 17*(64+64+192+64+64+256+64+64)*2 + 17*17*4 + 17*16 = 31 076 B  =>  56 FLOP/B, left of the 251 FLOP/B
  ridge: the kernel is bounded by the HBM roof (the stores for the backward), not the tensor roof."""
  ops, B = eng.ops, args.batch
  plan = eng.plan_pf
  flat = eng.pf_flat
  T, d = plan.T, plan.d
  R = B * T
  p = "visual_append_layers.0."
  x = plan.buf("tok0", (B, T, d))
  w = {"w_in": plan.W.fwd[p + "self_attn.in_proj_weight"].w, "w_o": plan.W.fwd[p + "self_attn.out_proj.weight"].w,
       "w_1": plan.W.fwd[p + "linear1.weight"].w, "w_2": plan.W.fwd[p + "linear2.weight"].w}
  par = {"b_in": plan._view(flat, p + "self_attn.in_proj_bias"), "b_o": plan._view(flat, p + "self_attn.out_proj.bias"),
         "g1": plan._view(flat, p + "norm1.weight"), "be1": plan._view(flat, p + "norm1.bias"),
         "b1": plan._view(flat, p + "linear1.bias"), "b2": plan._view(flat, p + "linear2.bias"),
         "g2": plan._view(flat, p + "norm2.weight"), "be2": plan._view(flat, p + "norm2.bias")}
  h16 = lambda *s: torch.empty(s, device=ops.device, dtype=torch.float16)
  f32 = lambda *s: torch.empty(s, device=ops.device)
  out = dict(qkv=h16(R, 192), o=h16(R, d), h=h16(R, d), f1=h16(R, 256), y=h16(R, d), p=f32(B, 1, T, T),
             st1=f32(R, 2), st2=f32(R, 2), xh1=h16(R, d), xh2=h16(R, d))
  run = lambda: ops.tc_block_fwd(x, B, T, w, par, out)
  for _ in range(3):
    run()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  reps = 20
  e0.record()
  for _ in range(reps):
    run()
  e1.record()
  torch.cuda.synchronize()
  sec = e0.elapsed_time(e1) / 1e3 / reps
  flops = B * (2.0 * T * d * (192 + 64 + 256 + 256) + 4.0 * T * T * d)
  by = B * (T * (64 + 64 + 192 + 64 + 64 + 256 + 64 + 64) * 2 + T * T * 4 + T * 16)
  gbs = by / sec / 1e9
  return {"kernel": "tc_block_fwd_kernel (one TransformerEncoderLayer forward: QKV, block-diagonal attention, "
                    "out-proj, LN, FFN, LN as six chained tcgen05.mma contractions; TMA loads/stores)",
          "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
          # dram__bytes_read.sum + dram__bytes_write.sum of one launch at minibatch 1024 from the
          # `ncu --set full` capture (profiles/r1_tc_block_ncu.txt)
          "traffic": TC_BLOCK_FWD_DRAM_BYTES_B1024 if B == 1024 else None,
          "us_per_launch": sec * 1e6, "peak_src": pk["src"],
          "algorithmic_bytes_per_launch": by, "algorithmic_flops_per_launch": flops,
          "tensor_tflops_achieved": flops / sec / 1e12, "tensor_frac": flops / sec / 1e12 / pk["bf16_tflops"],
          "note": "minibatch 1024 = 147 tiles = ONE wave of 148 SMs: the launch time is the latency of one tile's "
                  "six-deep MMA->epilogue chain (profiles/r1_tc_block_timeline.txt), so the fraction grows with "
                  "the minibatch (tools/bench_block.py: 2.0 TB/s at 65536)"}


TC_BLOCK_FWD_DRAM_BYTES_B1024 = 2430720 + 17152     # dram__bytes_read.sum + dram__bytes_write.sum, profiles/r1_tc_block_ncu.txt


def tc_conv1_roofline(args, eng, pk):
  """Tensor-core tier: the conv1 forward launch of tc_gemm_kernel (the largest GEMM of the step:
  M = B*225 output pixels, N = 32, K = 256 as 4 tap-shifted TMA boxes of the space-to-depth image)
  timed alone with CUDA events."""
  from vision4leg_b200 import engine as E
  ops, B = eng.ops, args.batch
  plan, r = eng.plan_pf, eng._roll
  idx = eng._bufs(B)["cur_idx"]
  a1c = plan.buf("a1c", (B, 8, 8, 128), zero=True)
  pre = "encoder.depth_visual_base.layers." if args.model == "loco" else "encoder.visual_base.layers."
  pkw = plan.W.fwd[pre + "0.weight"]
  bias = plan._view(eng.pf_flat, pre + "0.bias")
  # the launch the plan issues (engine_tc._trunk_fwd): single-load kernel unless FLAT_CONV1 is off
  from vision4leg_b200 import engine_tc as ET
  cmap = lambda: E.RM(225, 8 * 8 * 128, 0, 0, pos_off=plan.pos_a1)
  tap_box = lambda: ops.tc_gemm(r["imgs"], (r["imgs"].shape[0], 16, 16, 64), (B, 15, 15), (15, 8, 1), plan.taps2, 1,
                                pkw.w, pkw.rows, 32, bias, a1c, cmap(), flags=E.RELU, a_idx=idx)
  flat = lambda: ops.tc_conv_flat(r["imgs"], 64, 256, 16, 15, 15, plan.taps2, pkw.w, pkw.rows, 32, bias, a1c, cmap(), B,
                                  x_idx=idx, flags=E.RELU, mode=1 + (2 << 4))
  run, kname = tap_box, "tc_gemm_kernel (conv1 forward: tcgen05.mma fp16, 4 tap-shifted TMA boxes, fused bias+ReLU)"
  if getattr(ET, "FLAT_CONV1", False):
    try:
      flat()
      torch.cuda.synchronize()
      run, kname = flat, ("tc_conv_flat_kernel (conv1 forward: tcgen05.mma fp16, one TMA load per 128-position tile, "
                          "taps = shifted UMMA descriptors, fused bias+ReLU)")
    except Exception:
      pass
  for _ in range(3):
    run()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  reps = 20
  e0.record()
  for _ in range(reps):
    run()
  e1.record()
  torch.cuda.synchronize()
  sec = e0.elapsed_time(e1) / 1e3 / reps
  flops = 2.0 * B * 225 * 32 * 256
  by = B * (16 * 16 * 64 * 2 + 225 * 32 * 2)
  ach = flops / sec / 1e12
  return {"kernel": kname,
          "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
          "frac": ach / pk["bf16_tflops"],
          # dram__bytes_read.sum + dram__bytes_write.sum of this launch at minibatch 1024 from the
          # `ncu --set full` captures in profiles/r1_conv1_tc_gemm_ncu.txt: 33.67 MB + 0.07 MB for the tap-box
          # form, 35.95 MB + 0.12 MB for the single-load form (the 16.8 MB output stays in L2)
          "traffic": (36.07e6 if run is flat else 33.74e6) if B == 1024 else None,
          "us_per_launch": sec * 1e6, "peak_src": pk["src"],
          "algorithmic_flops_per_launch": flops, "hbm_bytes_per_launch_algorithmic": by,
          "hbm_gbs_achieved": by / sec / 1e9, "hbm_frac": by / sec / 1e9 / pk["hbm_gbs"],
          "note": "N=32: every tcgen05.mma (K=16) still streams its 128x16 A operand from shared memory "
                  "(~128 cycles), so a 128-row tile costs 16 x 128 cycles whatever N is: 12.5 % of the tensor "
                  "roof is this shape's ceiling with A in shared memory (measured: ring depth, epilogue "
                  "warpgroups, accumulator stages, a single-load descriptor-shifted variant (tc_conv.cu) and a "
                  "table-driven issue loop all leave 25 us unchanged)"}


if __name__ == "__main__":
  main()
