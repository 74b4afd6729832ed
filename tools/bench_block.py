"""Time the fused transformer-layer forward kernel alone (CUDA events, 50 launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_tc import _block_tensors, _ops, DEV

engine, ops = _ops()
for B in (1024, 8192, 65536):
  T = 17
  layer, w, par, x = _block_tensors(B, T, 1)
  R = B * T
  h = lambda *s: torch.empty(s, device=DEV, dtype=torch.float16)
  f = lambda *s: torch.empty(s, device=DEV)
  out = {"qkv": h(R, 192), "o": h(R, 64), "h": h(R, 64), "f1": h(R, 256), "y": h(R, 64),
         "p": f(B, T, T), "z1": f(R, 64), "st1": f(R, 2), "z2": f(R, 64), "st2": f(R, 2)}
  for _ in range(5):
    ops.tc_block_fwd(x, B, T, w, par, out)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(50):
    ops.tc_block_fwd(x, B, T, w, par, out)
  e1.record(); torch.cuda.synchronize()
  us = e0.elapsed_time(e1) * 1e3 / 50
  flops = 2.0 * R * (64 * 192 + 64 * 64 + 64 * 256 * 2) + 2.0 * (R / 119) * 128 * 128 * 64 * 2
  byts = R * (64 * 2 + 192 * 2 + 64 * 2 * 3 + 256 * 2 + 64 * 4 * 2 + 16 + T * 4)
  print("B=%d  %.1f us/launch  %.1f TFLOP/s  %.0f GB/s stored+loaded" % (B, us, flops / us / 1e6, byts / us / 1e3))
