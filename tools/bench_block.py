"""Time the fused encoder-layer kernels alone (CUDA events, 50 launches) and print CTA 0's phase
timeline (v4l_tc_block_timeline)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_tc import _block_tensors, _ops, DEV

engine, ops = _ops()


def timeit(fn, n=50):
  for _ in range(5):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n):
    fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) * 1e3 / n


for B in (1024, 8192, 65536):
  T = 17
  layer, w, par, x = _block_tensors(B, T, 1)
  R = B * T
  h = lambda *s: torch.empty(s, device=DEV, dtype=torch.float16)
  f = lambda *s: torch.empty(s, device=DEV)
  out = {"qkv": h(R, 192), "o": h(R, 64), "h": h(R, 64), "f1": h(R, 256), "y": h(R, 64), "p": f(B, T, T),
         "st1": f(R, 2), "st2": f(R, 2), "xh1": h(R, 64), "xh2": h(R, 64)}
  us = timeit(lambda: ops.tc_block_fwd(x, B, T, w, par, out))
  flops = 2.0 * R * (64 * 192 + 64 * 64 + 64 * 256 * 2) + 2.0 * (R / 119) * 128 * 128 * 64 * 2
  byts = R * (64 * 2 * 2 + 192 * 2 + 64 * 2 * 4 + 256 * 2 + 16 + T * 4)
  print("fwd B=%d  %.1f us/launch  %.1f TFLOP/s(useful)  %.0f GB/s" % (B, us, flops / us / 1e6, byts / us / 1e3))
  wd = {"w2d": w["w_2"].t().contiguous(), "w1d": w["w_1"].t().contiguous(), "wod": w["w_o"].t().contiguous(),
        "wind": w["w_in"].t().contiguous()}
  g = {"dz2": h(R, 64), "df1": h(R, 256), "dh": h(R, 64), "dz1": h(R, 64), "dqkv": h(R, 192), "dx": h(R, 64)}
  dy = torch.randn(R, 64, device=DEV).half()
  us = timeit(lambda: ops.tc_block_bwd(dy, B, T, out, wd, par["g1"], par["g2"], g))
  byts = R * (64 * 2 * 3 + 192 * 2 + 256 * 2 + 8 + T * 4 + 64 * 2 * 4 + 256 * 2 + 192 * 2)
  print("bwd B=%d  %.1f us/launch  %.1f TFLOP/s(useful)  %.0f GB/s" % (B, us, flops / us / 1e6, byts / us / 1e3))
  if B == 1024:
    buf = (C.c_uint64 * 64)()
    from vision4leg_b200 import _lib
    _lib.check(_lib.load().v4l_tc_block_timeline(buf))
    for which, name in ((0, "fwd"), (1, "bwd")):
      t = [buf[which * 32 + i] for i in range(32)]
      t0 = t[0]
      print(name, "timeline (us from start):", " ".join("%d:%.2f" % (i, (t[i] - t0) / 1e3) for i in range(21) if t[i]))
