"""Micro-benchmarks of the tensor-core kernels (CUDA events, warm L2): python tools/bench_tc.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision4leg_b200 import engine  # noqa: E402

DEV = "cuda:0"
ops = engine.ops_for(DEV)
RM = engine.RM


def timeit(fn, reps=50):
  for _ in range(5):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps * 1e3   # us


def linear(M, N, K, f32=False):
  x = torch.randn(M, K, device=DEV).half()
  w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).half()
  b = torch.randn(N, device=DEV)
  out = torch.empty(M, N, device=DEV, dtype=torch.float32 if f32 else torch.float16)
  f = lambda: ops.tc_gemm(x, (M, 1, 1, K), (M, 1, 1), (1, 1, 128), [(0, 0)], K // 64, w, N, N, b, out, RM.dense(N),
                          c_f32=f32, flags=engine.RELU)
  us = timeit(f)
  print("linear  M=%6d N=%4d K=%4d : %7.1f us  %7.1f TFLOP/s" % (M, N, K, us, 2.0 * M * N * K / us / 1e6))


def wgrad(M, N, K):
  x = torch.randn(M, K, device=DEV).half()
  dy = torch.randn(M, N, device=DEV).half()
  dw = torch.empty(N, K, device=DEV)
  f = lambda: ops.tc_wgrad(x, (M, 1, 1, K), dy, N, (M, 1, 1), (1, 1, 128), [(0, 0)], N, None, dw)
  us = timeit(f)
  print("wgrad   M=%6d N=%4d K=%4d : %7.1f us  %7.1f TFLOP/s (incl. reduce)" % (M, N, K, us, 2.0 * M * N * K / us / 1e6))


def conv1(B):
  img = torch.randn(B, 16, 16, 64, device=DEV).half()
  w = (torch.randn(32, 256, device=DEV) / 16).half()
  b = torch.randn(32, device=DEV)
  out = torch.zeros(B, 15, 15, 32, device=DEV, dtype=torch.float16)
  taps = [(dx, dy) for dy in range(2) for dx in range(2)]
  f = lambda: ops.tc_gemm(img, (B, 16, 16, 64), (B, 15, 15), (15, 8, 1), taps, 1, w, 32, 32, b, out,
                          RM(225, 225 * 32, 32, 0), flags=engine.RELU)
  us = timeit(f)
  fl = 2.0 * B * 225 * 32 * 256
  by = B * (32768 + 225 * 64)
  print("conv1   B=%6d             : %7.1f us  %7.1f TFLOP/s  %7.1f GB/s" % (B, us, fl / us / 1e6, by / us / 1e3))


def conv3(B):
  x = torch.randn(B, 6, 6, 64, device=DEV).half()
  w = (torch.randn(64, 576, device=DEV) / 24).half()
  b = torch.randn(64, device=DEV)
  out = torch.zeros(B, 16, 64, device=DEV, dtype=torch.float16)
  taps = [(kw, kh) for kh in range(3) for kw in range(3)]
  f = lambda: ops.tc_gemm(x, (B, 6, 6, 64), (B, 4, 4), (4, 4, 8), taps, 1, w, 64, 64, b, out, RM(16, 1024, 64, 0),
                          flags=engine.RELU)
  us = timeit(f)
  print("conv3   B=%6d             : %7.1f us  %7.1f TFLOP/s" % (B, us, 2.0 * B * 16 * 64 * 576 / us / 1e6))


if __name__ == "__main__":
  if len(sys.argv) > 1 and sys.argv[1] == "one":
    linear(1024, 256, 256)
    conv1(1024)
    sys.exit(0)
  for M, N, K in [(1024, 256, 256), (1024, 64, 256), (1024, 256, 128), (8192, 256, 256), (65536, 256, 256),
                  (17408, 192, 64), (17408, 64, 64), (17408, 256, 64), (17408, 64, 256), (1114112, 256, 64)]:
    linear(M, N, K)
  linear(1024, 16, 256, f32=True)
  for B in (1024, 8192, 65536):
    conv1(B)
  for B in (1024, 65536):
    conv3(B)
  for M, N, K in [(1024, 256, 256), (17408, 192, 64), (17408, 64, 256), (1114112, 256, 64)]:
    wgrad(M, N, K)
