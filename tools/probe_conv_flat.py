"""Flat (single-load, descriptor-shifted) trunk convolutions vs the tap-box tc_gemm path: error per
mode and time per launch (CUDA events)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vision4leg_b200 import engine
from vision4leg_b200.engine import RM, RELU

DEV = "cuda:0"
ops = engine.ops_for(DEV)
torch.manual_seed(0)


def timeit(fn, n=30):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n):
    fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) * 1e3 / n


def rel(a, b):
  return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12))


B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
Nimg = 4 * B
taps2 = [(dx, dy) for dy in range(2) for dx in range(2)]
taps3 = [(kw, kh) for kh in range(3) for kw in range(3)]
oh, ow = np.meshgrid(np.arange(15), np.arange(15), indexing="ij")
pos = ((oh // 2) * 8 + ow // 2) * 128 + ((oh % 2) * 2 + ow % 2) * 32
pos_a1 = torch.tensor(pos.ravel().astype(np.int32), device=DEV)

imgs = (torch.randn(Nimg, 16, 16, 64, device=DEV) * 0.5).half()
idx = torch.randperm(Nimg, device=DEV)[:B].int().contiguous()
cases = []
# conv1
w1 = (torch.randn(32, 4 * 64, device=DEV) * 0.05).half(); b1 = torch.randn(32, device=DEV) * 0.1
cases.append(("conv1", imgs, (Nimg, 16, 16, 64), (B, 15, 15), (15, 8, 1), taps2, 1, w1, 32, 32, b1, (B, 8, 8, 128),
              lambda: RM(225, 8 * 8 * 128, 0, 0, pos_off=pos_a1), idx, dict(C_=64, P=256, Wg=16, Hout=15, Wout=15)))
a1c = (torch.randn(B, 8, 8, 128, device=DEV) * 0.5).half()
w2 = (torch.randn(64, 4 * 128, device=DEV) * 0.05).half(); b2 = torch.randn(64, device=DEV) * 0.1
cases.append(("conv2", a1c, (B, 8, 8, 128), (B, 6, 6), (6, 6, 3), taps2, 2, w2, 64, 64, b2, (B, 6, 6, 64),
              lambda: RM(36, 36 * 64, 64, 0), None, dict(C_=128, P=64, Wg=8, Hout=6, Wout=6)))
a2 = (torch.randn(B, 6, 6, 64, device=DEV) * 0.5).half()
w3 = (torch.randn(64, 9 * 64, device=DEV) * 0.05).half(); b3 = torch.randn(64, device=DEV) * 0.1
cases.append(("conv3", a2, (B, 6, 6, 64), (B, 4, 4), (4, 4, 8), taps3, 1, w3, 64, 64, b3, (B, 16, 64),
              lambda: RM(16, 16 * 64, 64, 0), None, dict(C_=64, P=36, Wg=6, Hout=4, Wout=4)))

for name, x, xs, og, box, taps, kch, w, Np, Nv, bias, oshape, cmap, xi, fl in cases:
  ref = torch.zeros(oshape, device=DEV, dtype=torch.float16)
  f_ref = lambda: ops.tc_gemm(x, xs, og, box, taps, kch, w, Np, Nv, bias, ref, cmap(), flags=RELU, a_idx=xi)
  f_ref(); t_ref = timeit(f_ref)
  line = "%s B=%d  tap-box %.1f us |" % (name, B, t_ref)
  for mode in (1 + 16, 1 + 32, 1 + 64):
    out = torch.zeros(oshape, device=DEV, dtype=torch.float16)
    f = lambda: ops.tc_conv_flat(x, fl["C_"], fl["P"], fl["Wg"], fl["Hout"], fl["Wout"], taps, w, Np, Nv, bias, out, cmap(),
                                 B, x_idx=xi, flags=RELU, mode=mode)
    try:
      f(); torch.cuda.synchronize()
      e = rel(out, ref)
      t = timeit(f)
      line += "  group %d: err %.2e %.1f us" % (mode >> 4, e, t)
    except Exception as ex:
      line += "  mode %d: FAILED %s" % (mode, str(ex)[:60])
  print(line)
