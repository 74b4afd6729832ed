"""Time conv1 forward at B=1024: tap-box tc_gemm vs single-load flat kernel vs swapped-role kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vision4leg_b200 import engine
from vision4leg_b200.engine import RM, RELU
DEV = "cuda:0"
ops = engine.ops_for(DEV)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
Nimg = 4 * B
taps2 = [(dx, dy) for dy in range(2) for dx in range(2)]
oh, ow = np.meshgrid(np.arange(15), np.arange(15), indexing="ij")
pos = torch.tensor((((oh // 2) * 8 + ow // 2) * 128 + ((oh % 2) * 2 + ow % 2) * 32).ravel().astype(np.int32), device=DEV)
x = (torch.randn(Nimg, 16, 16, 64, device=DEV) * 0.5).half()
idx = torch.randperm(Nimg, device=DEV)[:B].int().contiguous()
w = (torch.randn(32, 256, device=DEV) * 0.05).half(); b = torch.randn(32, device=DEV) * 0.1
out = torch.zeros(B, 8, 8, 128, device=DEV, dtype=torch.float16)
cm = lambda: RM(225, 8 * 8 * 128, 0, 0, pos_off=pos)
fns = {"tap-box": lambda: ops.tc_gemm(x, (Nimg, 16, 16, 64), (B, 15, 15), (15, 8, 1), taps2, 1, w, 32, 32, b, out, cm(), flags=RELU, a_idx=idx),
       "flat": lambda: ops.tc_conv_flat(x, 64, 256, 16, 15, 15, taps2, w, 32, 32, b, out, cm(), B, x_idx=idx, flags=RELU, mode=1 + 32),
       "swapped": lambda: ops.tc_conv_flat(x, 64, 256, 16, 15, 15, taps2, w, 32, 32, b, out, cm(), B, x_idx=idx, flags=RELU, mode=0x101)}
for name, f in fns.items():
  for _ in range(3):
    f()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(30):
    f()
  e1.record(); torch.cuda.synchronize()
  print("%s B=%d: %.1f us" % (name, B, e0.elapsed_time(e1) * 1e3 / 30))
