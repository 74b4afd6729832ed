"""Runs the fused encoder-layer forward at the bench shape a few times for an `ncu --set full` capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_tc import _block_tensors, _ops, DEV
engine, ops = _ops()
B, T = 1024, 17
layer, w, par, x = _block_tensors(B, T, 1)
R = B * T
h = lambda *s: torch.empty(s, device=DEV, dtype=torch.float16)
f = lambda *s: torch.empty(s, device=DEV)
out = {"qkv": h(R, 192), "o": h(R, 64), "h": h(R, 64), "f1": h(R, 256), "y": h(R, 64), "p": f(B, T, T),
       "st1": f(R, 2), "st2": f(R, 2), "xh1": h(R, 64), "xh2": h(R, 64)}
wd = {"w2d": w["w_2"].t().contiguous(), "w1d": w["w_1"].t().contiguous(), "wod": w["w_o"].t().contiguous(),
      "wind": w["w_in"].t().contiguous()}
g = {"dz2": h(R, 64), "df1": h(R, 256), "dh": h(R, 64), "dz1": h(R, 64), "dqkv": h(R, 192), "dx": h(R, 64)}
dy = torch.randn(R, 64, device=DEV).half()
for _ in range(4):
  ops.tc_block_fwd(x, B, T, w, par, out)
  ops.tc_block_bwd(dy, B, T, out, wd, par["g1"], par["g2"], g)
torch.cuda.synchronize()
