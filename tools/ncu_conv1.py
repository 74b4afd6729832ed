"""Runs conv1 forward (tap-box tc_gemm, then flat) a few times for an `ncu --set full` capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vision4leg_b200 import engine
from vision4leg_b200.engine import RM, RELU
DEV = "cuda:0"
ops = engine.ops_for(DEV)
torch.manual_seed(0)
B = 1024; Nimg = 4 * B
taps2 = [(dx, dy) for dy in range(2) for dx in range(2)]
oh, ow = np.meshgrid(np.arange(15), np.arange(15), indexing="ij")
pos = ((oh // 2) * 8 + ow // 2) * 128 + ((oh % 2) * 2 + ow % 2) * 32
pos_a1 = torch.tensor(pos.ravel().astype(np.int32), device=DEV)
imgs = (torch.randn(Nimg, 16, 16, 64, device=DEV) * 0.5).half()
idx = torch.randperm(Nimg, device=DEV)[:B].int().contiguous()
w1 = (torch.randn(32, 256, device=DEV) * 0.05).half(); b1 = torch.randn(32, device=DEV) * 0.1
out = torch.zeros(B, 8, 8, 128, device=DEV, dtype=torch.float16)
cm = lambda: RM(225, 8 * 8 * 128, 0, 0, pos_off=pos_a1)
for _ in range(3):
  ops.tc_gemm(imgs, (Nimg, 16, 16, 64), (B, 15, 15), (15, 8, 1), taps2, 1, w1, 32, 32, b1, out, cm(), flags=RELU, a_idx=idx)
for _ in range(3):
  ops.tc_conv_flat(imgs, 64, 256, 16, 15, 15, taps2, w1, 32, 32, b1, out, cm(), B, x_idx=idx, flags=RELU, mode=1)
torch.cuda.synchronize()
