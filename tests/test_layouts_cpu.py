"""Host-side layout logic of the tensor-core tier, checked on the CPU against torch's convolutions:
the weight-packing index tables (forward and data-gradient orientation), the 4x4 space-to-depth image,
the conv1-output "cells" addressing and the projector's K permutation are what turn the reference's
Conv2d / Linear layers (torchrl/networks/base.py:304-342, 209-230) into "sum over taps of a shifted box x
packed weight slice".  The kernels only ever see these tables, so an error here would be a silent layout
bug; no GPU is needed to pin it."""
import numpy as np
import torch
import torch.nn.functional as F

from vision4leg_b200 import engine_tc as ET


def _pack(w, table):
  flat = np.concatenate([w.reshape(-1), [0.0]])
  return flat[np.where(table >= 0, table, flat.size - 1)]


def _s2d_image(img):
  """[n,4,64,64] -> [n,16,16,64], channel (py*4+px)*4+c (csrc/tc_gemm.cu ingest_img_kernel)"""
  n = img.shape[0]
  x = img.reshape(n, 4, 16, 4, 16, 4)                  # n c Y py X px
  return x.transpose(0, 2, 4, 3, 5, 1).reshape(n, 16, 16, 64)


def _conv_taps(x, wp, taps, out_h, out_w, chunk):
  """sum over taps (dw, dh) of x[:, h+dh, w+dw, :] @ wp[:, tap*chunk:(tap+1)*chunk]^T"""
  n = x.shape[0]
  out = np.zeros((n, out_h, out_w, wp.shape[0]))
  for t, (dw, dh) in enumerate(taps):
    out += x[:, dh:dh + out_h, dw:dw + out_w, :] @ wp[:, t * chunk:(t + 1) * chunk].T
  return out


TAPS2 = [(dx, dy) for dy in range(2) for dx in range(2)]
TAPS3 = [(kw, kh) for kh in range(3) for kw in range(3)]


def test_conv_trunk_tables_reproduce_conv2d():
  rng = np.random.default_rng(0)
  n = 3
  img = rng.standard_normal((n, 4, 64, 64))
  w1, w2, w3 = rng.standard_normal((32, 4, 8, 8)), rng.standard_normal((64, 32, 4, 4)), rng.standard_normal((64, 64, 3, 3))
  t = lambda a: torch.tensor(a)
  a1 = F.conv2d(t(img), t(w1), stride=4)               # [n,32,15,15]
  a2 = F.conv2d(a1, t(w2), stride=2)                   # [n,64,6,6]
  a3 = F.conv2d(a2, t(w3), stride=1)                   # [n,64,4,4]
  # conv1 = 2x2 stride-1 over the space-to-depth image
  f1, _ = ET._conv_tables(0, 32, 4, 8, 8, 4)
  o1 = _conv_taps(_s2d_image(img), _pack(w1, f1.table), TAPS2, 15, 15, 64)
  np.testing.assert_allclose(o1, a1.numpy().transpose(0, 2, 3, 1), rtol=1e-9, atol=1e-9)
  # conv1's output stored as cells [n,8,8,128] through the position table the plan builds
  oh, ow = np.meshgrid(np.arange(15), np.arange(15), indexing="ij")
  pos = (((oh // 2) * 8 + ow // 2) * 128 + ((oh % 2) * 2 + ow % 2) * 32).ravel()
  cells = np.zeros((n, 8 * 8 * 128))
  for p_, off in enumerate(pos):
    cells[:, off:off + 32] = o1.reshape(n, 225, 32)[:, p_]
  cells = cells.reshape(n, 8, 8, 128)
  # conv2 = 2x2 stride-1 over the cells
  f2, _ = ET._conv_tables(0, 64, 32, 4, 4, 2)
  o2 = _conv_taps(cells, _pack(w2, f2.table), TAPS2, 6, 6, 128)
  np.testing.assert_allclose(o2, a2.numpy().transpose(0, 2, 3, 1), rtol=1e-9, atol=1e-8)
  # conv3 = 3x3 stride-1
  f3, _ = ET._conv_tables(0, 64, 64, 3, 3, 1)
  o3 = _conv_taps(o2, _pack(w3, f3.table), TAPS3, 4, 4, 64)
  np.testing.assert_allclose(o3, a3.numpy().transpose(0, 2, 3, 1), rtol=1e-9, atol=1e-7)


def test_conv_dgrad_tables_reproduce_autograd():
  """data gradient = the same sum-over-taps with NEGATED shifts and the transposed packing
  (engine_tc._trunk_bwd): conv3 and conv2 against torch autograd."""
  rng = np.random.default_rng(1)
  n = 2
  w3 = rng.standard_normal((64, 64, 3, 3)); x = rng.standard_normal((n, 64, 6, 6))
  xt = torch.tensor(x, requires_grad=True)
  g = rng.standard_normal((n, 64, 4, 4))
  F.conv2d(xt, torch.tensor(w3)).backward(torch.tensor(g))
  _, d3 = ET._conv_tables(0, 64, 64, 3, 3, 1)
  wd = _pack(w3, d3.table)                               # [cell channel rows, tap * 64 + n]
  gp = np.zeros((n, 4 + 4, 4 + 4, 64)); gp[:, 2:6, 2:6] = g.transpose(0, 2, 3, 1)   # zero fill = TMA out-of-bounds
  dx = np.zeros((n, 6, 6, 64))
  for t_, (kw, kh) in enumerate(TAPS3):                  # output (h, w) reads dy at (h - kh, w - kw)
    dx += gp[:, 2 - kh:8 - kh, 2 - kw:8 - kw, :] @ wd[:64, t_ * 64:(t_ + 1) * 64].T
  np.testing.assert_allclose(dx, xt.grad.numpy().transpose(0, 2, 3, 1), rtol=1e-9, atol=1e-8)
  # conv2 (4x4 stride 2 on the 15x15x32 map) as 2x2 on cells: gradient w.r.t. the cells
  w2 = rng.standard_normal((64, 32, 4, 4)); a1 = rng.standard_normal((n, 32, 15, 15))
  a1t = torch.tensor(a1, requires_grad=True)
  g2 = rng.standard_normal((n, 64, 6, 6))
  F.conv2d(a1t, torch.tensor(w2), stride=2).backward(torch.tensor(g2))
  _, d2 = ET._conv_tables(0, 64, 32, 4, 4, 2)
  wd2 = _pack(w2, d2.table)                              # [128 cell channels, tap * 64 + n]
  gp2 = np.zeros((n, 9, 9, 64)); gp2[:, 1:7, 1:7] = g2.transpose(0, 2, 3, 1)    # cells -1..7: zero fill outside 0..5
  dcells = np.zeros((n, 8, 8, 128))
  for t_, (dx_, dy_) in enumerate(TAPS2):
    dcells += gp2[:, 1 - dy_:9 - dy_, 1 - dx_:9 - dx_, :] @ wd2[:128, t_ * 64:(t_ + 1) * 64].T
  ref = np.zeros((n, 16, 16, 32)); ref[:, :15, :15] = a1t.grad.numpy().transpose(0, 2, 3, 1)
  ref = ref.reshape(n, 8, 2, 8, 2, 32).transpose(0, 1, 3, 2, 4, 5).reshape(n, 8, 8, 128)   # (cy,cx,(sy,sx,ch))
  # pad positions (row / column 15) are not part of the reference gradient: the kernel masks them with
  # the (zero) activation there
  valid = np.zeros((16, 16), bool); valid[:15, :15] = True
  vmask = valid.reshape(8, 2, 8, 2).transpose(0, 2, 1, 3).reshape(8, 8, 4)
  vmask = np.repeat(vmask, 32, axis=-1)
  np.testing.assert_allclose(dcells * vmask, ref, rtol=1e-9, atol=1e-8)


def test_linear_tables_and_projector_permutation():
  rng = np.random.default_rng(2)
  W = rng.standard_normal((12, 200))
  f, d = ET._linear_tables(5, 12, 200)
  flat = np.concatenate([np.zeros(5), W.reshape(-1)])
  fp = _pack(flat, f.table); dp = _pack(flat, d.table)
  assert fp.shape == (16, 256) and dp.shape[1] == 64
  np.testing.assert_array_equal(fp[:12, :200], W); assert not fp[12:].any() and not fp[:, 200:].any()
  np.testing.assert_array_equal(dp[:200, :12], W.T)
  # NatureCNN projector: torch flattens [64,4,4] as (c, p); the tier's activation is [p, c]
  Wp = rng.standard_normal((8, 1024))
  pp, cc = np.meshgrid(np.arange(16), np.arange(64), indexing="ij")
  f, _ = ET._linear_tables(0, 8, 1024, kperm=(cc * 16 + pp).ravel())
  a3 = rng.standard_normal((64, 4, 4))                   # c, h, w
  ours = _pack(Wp, f.table)[:8] @ a3.reshape(64, 16).T.reshape(-1)      # activation laid out (p, c)
  np.testing.assert_allclose(ours, Wp @ a3.reshape(-1), rtol=1e-12)
