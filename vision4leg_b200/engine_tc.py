"""Tensor-core tier of the PPO networks: fp16 activations, fp32 master weights / gradients,
every GEMM-shaped layer on tcgen05 (v4l_tc_gemm / v4l_tc_wgrad), fed by TMA.

Layouts (all fp16 unless noted)
  image      [N,16,16,64]  4x4 space-to-depth of the 4x64x64 depth stack (v4l_ingest_img): conv1
                           (8x8 stride 4) is a 2x2 stride-1 conv over 64-channel pixels
  a1 cells   [B,8,8,128]   conv1 output stored directly as the 2x2 space-to-depth "cells" of the
                           (zero-padded to 16x16) 15x15x32 map: conv2 (4x4 stride 2) is again a
                           2x2 stride-1 conv; pad cells stay zero (allocated zeroed, never written)
  a2 / a3    [B,6,6,64] / [B,4,4,64] = [B,16,64]
  tokens     [B,17,64]     slot 0 = proprio token, 1..16 = depth tokens (reference base.py:602-622)
Every conv/linear is "sum over taps of a shifted TMA box x packed weight slice"; the data
gradient is the same kernel with negated shifts and the transposed packing, the weight gradient
reads the same boxes as MN-major operands.  Weights are re-packed (fp32 -> fp16, tap-major) from
the flat parameter bucket by ONE gather kernel per optimiser step through a precomputed index
table; the same table scatters the fp32 weight gradients back into the reference layout.

Reference semantics: torchrl/networks/nets.py:909-1038 + base.py:497-626 (LocoTransformer).
"""
import numpy as np
import torch

from . import engine
from .engine import RM, RELU, ACCUM
from ._lib import V4LError

F16 = torch.float16
LOSS_SCALE_OVERRIDE = None        # tests/experiments: force the static loss scale


def _ceil(a, b):
  return (a + b - 1) // b * b


class _Packed:
  """One packed weight matrix: where it lives in the packed buffer and its gather table."""

  def __init__(self, rows, cols, table):
    self.rows, self.cols, self.table = rows, cols, table     # table: int64 [rows, cols], -1 = zero
    self.off = None


def _linear_tables(off, N, K, kperm=None):
  """W[N,K] fp32 at flat offset `off` -> (fwd [N_pad, ceil64(K)], dgrad [Kd_pad, ceil64(N)]).
  kperm[kp] = reference K index held at packed position kp (identity if None)."""
  Np, Kp = _ceil(N, 16), _ceil(K, 64)
  fwd = -np.ones((Np, Kp), np.int64)
  n, k = np.meshgrid(np.arange(N), np.arange(K), indexing="ij")
  if kperm is not None:
    k = np.asarray(kperm)[k]
  fwd[:N, :K] = off + n * K + k
  Kd = _ceil(K, 16)
  if Kd > 256:
    Kd = _ceil(Kd, 256)
  dg = -np.ones((Kd, _ceil(N, 64)), np.int64)
  dg[:K, :N] = (off + n * K + k).T
  return _Packed(Np, Kp, fwd), _Packed(Kd, _ceil(N, 64), dg)


def _conv_tables(off, N, C, KH, KW, s):
  """OIHW conv weight with stride s -> taps of the s x s space-to-depth form.
  packed k = ((dy*T + dx) * (s*s*C)) + (py*s + px)*C + c, kh = s*dy + py, kw = s*dx + px, T = KH//s."""
  T = KH // s
  Cc = s * s * C
  n, dy, dx, py, px, c = np.meshgrid(np.arange(N), np.arange(T), np.arange(T), np.arange(s), np.arange(s),
                                     np.arange(C), indexing="ij")
  src = off + ((n * C + c) * KH + (s * dy + py)) * KW + (s * dx + px)
  fwd = -np.ones((_ceil(N, 16), T * T * Cc), np.int64)
  fwd[:N] = src.reshape(N, T * T * Cc)
  # dgrad: rows = cell channel (py,px,c), cols = (tap, n)
  dg = -np.ones((_ceil(Cc, 16), T * T * _ceil(N, 64)), np.int64)
  d = src.transpose(3, 4, 5, 1, 2, 0).reshape(Cc, T * T, N)     # [(py,px,c), tap, n]
  dgv = dg.reshape(_ceil(Cc, 16), T * T, _ceil(N, 64))
  dgv[:Cc, :, :N] = d
  return _Packed(_ceil(N, 16), T * T * Cc, fwd), _Packed(_ceil(Cc, 16), T * T * _ceil(N, 64), dg)


# LocoTransformer encoder layers (1 head, d=64) run as one fused kernel (csrc/tc_block.cu);
# False = the unfused GEMM / attention / LayerNorm launches (kept for multi-head configurations).
FUSED_LAYER = True
N_WGRAD_STREAMS = 4
FLAT_CONV1 = True
# 3-layer heads / proprio MLP and their data-gradient chains as ONE launch (csrc/tc_mlp.cu).  Correct (tests) but
# NOT faster at minibatch 1024, so off: 8 row tiles = 8 CTAs each streaming all 200 KB of weights (17-24 us per chain)
# against three launches that slice N over 32 CTAs (5-6 us each); profiles/r2_mlp_chain_note.txt
MLP_CHAIN = False
CONV1_WGRAD_S2D = True            # conv1 weight gradient by the single-load kernel (False: generic v4l_tc_wgrad)


class TcWeights:
  """Packed fp16 copies (forward and data-gradient orientations) of one network's GEMM weights."""

  def __init__(self, ops, layout, with_dgrad=True, flatten_names=("encoder.visual_projector.projection.0.weight",)):
    """layout: {param name: (flat offset, shape)} relative to the flat fp32 bucket handed to pack();
    flatten_names: Linear layers fed by torch's (c, p) flatten of the [64,4,4] conv output"""
    self.ops = ops
    self.layout = layout
    self.fwd, self.dgr = {}, {}
    tables = []
    cursor = 0
    for name, (off, shape) in layout.items():
      if not name.endswith("weight") or len(shape) < 2 or ".norm" in name:
        continue
      if len(shape) == 4 and shape[2] > 1:
        stride = {8: 4, 4: 2, 3: 1}[shape[2]]
        f, d = _conv_tables(off, shape[0], shape[1], shape[2], shape[3], stride)
      elif name in flatten_names:
        # torch flattens [64,4,4] as (c, p); our a3 is [16 positions, 64 channels] = (p, c)
        pp, cc = np.meshgrid(np.arange(16), np.arange(64), indexing="ij")
        f, d = _linear_tables(off, shape[0], shape[1], kperm=(cc * 16 + pp).ravel())
      else:
        f, d = _linear_tables(off, shape[0], int(np.prod(shape[1:])))
      for store, pk in ((self.fwd, f), (self.dgr, d)):
        if store is self.dgr and not with_dgrad:
          continue
        pk.off = cursor
        cursor += pk.rows * pk.cols
        cursor = _ceil(cursor, 64)                # 128-byte alignment of every packed matrix (TMA)
        tables.append((pk.off, pk.table))
        store[name] = pk
    self.size = cursor
    table = -np.ones(cursor, np.int64)
    for o, t in tables:
      table[o:o + t.size] = t.ravel()
    dev = ops.device
    self.table = torch.tensor(table.astype(np.int32), device=dev)
    self.packed = torch.zeros(cursor, device=dev, dtype=F16)
    for pk in list(self.fwd.values()) + list(self.dgr.values()):
      pk.dev_table = self.table[pk.off:pk.off + pk.rows * pk.cols]
      pk.w = self.packed[pk.off:pk.off + pk.rows * pk.cols]

  def pack(self, flat):
    self.ops.pack_f16(flat, self.table, self.packed, self.size)


class _PlanTC:
  """Shared machinery: workspace, packed weights, Linear helpers and the NatureCNN trunk."""

  def __init__(self, ops, S, out_dim, layout, with_backward, head_prefix, flatten_names=None):
    self.ops, self.device = ops, ops.device
    self.S, self.Sp = S, _ceil(S, 64)
    self.out_dim = out_dim
    self.layout = layout
    self.W = (TcWeights(ops, layout, with_dgrad=with_backward, flatten_names=flatten_names) if flatten_names
              else TcWeights(ops, layout, with_dgrad=with_backward))
    self._ws = {}
    self.world = 1                    # data-parallel world size (loss scale, see _begin_backward)
    self._rr, self._forked = 0, set()
    self.k_base = [k for k in layout if k.startswith("encoder.base.seq_fcs.") and k.endswith("weight")]
    self.k_head = sorted((k for k in layout if k.startswith(head_prefix) and k.endswith("weight")),
                         key=lambda k: int(k.split(".")[-2]))
    oh, ow = np.meshgrid(np.arange(15), np.arange(15), indexing="ij")
    pos = ((oh // 2) * 8 + ow // 2) * 128 + ((oh % 2) * 2 + ow % 2) * 32
    self.pos_a1 = torch.tensor(pos.ravel().astype(np.int32), device=self.device)
    self.taps2 = [(dx, dy) for dy in range(2) for dx in range(2)]
    self.taps3 = [(kw, kh) for kh in range(3) for kw in range(3)]

  # ---- workspace ------------------------------------------------------------------------------
  def buf(self, name, shape, dtype=F16, zero=False):
    key = (name,) + tuple(shape)
    t = self._ws.get(key)
    if t is None:
      t = (torch.zeros if zero else torch.empty)(shape, device=self.device, dtype=dtype)
      self._ws[key] = t
    return t

  def pack(self, flat):
    self.W.pack(flat)

  def _view(self, flat, name):
    off, shape = self.layout[name]
    return flat[off:off + int(np.prod(shape))]

  # ---- generic layer helpers ------------------------------------------------------------------
  def _lin_fwd(self, flat, wname, x, M, K, out, out_map, relu, c_f32=False):
    pk = self.W.fwd[wname]
    N = self.layout[wname][1][0]
    bias = self._view(flat, wname[:-6] + "bias")
    self.ops.tc_gemm(x, (M, 1, 1, K), (M, 1, 1), (1, 1, 128), [(0, 0)], pk.cols // 64, pk.w, pk.rows, N, bias,
                     out, out_map, c_f32=c_f32, flags=RELU if relu else 0)

  def _chain(self, flat, x, M, x_cols, x_ld, specs, dgrad=False):
    """Linear layers `specs` = [(weight name, relu, out, out_map, out_f32, mask)] chained in one launch.
    dgrad: use the data-gradient (transposed) packing and no bias."""
    layers = []
    for wname, relu, out, out_map, f32, mask in specs:
      pk = (self.W.dgr if dgrad else self.W.fwd)[wname]
      shape = self.layout[wname][1]
      N_out, K_in = (int(np.prod(shape[1:])), shape[0]) if dgrad else (shape[0], int(np.prod(shape[1:])))
      layers.append(dict(w=pk.w, K=pk.cols, N_pad=pk.rows, N_valid=N_out, K_true=K_in,
                         bias=None if dgrad else self._view(flat, wname[:-6] + "bias"), relu=relu,
                         mask=mask, mask_ld=(mask.shape[-1] if mask is not None else 0),
                         out=out, out_f32=f32, out_map=out_map))
    self.ops.tc_mlp_chain(x, M, x_cols, x_ld, layers)

  def _side(self, fn, which=None):
    """Run `fn` on a side stream, ordered after what has been issued so far.  Weight-gradient
    launches only READ activations and gradients that are never overwritten during this backward,
    so they leave the critical path; they are small grids (8-96 CTAs), hence round-robin over
    N_WGRAD_STREAMS streams so that several run at once.  `which` pins a branch to one stream."""
    if which is None:
      which = "w%d" % (self._rr % N_WGRAD_STREAMS)
      self._rr += 1
    if not self._forked:
      self._flushes_at_fork = self.ops.lib.v4l_ctx_early_flushes(self.ops.h)
    self._forked.add(which)
    with self.ops.fork(which):
      fn()

  def _join_all(self):
    for w in sorted(self._forked, key=str):
      self.ops.join(w)
    # A deferred weight gradient that found the scratch full reduced the pending ones early on ITS stream, which is
    # not ordered after the other side streams' launches: refuse to go on rather than race (the default scratch
    # holds a whole backward pass of every shipped network at any minibatch, so this means a custom, smaller one).
    overflowed = len(self._forked) > 1 and self.ops.lib.v4l_ctx_early_flushes(self.ops.h) != self._flushes_at_fork
    if overflowed:
      self._forked.clear()
      raise V4LError("weight-gradient scratch overflowed while launches were spread over side streams: "
                     "create the context with a larger scratch (v4l_ctx_create scratch_bytes)")
    self._forked.clear()

  def _lin_bwd(self, gflat, wname, x, x_cols, dy, dy_cols, M, dx=None, dx_map=None, mask=None, res=None,
               need_dx=True, dy_pitch=0, dy_off=0, side=True):
    """dW, db from (x [M,x_cols], dy [M,dy_cols]) on the side stream; optionally
    dx = (dy @ W) * (mask > 0) + res on the main stream.  dy may be a column window of a wider
    matrix (row pitch dy_pitch elements, first column dy_off)."""
    pk = self.W.fwd[wname]
    N, K = self.layout[wname][1][0], int(np.prod(self.layout[wname][1][1:]))
    inv = self._inv_scale
    st = (dy_pitch, dy_pitch, dy_pitch) if dy_pitch else None
    wg = lambda: self.ops.tc_wgrad(
      x, (M, 1, 1, x_cols), dy, dy_cols, (M, 1, 1), (1, 1, 128), [(0, 0)], N, pk.dev_table, gflat,
      out_scale=inv, dbias=self._view(gflat, wname[:-6] + "bias"), defer=True, dy_strides=st, dy_off=dy_off)
    if side:
      self._side(wg)
    else:
      wg()
    if need_dx:
      pd = self.W.dgr[wname]
      self.ops.tc_gemm(dy, (M, 1, 1, dy_cols), (M, 1, 1), (1, 1, 128), [(0, 0)], pd.cols // 64, pd.w, pd.rows, K,
                       None, dx, dx_map, mask=mask, res=res, a_strides=st, a_off=dy_off)

  # ---- NatureCNN trunk (reference base.py:304-342) on tap-shifted TMA boxes -----------------------
  def _trunk_fwd(self, flat, imgs, idx, B, pre):
    ops = self.ops
    Nimg = imgs.shape[0]
    # conv1 (8x8/4 = 2x2/1 on the s2d image) -> a1 stored as the cells conv2 reads
    a1c = self.buf("a1c", (B, 8, 8, 128), zero=True)
    pk = self.W.fwd[pre + "0.weight"]
    if FLAT_CONV1:
      # single-load form: each 128-position tile of the gathered image is loaded once, the 4 taps are
      # UMMA descriptors shifted by dh*16+dw rows (csrc/tc_conv.cu; bit-identical to the tap boxes, ~10 %
      # faster, 4x less L2->SM traffic); MMAs of 2 tiles interleaved
      ops.tc_conv_flat(imgs, 64, 256, 16, 15, 15, self.taps2, pk.w, pk.rows, 32, self._view(flat, pre + "0.bias"), a1c,
                       RM(225, 8 * 8 * 128, 0, 0, pos_off=self.pos_a1), B, x_idx=idx, flags=RELU, mode=1 + (2 << 4))
    else:
      ops.tc_gemm(imgs, (Nimg, 16, 16, 64), (B, 15, 15), (15, 8, 1), self.taps2, 1, pk.w, pk.rows, 32,
                  self._view(flat, pre + "0.bias"), a1c, RM(225, 8 * 8 * 128, 0, 0, pos_off=self.pos_a1),
                  flags=RELU, a_idx=idx)
    a2 = self.buf("a2", (B, 6, 6, 64))
    pk = self.W.fwd[pre + "2.weight"]
    ops.tc_gemm(a1c, (B, 8, 8, 128), (B, 6, 6), (6, 6, 3), self.taps2, 2, pk.w, pk.rows, 64,
                self._view(flat, pre + "2.bias"), a2, RM(36, 36 * 64, 64, 0), flags=RELU)
    a3 = self.buf("a3", (B, 16, 64))
    pk = self.W.fwd[pre + "4.weight"]
    ops.tc_gemm(a2, (B, 6, 6, 64), (B, 4, 4), (4, 4, 8), self.taps3, 1, pk.w, pk.rows, 64,
                self._view(flat, pre + "4.bias"), a3, RM(16, 16 * 64, 64, 0), flags=RELU)
    return a3

  def _trunk_bwd(self, gflat, da3, B, pre):
    """da3 [B,16,64]: gradient w.r.t. conv3's pre-activation."""
    ops, ws, inv = self.ops, self._ws, self._inv_scale
    a2, a1c = ws[("a2", B, 6, 6, 64)], ws[("a1c", B, 8, 8, 128)]
    self._side(lambda: ops.tc_wgrad(
      a2, (B, 6, 6, 64), da3, 64, (B, 4, 4), (4, 4, 4), self.taps3, 64, self.W.fwd[pre + "4.weight"].dev_table, gflat,
      out_scale=inv, dbias=self._view(gflat, pre + "4.bias"), defer=True))
    da2 = self.buf("da2", (B, 6, 6, 64))
    pd = self.W.dgr[pre + "4.weight"]
    ops.tc_gemm(da3, (B, 4, 4, 64), (B, 6, 6), (6, 6, 3), [(-kw, -kh) for kw, kh in self.taps3], 1, pd.w, pd.rows, 64,
                None, da2, RM(36, 36 * 64, 64, 0), mask=a2)
    self._side(lambda: ops.tc_wgrad(
      a1c, (B, 8, 8, 128), da2, 64, (B, 6, 6), (6, 6, 3), self.taps2, 64, self.W.fwd[pre + "2.weight"].dev_table, gflat,
      out_scale=inv, dbias=self._view(gflat, pre + "2.bias"), defer=True))
    da1c = self.buf("da1c", (B, 8, 8, 128))
    pd = self.W.dgr[pre + "2.weight"]
    ops.tc_gemm(da2, (B, 6, 6, 64), (B, 8, 8), (8, 8, 2), [(-dx_, -dy_) for dx_, dy_ in self.taps2], 1, pd.w, pd.rows, 128,
                None, da1c, RM(64, 64 * 128, 128, 0), mask=a1c)
    if CONV1_WGRAD_S2D:
      # dedicated kernel: each pixel window and each dY cell loaded once per image (csrc/tc_wgrad_s2d.cu);
      # it is the LAST link of the backward chain, so it runs on the main stream with the whole machine
      ops.tc_wgrad_conv1(self._imgs, self._idx, da1c, B, self.W.fwd[pre + "0.weight"].dev_table, gflat,
                         self._view(gflat, pre + "0.bias"), out_scale=inv, defer=True)
      return
    # conv1: per sub-position (py,px) of a cell, X is the stride-2 sub-grid of the s2d image
    subs = [(px, py, (py * 2 + px) * 32) for py in range(2) for px in range(2)]
    self._side(lambda: ops.tc_wgrad(
      self._imgs, (self._imgs.shape[0], 16, 16, 64), da1c, 128, (B, 8, 8), (8, 8, 1), self.taps2, 32,
      self.W.fwd[pre + "0.weight"].dev_table, gflat, x_idx=self._idx, x_estride=2, subs=subs, out_scale=inv,
      dbias=self._view(gflat, pre + "0.bias"), defer=True))

  def loss_scale(self, B):
    """Static loss scale of the fp16 backward: d_out ~ 1/B_global, so scale ~ 4 B_global keeps fp16
    gradients O(1e2); every fp32 result (dW, db, dgamma, dbeta) is multiplied by 1/scale where it is
    produced.  Data parallel: d_out carries 1 / (B * world), so the scale follows the GLOBAL minibatch."""
    if LOSS_SCALE_OVERRIDE:
      return float(LOSS_SCALE_OVERRIDE)
    scale = float(min(32768, max(64, 1 << int(np.floor(np.log2(4 * B * self.world))))))
    return min(scale, 4096.0) if self.world == 1 else scale

  def grad_in(self, B):
    """fp16 [B,16] buffer the backward starts from; the loss kernels write it directly
    (scale_f16 = loss_scale(B)), see backward(..., d_out=None)"""
    return self.buf("dout16", (B, 16))

  def _begin_backward(self, d_out, B):
    """fp32 d_out [B, out_dim] -> loss-scaled fp16 [B,16] (d_out None: grad_in(B) was already written)"""
    scale = self.loss_scale(B)
    self._inv_scale = 1.0 / scale
    g16 = self.grad_in(B)
    if d_out is not None:
      self.ops.gather_rows_f16(d_out, True, None, g16, B, self.out_dim, self.out_dim, 16, scale=scale)
    return g16


class LocoPlanTC(_PlanTC):
  """LocoTransformer forward/backward on the tensor-core tier for one batch size."""
  family = "loco"

  def __init__(self, ops, S, out_dim, layout, n_heads=(1, 1), with_backward=True, has_state=True):
    super().__init__(ops, S, out_dim, layout, with_backward, "visual_seq_append_fcs.")
    self.n_heads = list(n_heads)
    self.has_state = has_state          # False: vision-only Transformer (reference nets.py:784-906): 16 tokens, mean pool
    self.T, self.d = (17 if has_state else 16), 64
    self.first = 1 if has_state else 0  # token slot of the first depth token
    self.pd = 2 * self.d if has_state else self.d
    self._diag = {}

  # ---- forward --------------------------------------------------------------------------------
  def forward(self, flat, imgs, idx, st, B, out, out_map=None, enc_from=None):
    """imgs [N,16,16,64] fp16 (whole rollout), idx int32 [B] or None, st [B,Sp] fp16 proprio rows,
    out fp32 [B,out_dim].  `flat` = the fp32 bucket the layout offsets refer to (biases, LN).
    enc_from: another plan of the same batch whose SHARED-ENCODER output (the 17 / 16 tokens) is reused instead
    of running the conv trunk and the proprio branch again (actor + critic inference on one observation batch)."""
    ops, T, d = self.ops, self.T, self.d
    self._flat, self._B, self._imgs, self._idx, self._st = flat, B, imgs, idx, st
    if enc_from is not None:
      return self._forward_layers(flat, enc_from._ws[("tok0", B, T, d)], B, out, out_map)
    tok = self.buf("tok0", (B, T, d))
    s1 = self.buf("s1", (B, 256)); s2 = self.buf("s2", (B, 256))

    def state_branch():      # proprio MLP -> token 0: independent of the conv trunk
      if MLP_CHAIN:
        self._chain(flat, st, B, self.Sp, self.Sp, [
          (self.k_base[0], True, s1, RM.dense(256), False, None), (self.k_base[1], True, s2, RM.dense(256), False, None),
          ("encoder.state_projector.projection.0.weight", True, tok, RM.slots(1, T, d, 0), False, None)])
        return
      self._lin_fwd(flat, self.k_base[0], st, B, self.Sp, s1, RM.dense(256), True)
      self._lin_fwd(flat, self.k_base[1], s1, B, 256, s2, RM.dense(256), True)
      self._lin_fwd(flat, "encoder.state_projector.projection.0.weight", s2, B, 256, tok, RM.slots(1, T, d, 0), True)
    if self.has_state:
      self._side(state_branch, which=0)
    a3 = self._trunk_fwd(flat, imgs, idx, B, "encoder.depth_visual_base.layers.")
    self._lin_fwd(flat, "encoder.depth_up_conv.weight", a3, B * 16, 64, tok, RM.slots(16, T, d, self.first), False)
    self._join_all()
    return self._forward_layers(flat, tok, B, out, out_map)

  def _forward_layers(self, flat, tok, B, out, out_map):
    """encoder layers, pooling and the head on the token tensor"""
    ops, T, d = self.ops, self.T, self.d
    R = B * T
    x = tok
    self._layers = []
    for l, nh in enumerate(self.n_heads):
      p = "visual_append_layers.%d." % l
      qkv = self.buf("qkv%d" % l, (R, 3 * d))
      o = self.buf("o%d" % l, (R, d))
      pr = self.buf("p%d" % l, (B, nh, T, T), torch.float32)
      if nh == 1 and FUSED_LAYER:          # whole encoder layer in one tcgen05 kernel
        h = self.buf("h%d" % l, (R, d)); f1 = self.buf("f1_%d" % l, (R, 256)); y = self.buf("y%d" % l, (R, d))
        st1 = self.buf("st1_%d" % l, (R, 2), torch.float32); st2 = self.buf("st2_%d" % l, (R, 2), torch.float32)
        xh1 = self.buf("xh1_%d" % l, (R, d)); xh2 = self.buf("xh2_%d" % l, (R, d))
        w = {"w_in": self.W.fwd[p + "self_attn.in_proj_weight"].w, "w_o": self.W.fwd[p + "self_attn.out_proj.weight"].w,
             "w_1": self.W.fwd[p + "linear1.weight"].w, "w_2": self.W.fwd[p + "linear2.weight"].w}
        par = {"b_in": self._view(flat, p + "self_attn.in_proj_bias"), "b_o": self._view(flat, p + "self_attn.out_proj.bias"),
               "g1": self._view(flat, p + "norm1.weight"), "be1": self._view(flat, p + "norm1.bias"),
               "b1": self._view(flat, p + "linear1.bias"), "b2": self._view(flat, p + "linear2.bias"),
               "g2": self._view(flat, p + "norm2.weight"), "be2": self._view(flat, p + "norm2.bias")}
        ops.tc_block_fwd(x, B, T, w, par, dict(qkv=qkv, o=o, h=h, f1=f1, y=y, p=pr, st1=st1, st2=st2, xh1=xh1, xh2=xh2))
        self._layers.append(dict(p=p, nh=nh, x=x, qkv=qkv, o=o, pr=pr, h=h, st1=st1, f1=f1, st2=st2, xh1=xh1, xh2=xh2,
                                 fused=True))
        x = y
        continue
      self._lin_fwd(flat, p + "self_attn.in_proj_weight", x, R, d, qkv, RM.dense(3 * d), False)
      ops.attn_fwd_f16(qkv, o, pr, B, T, d, nh)
      proj = self.buf("proj", (R, d))
      self._lin_fwd(flat, p + "self_attn.out_proj.weight", o, R, d, proj, RM.dense(d), False)
      h = self.buf("h%d" % l, (R, d))
      z1 = self.buf("z1_%d" % l, (R, d), torch.float32); st1 = self.buf("st1_%d" % l, (R, 2), torch.float32)
      ops.ln_fwd_f16(proj, x, self._view(flat, p + "norm1.weight"), self._view(flat, p + "norm1.bias"), h, z1, st1, R, d)
      f1 = self.buf("f1_%d" % l, (R, 256))
      self._lin_fwd(flat, p + "linear1.weight", h, R, d, f1, RM.dense(256), True)
      f2 = self.buf("f2", (R, d))
      self._lin_fwd(flat, p + "linear2.weight", f1, R, 256, f2, RM.dense(d), False)
      y = self.buf("y%d" % l, (R, d))
      z2 = self.buf("z2_%d" % l, (R, d), torch.float32); st2 = self.buf("st2_%d" % l, (R, 2), torch.float32)
      ops.ln_fwd_f16(f2, h, self._view(flat, p + "norm2.weight"), self._view(flat, p + "norm2.bias"), y, z2, st2, R, d)
      self._layers.append(dict(p=p, nh=nh, x=x, qkv=qkv, o=o, pr=pr, h=h, z1=z1, st1=st1, f1=f1, z2=z2, st2=st2))
      x = y
    pooled = self.buf("pooled", (B, self.pd))
    ops.pool_fwd_f16(x, pooled, B, T, d, 0 if self.has_state else 1)
    h1 = self.buf("h1", (B, 256)); h2 = self.buf("h2", (B, 256))
    if MLP_CHAIN:
      self._chain(flat, pooled, B, self.pd, self.pd, [
        (self.k_head[0], True, h1, RM.dense(256), False, None), (self.k_head[1], True, h2, RM.dense(256), False, None),
        (self.k_head[2], False, out, out_map or RM.dense(self.out_dim), True, None)])
      return out
    self._lin_fwd(flat, self.k_head[0], pooled, B, self.pd, h1, RM.dense(256), True)
    self._lin_fwd(flat, self.k_head[1], h1, B, 256, h2, RM.dense(256), True)
    self._lin_fwd(flat, self.k_head[2], h2, B, 256, out, out_map or RM.dense(self.out_dim), False, c_f32=True)
    return out

  # ---- backward -------------------------------------------------------------------------------
  def _diag_table(self, name):
    """index table that scatters the diagonal of a [64,64] xhat^T dy product to the LayerNorm weight
    gradient (LayerNorm affine gradients ride on the weight-gradient GEMM)."""
    t = self._diag.get(name)
    if t is None:
      tab = -np.ones((64, 64), np.int64)
      tab[np.arange(64), np.arange(64)] = self.layout[name][0] + np.arange(64)
      t = self._diag[name] = torch.tensor(tab.ravel().astype(np.int32), device=self.device)
    return t

  def _layer_bwd_fused(self, gflat, l, Ly, dy, B):
    """One encoder layer: data gradients in one kernel (tc_block.cu), weight / bias / LayerNorm
    gradients as side-stream GEMMs over the row gradients that kernel stores."""
    ops, T, d, flat = self.ops, self.T, self.d, self._flat
    R = B * T
    p = Ly["p"]
    inv = self._inv_scale
    g = {k: self.buf("%s_%d" % (k, l), (R, n)) for k, n in (("dz2", d), ("df1", 256), ("dh", d), ("dz1", d),
                                                            ("dqkv", 3 * d), ("dx", d))}
    D = self.W.dgr
    w = {"w2d": D[p + "linear2.weight"].w, "w1d": D[p + "linear1.weight"].w,
         "wod": D[p + "self_attn.out_proj.weight"].w, "wind": D[p + "self_attn.in_proj_weight"].w}
    saved = dict(qkv=Ly["qkv"], xh1=Ly["xh1"], xh2=Ly["xh2"], f1=Ly["f1"], p=Ly["pr"], st1=Ly["st1"], st2=Ly["st2"])
    ops.tc_block_bwd(dy, B, T, saved, w, self._view(flat, p + "norm1.weight"), self._view(flat, p + "norm2.weight"), g)

    self._lin_bwd(gflat, p + "linear2.weight", Ly["f1"], 256, g["dz2"], d, R, need_dx=False)
    self._lin_bwd(gflat, p + "linear1.weight", Ly["h"], d, g["df1"], 256, R, need_dx=False)
    self._lin_bwd(gflat, p + "self_attn.out_proj.weight", Ly["o"], d, g["dz1"], d, R, need_dx=False)
    self._lin_bwd(gflat, p + "self_attn.in_proj_weight", Ly["x"], d, g["dqkv"], 3 * d, R, need_dx=False)
    for norm, xh, dyn in (("norm2", Ly["xh2"], dy), ("norm1", Ly["xh1"], g["dh"])):
      self._side(lambda: ops.tc_wgrad(
        xh, (R, 1, 1, d), dyn, d, (R, 1, 1), (1, 1, 128), [(0, 0)], d, self._diag_table(p + norm + ".weight"),
        gflat, out_scale=inv, dbias=self._view(gflat, p + norm + ".bias"), defer=True,
        algo_flops=4.0 * R * d))        # dgamma = diag(xhat^T dy), dbeta = colsum(dy): 2 x 2 x R x d useful FLOPs
    return g["dx"]

  def backward(self, gflat, d_out, flush=True):
    """d_out fp32 [B,out_dim] (None: the loss kernel already wrote grad_in(B)); writes every
    weight/bias/LayerNorm gradient (fp32) into gflat at the layout offsets.  flush=False leaves the
    split-K partials pending for the fused optimiser tail (v4l_opt_tail)."""
    ops, T, d, B, flat = self.ops, self.T, self.d, self._B, self._flat
    R = B * T
    A = self.out_dim
    ws = self._ws
    g16 = self._begin_backward(d_out, B)
    inv = self._inv_scale
    pd = self.pd
    h1, h2, pooled = ws[("h1", B, 256)], ws[("h2", B, 256)], ws[("pooled", B, pd)]
    dh2 = self.buf("dh2", (B, 256)); dh1 = self.buf("dh1", (B, 256)); dpool = self.buf("dpool", (B, pd))
    if MLP_CHAIN:      # the three data gradients in one launch, then the weight gradients (side streams) from its outputs
      self._chain(flat, g16, B, 16, 16, [
        (self.k_head[2], False, dh2, RM.dense(256), False, h2), (self.k_head[1], False, dh1, RM.dense(256), False, h1),
        (self.k_head[0], False, dpool, RM.dense(pd), False, None)], dgrad=True)
    chain = MLP_CHAIN
    self._lin_bwd(gflat, self.k_head[2], h2, 256, g16, 16, B, dh2, RM.dense(256), mask=h2, need_dx=not chain)
    self._lin_bwd(gflat, self.k_head[1], h1, 256, dh2, 256, B, dh1, RM.dense(256), mask=h1, need_dx=not chain)
    self._lin_bwd(gflat, self.k_head[0], pooled, pd, dh1, 256, B, dpool, RM.dense(pd), need_dx=not chain)
    dx = self.buf("dx_top", (R, d))
    ops.pool_bwd_f16(dpool, dx, B, T, d, 0 if self.has_state else 1)
    # every gradient tensor below is written once and then only read (no in-place accumulation,
    # per-layer buffers): the weight-gradient launches on the side stream can lag behind safely
    for l in reversed(range(len(self._layers))):
      Ly = self._layers[l]
      p = Ly["p"]
      if Ly.get("fused"):
        dx = self._layer_bwd_fused(gflat, l, Ly, dx, B)
        continue
      dz2 = self.buf("dz2_%d" % l, (R, d))
      ops.ln_bwd_f16(dx, Ly["z2"], Ly["st2"], self._view(flat, p + "norm2.weight"), dz2,
                     self._view(gflat, p + "norm2.weight"), self._view(gflat, p + "norm2.bias"), R, d, out_scale=inv)
      df1 = self.buf("df1_%d" % l, (R, 256))
      self._lin_bwd(gflat, p + "linear2.weight", Ly["f1"], 256, dz2, d, R, df1, RM.dense(256), mask=Ly["f1"])
      dh = self.buf("dh_%d" % l, (R, d))
      self._lin_bwd(gflat, p + "linear1.weight", Ly["h"], d, df1, 256, R, dh, RM.dense(d), res=dz2)
      dz1 = self.buf("dz1_%d" % l, (R, d))
      ops.ln_bwd_f16(dh, Ly["z1"], Ly["st1"], self._view(flat, p + "norm1.weight"), dz1,
                     self._view(gflat, p + "norm1.weight"), self._view(gflat, p + "norm1.bias"), R, d, out_scale=inv)
      do = self.buf("do", (R, d))
      self._lin_bwd(gflat, p + "self_attn.out_proj.weight", Ly["o"], d, dz1, d, R, do, RM.dense(d))
      dqkv = self.buf("dqkv_%d" % l, (R, 3 * d))
      ops.attn_bwd_f16(Ly["qkv"], Ly["pr"], do, dqkv, B, T, d, Ly["nh"])
      dxn = self.buf("dx_%d" % l, (R, d))
      self._lin_bwd(gflat, p + "self_attn.in_proj_weight", Ly["x"], d, dqkv, 3 * d, R, dxn, RM.dense(d), res=dz1)
      dx = dxn
    tok = ws[("tok0", B, T, d)]
    # proprio token -> state MLP
    ds = self.buf("ds", (B, d))
    smap = RM.slots(1, T, d, 0)
    s1, s2 = ws[("s1", B, 256)], ws[("s2", B, 256)]
    ds2 = self.buf("ds2", (B, 256)); ds1 = self.buf("ds1", (B, 256))

    def state_branch():      # whole proprio-branch backward off the critical path
      ops.relu_bwd_f16(dx, smap, tok, smap, ds, RM.dense(d), B, d)
      proj = "encoder.state_projector.projection.0.weight"
      if MLP_CHAIN:
        self._chain(flat, ds, B, d, d, [(proj, False, ds2, RM.dense(256), False, s2),
                                        (self.k_base[1], False, ds1, RM.dense(256), False, s1)], dgrad=True)
      self._lin_bwd(gflat, proj, s2, 256, ds, d, B, ds2, RM.dense(256), mask=s2, need_dx=not MLP_CHAIN)
      self._lin_bwd(gflat, self.k_base[1], s1, 256, ds2, 256, B, ds1, RM.dense(256), mask=s1, need_dx=not MLP_CHAIN)
      self._lin_bwd(gflat, self.k_base[0], self._st, self.Sp, ds1, 256, B, need_dx=False)
    if self.has_state:
      self._side(state_branch, which=0)
    # depth tokens -> 1x1 up-conv (dY is the strided [B,16,64] window of the token gradient)
    a3 = ws[("a3", B, 16, 64)]
    up = "encoder.depth_up_conv.weight"
    strides = (d, T * d, T * d)
    off = self.first * d
    self._side(lambda: ops.tc_wgrad(
      a3, (B, 1, 16, 64), dx, d, (B, 1, 16), (16, 1, 8), [(0, 0)], 64, self.W.fwd[up].dev_table, gflat,
      dy_strides=strides, dy_off=off, out_scale=inv, dbias=self._view(gflat, "encoder.depth_up_conv.bias"), defer=True))
    da3 = self.buf("da3", (B, 16, 64))
    pdw = self.W.dgr[up]
    ops.tc_gemm(dx, (B, 1, 16, 64), (B, 1, 16), (16, 1, 8), [(0, 0)], 1, pdw.w, pdw.rows, 64, None, da3,
                RM(16, 16 * 64, 64, 0), mask=a3, a_strides=strides, a_off=off)
    self._trunk_bwd(gflat, da3, B, "encoder.depth_visual_base.layers.")
    self._join_all()
    if flush:
      ops.tc_wgrad_flush()


class NaturePlanTC(_PlanTC):
  """NatureCNN + concat MLP (reference nets.py:194-262, base.py:345-385) on the tensor-core tier."""
  family = "nature"

  def __init__(self, ops, S, out_dim, layout, n_heads=None, with_backward=True):
    super().__init__(ops, S, out_dim, layout, with_backward, "seq_append_fcs.")
    self.k_proj = "encoder.visual_projector.projection.0.weight"
    self.vd = layout[self.k_proj][1][0]
    self.sd = layout[self.k_base[-1]][1][0]

  def forward(self, flat, imgs, idx, st, B, out, out_map=None, enc_from=None):
    self._flat, self._B, self._imgs, self._idx, self._st = flat, B, imgs, idx, st
    W = self.vd + self.sd
    if enc_from is not None:            # shared encoder output [visual | proprio] of another plan of this batch
      cat = enc_from._ws[("cat", B, W)]
    else:
      a3 = self._trunk_fwd(flat, imgs, idx, B, "encoder.visual_base.layers.")
      cat = self.buf("cat", (B, W))
      # flatten + Linear(1024, vd) + ReLU: the (c,p) -> (p,c) reorder lives in the packing table
      self._lin_fwd(flat, self.k_proj, a3, B, 1024, cat, RM(1, W, 0, 0), True)
      s1 = self.buf("s1", (B, 256))
      self._lin_fwd(flat, self.k_base[0], st, B, self.Sp, s1, RM.dense(256), True)
      self._lin_fwd(flat, self.k_base[1], s1, B, 256, cat, RM(1, W, 0, self.vd), True)
    h1 = self.buf("h1", (B, 256)); h2 = self.buf("h2", (B, 256))
    if MLP_CHAIN:
      self._chain(flat, cat, B, W, W, [
        (self.k_head[0], True, h1, RM.dense(256), False, None), (self.k_head[1], True, h2, RM.dense(256), False, None),
        (self.k_head[2], False, out, out_map or RM.dense(self.out_dim), True, None)])
      return out
    self._lin_fwd(flat, self.k_head[0], cat, B, W, h1, RM.dense(256), True)
    self._lin_fwd(flat, self.k_head[1], h1, B, 256, h2, RM.dense(256), True)
    self._lin_fwd(flat, self.k_head[2], h2, B, 256, out, out_map or RM.dense(self.out_dim), False, c_f32=True)
    return out

  def backward(self, gflat, d_out, flush=True):
    ops, B, ws = self.ops, self._B, self._ws
    W = self.vd + self.sd
    g16 = self._begin_backward(d_out, B)
    cat, h1, h2, s1, a3 = ws[("cat", B, W)], ws[("h1", B, 256)], ws[("h2", B, 256)], ws[("s1", B, 256)], ws[("a3", B, 16, 64)]
    dh2 = self.buf("dh2", (B, 256)); dh1 = self.buf("dh1", (B, 256)); dcat = self.buf("dcat", (B, W))
    if MLP_CHAIN:      # the first two data gradients in one launch (the third is 512 wide: its own GEMM)
      self._chain(self._flat, g16, B, 16, 16, [(self.k_head[2], False, dh2, RM.dense(256), False, h2),
                                               (self.k_head[1], False, dh1, RM.dense(256), False, h1)], dgrad=True)
    self._lin_bwd(gflat, self.k_head[2], h2, 256, g16, 16, B, dh2, RM.dense(256), mask=h2, need_dx=not MLP_CHAIN)
    self._lin_bwd(gflat, self.k_head[1], h1, 256, dh2, 256, B, dh1, RM.dense(256), mask=h1, need_dx=not MLP_CHAIN)
    self._lin_bwd(gflat, self.k_head[0], cat, W, dh1, 256, B, dcat, RM.dense(W), mask=cat)
    # proprio MLP: dy = dcat[:, vd:]
    ds1 = self.buf("ds1", (B, 256))
    self._lin_bwd(gflat, self.k_base[1], s1, 256, dcat, self.sd, B, ds1, RM.dense(256), mask=s1, dy_pitch=W, dy_off=self.vd)
    self._lin_bwd(gflat, self.k_base[0], self._st, self.Sp, ds1, 256, B, need_dx=False)
    # projector: dy = dcat[:, :vd]; da3 comes out in our (p,c) order through the packing table
    da3 = self.buf("da3", (B, 16, 64))
    self._lin_bwd(gflat, self.k_proj, a3, 1024, dcat, self.vd, B, da3, RM.dense(1024), mask=a3, dy_pitch=W, dy_off=0)
    self._trunk_bwd(gflat, da3, B, "encoder.visual_base.layers.")
    self._join_all()
    if flush:
      ops.tc_wgrad_flush()


class VitPlanTC(LocoPlanTC):
  """Vision-only Transformer policy / value net (reference nets.py:784-906 + base.py:388-494): the LocoTransformer
  plan without the proprio token — 16 depth tokens, mean pooling, 64-wide head input."""
  family = "vit"

  def __init__(self, ops, S, out_dim, layout, n_heads=(1, 1), with_backward=True):
    super().__init__(ops, 0, out_dim, layout, n_heads, with_backward, has_state=False)


class NatureVOPlanTC(_PlanTC):
  """Vision-only NatureCNN (reference nets.py:133-191 with a flattening NatureEncoder, base.py:334-342):
  conv trunk -> flatten -> 3-layer head; torch's (c, p) flatten order lives in the first head layer's packing table."""
  family = "nvo"

  def __init__(self, ops, S, out_dim, layout, n_heads=None, with_backward=True):
    super().__init__(ops, 0, out_dim, layout, with_backward, "seq_append_fcs.", flatten_names=("seq_append_fcs.0.weight",))

  def forward(self, flat, imgs, idx, st, B, out, out_map=None, enc_from=None):
    self._flat, self._B, self._imgs, self._idx, self._st = flat, B, imgs, idx, st
    a3 = enc_from._ws[("a3", B, 16, 64)] if enc_from is not None else self._trunk_fwd(flat, imgs, idx, B, "encoder.layers.")
    h1 = self.buf("h1", (B, 256)); h2 = self.buf("h2", (B, 256))
    if MLP_CHAIN:
      self._chain(flat, a3, B, 1024, 1024, [
        (self.k_head[0], True, h1, RM.dense(256), False, None), (self.k_head[1], True, h2, RM.dense(256), False, None),
        (self.k_head[2], False, out, out_map or RM.dense(self.out_dim), True, None)])
      return out
    self._lin_fwd(flat, self.k_head[0], a3, B, 1024, h1, RM.dense(256), True)
    self._lin_fwd(flat, self.k_head[1], h1, B, 256, h2, RM.dense(256), True)
    self._lin_fwd(flat, self.k_head[2], h2, B, 256, out, out_map or RM.dense(self.out_dim), False, c_f32=True)
    return out

  def backward(self, gflat, d_out, flush=True):
    ops, B, ws = self.ops, self._B, self._ws
    g16 = self._begin_backward(d_out, B)
    h1, h2, a3 = ws[("h1", B, 256)], ws[("h2", B, 256)], ws[("a3", B, 16, 64)]
    dh2 = self.buf("dh2", (B, 256)); dh1 = self.buf("dh1", (B, 256)); da3 = self.buf("da3", (B, 16, 64))
    if MLP_CHAIN:
      self._chain(self._flat, g16, B, 16, 16, [(self.k_head[2], False, dh2, RM.dense(256), False, h2),
                                               (self.k_head[1], False, dh1, RM.dense(256), False, h1)], dgrad=True)
    self._lin_bwd(gflat, self.k_head[2], h2, 256, g16, 16, B, dh2, RM.dense(256), mask=h2, need_dx=not MLP_CHAIN)
    self._lin_bwd(gflat, self.k_head[1], h1, 256, dh2, 256, B, dh1, RM.dense(256), mask=h1, need_dx=not MLP_CHAIN)
    # first head layer: da3 comes out in our (p, c) order through the packing table, masked by the trunk's ReLU
    self._lin_bwd(gflat, self.k_head[0], a3, 1024, dh1, 256, B, da3, RM.dense(1024), mask=a3)
    self._trunk_bwd(gflat, da3, B, "encoder.layers.")
    self._join_all()
    if flush:
      ops.tc_wgrad_flush()


PLANS = {"loco": LocoPlanTC, "nature": NaturePlanTC, "vit": VitPlanTC, "nvo": NatureVOPlanTC}
