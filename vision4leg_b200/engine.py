"""Host-side orchestration of the CUDA path: layer plans for the three policy families.

A *plan* knows, for one network (policy mean head or value head) and one batch size, which
C-ABI calls (include/v4l_b200.h) make up its forward and backward pass, and owns the device
workspace for the saved activations.  Parameters are passed in as {state_dict key: fp32 CUDA
tensor}; gradients are written to a parallel dict.  Nothing here computes on the CPU and
nothing calls a torch math op on the hot path (torch supplies memory and streams only).

Reference semantics reproduced (paths relative to the reference root):
  MLPPlan     torchrl/networks/nets.py:16-55 (Net) + base.py:8-44 (MLPBase)
  NaturePlan  nets.py:194-262 (ImpalaEncoderProjNet) + base.py:345-385 (NatureFuseEncoder)
  LocoPlan    nets.py:909-1038 (LocoTransformer) + base.py:497-626 (LocoTransformerEncoder)
  VisionTransformerPlan  nets.py:784-906 (Transformer) + base.py:388-494 (TransformerEncoder)
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import RowMap, GemmArgs, WgradArgs, TcGemmArgs, TcWgradArgs, RELU, ACCUM, check, ptr

IMG_C, IMG_H, IMG_W = 4, 64, 64
IMG_ELEMS = IMG_C * IMG_H * IMG_W


class RM:
  """Python mirror of v4l_rowmap (keeps the index tensors alive)."""
  __slots__ = ("P", "item_stride", "pos_stride", "base", "idx", "pos_off")

  def __init__(self, P=1, item_stride=0, pos_stride=0, base=0, idx=None, pos_off=None):
    self.P, self.item_stride, self.pos_stride, self.base = P, item_stride, pos_stride, base
    self.idx, self.pos_off = idx, pos_off

  def c(self):
    return RowMap(self.P, self.item_stride, self.pos_stride, self.base, ptr(self.idx),
                  ptr(self.pos_off))

  @staticmethod
  def dense(ld):
    return RM(1, ld, 0, 0)

  @staticmethod
  def slots(P, slot_count, d, first):
    """rows (b, p<P) -> token slot (first + p) of a [B, slot_count, d] tensor."""
    return RM(P, slot_count * d, d, first * d)


class Input:
  """Where a batch's observations live: proprio rows and CHW image planes, optionally
  through a row-index gather (minibatch rows of the device-resident rollout)."""
  __slots__ = ("state", "state_stride", "state_base", "img", "img_stride", "img_base", "idx", "B")

  def __init__(self, B, state=None, state_stride=0, state_base=0, img=None, img_stride=0,
               img_base=0, idx=None):
    self.B = B
    self.state, self.state_stride, self.state_base = state, state_stride, state_base
    self.img, self.img_stride, self.img_base = img, img_stride, img_base
    self.idx = idx

  @staticmethod
  def from_flat(x, S, has_img=True):
    """x [B, S (+16384)] fp32 contiguous CUDA (the layout the reference modules receive,
    nets.py:997-1000)."""
    B, D = x.shape
    if has_img:
      return Input(B, x, D, 0, x, D, S)
    return Input(B, x, D, 0)


class Ops:
  """Typed wrappers over the C ABI for one device."""

  def __init__(self, device, ctx=None):
    self.ctx = ctx if ctx is not None else _lib.ctx(device)     # ctx: a private _lib.Context (own scratch size)
    self.lib = self.ctx.lib
    self.h = self.ctx.handle
    self.device = self.ctx.device
    self.launches = 0     # kernel launches issued through this object (bench `gpu_launches`)
    self._rec = None      # when a list: (entry point, argument struct copy, FLOPs) of the tensor-core launches,
                          # so that a profiler / bench.py can replay a kernel family on its own (record())

  def record(self, on=True):
    """start (returns nothing) / stop (returns the list) recording tensor-core launches for replay()"""
    if on:
      self._rec = []
      return None
    rec, self._rec = self._rec, None
    return rec

  def _note(self, fn, g, flops, keep=()):
    if self._rec is not None:
      self._rec.append((fn, type(g).from_buffer_copy(g), float(flops), keep))

  def replay(self, rec):
    """re-issue recorded launches on the current stream (same arguments, same buffers)"""
    for fn, g, _, _ in rec:
      if isinstance(g, tuple):
        check(getattr(self.lib, fn)(self.h, self.ctx.stream(), *g))
      else:
        check(getattr(self.lib, fn)(self.h, self.ctx.stream(), C.byref(g)))

  # ---- side stream: independent work (weight gradients, frozen-target forward) leaves the
  #      critical path; under CUDA-graph capture this becomes a parallel branch of the graph
  def _side_stream(self, which):
    sides = self.__dict__.setdefault("_sides", {})
    if which not in sides:
      sides[which] = torch.cuda.Stream(device=self.device)
    return sides[which]

  def fork(self, which=0, wait=True):
    """Context manager: launches inside go to side stream `which`, ordered after everything
    issued so far on the current stream (wait=False: free-running, e.g. the copy stream)."""
    side = self._side_stream(which)
    cur = torch.cuda.current_stream(self.device)
    if wait and cur != side:
      side.wait_stream(cur)
    return torch.cuda.stream(side)

  def join(self, which=0):
    side = self._side_stream(which)
    cur = torch.cuda.current_stream(self.device)
    if cur != side:
      cur.wait_stream(side)

  # ---- GEMMs
  def gemm(self, a, a_map, koff, b, b_sk, b_sn, bias, c, c_map, M, N, K, flags=0, mask=None,
           mask_map=None, c_koff=None):
    g = GemmArgs(ptr(a), a_map.c(), ptr(koff), ptr(b), b_sk, b_sn, ptr(bias), ptr(c), c_map.c(),
                 ptr(c_koff), ptr(mask), (mask_map or RM()).c(), M, N, K, flags)
    check(self.lib.v4l_gemm_rows(self.h, self.ctx.stream(), C.byref(g)))
    self.launches += 1

  def linear_fwd(self, a, a_map, koff, w, bias, c, c_map, M, N, K, relu):
    """c = act(a_gather @ w[N,K]^T + bias)"""
    self.gemm(a, a_map, koff, w, 1, K, bias, c, c_map, M, N, K, RELU if relu else 0)

  def linear_dgrad(self, dy, dy_map, w, dx, dx_map, M, N, K, mask=None, mask_map=None, accum=False,
                   dx_koff=None):
    """dx[M,K] (+)= dy[M,N] @ w[N,K], optionally masked by (mask > 0) (ReLU of the producer)"""
    self.gemm(dy, dy_map, None, w, K, 1, None, dx, dx_map, M, K, N, ACCUM if accum else 0, mask,
              mask_map, dx_koff)

  def relu_bwd(self, dy, dy_map, act, act_map, out, out_map, M, N):
    a, b, c = dy_map.c(), act_map.c(), out_map.c()
    check(self.lib.v4l_relu_bwd(self.h, self.ctx.stream(), ptr(dy), C.byref(a), ptr(act), C.byref(b),
                                ptr(out), C.byref(c), M, N))
    self.launches += 1

  def linear_wgrad(self, dy, dy_map, a, a_map, koff, dw, dbias, M, N, K):
    g = WgradArgs(ptr(dy), dy_map.c(), ptr(a), a_map.c(), ptr(koff), ptr(dw), K, ptr(dbias), M, N, K)
    check(self.lib.v4l_gemm_wgrad(self.h, self.ctx.stream(), C.byref(g)))
    self.launches += 2

  def col2im(self, dcol, x, dx, B, Hin, Win, Cc, KH, KW, stride, Hout, Wout):
    check(self.lib.v4l_col2im(self.h, self.ctx.stream(), ptr(dcol), ptr(x), ptr(dx), B, Hin, Win, Cc,
                              KH, KW, stride, Hout, Wout))
    self.launches += 1

  # ---- tensor-core tier
  def tc_gemm(self, a, a_shape, out_grid, box, taps, kchunks, w, N_pad, N_valid, bias, c, c_map,
              c_f32=False, mask=None, flags=0, a_strides=None, a_idx=None, a_off=0, res=None):
    """a: fp16 [a_B, a_H, a_W, a_C]; out_grid (B, Hout, Wout); box (bw, bh, bb); taps [(dw, dh)];
    a_strides: element strides (sW, sH, sB) of a non-packed view; a_off: element offset"""
    g = TcGemmArgs()
    g.a = ptr(a) + 2 * a_off
    g.a_B, g.a_H, g.a_W, g.a_C = a_shape
    if a_strides:
      g.a_sW, g.a_sH, g.a_sB = a_strides
    g.a_idx = ptr(a_idx)
    g.B, g.Hout, g.Wout = out_grid
    g.bw, g.bh, g.bb = box
    g.n_taps, g.kchunks = len(taps), kchunks
    for i, (dw, dh) in enumerate(taps):
      g.tap_dw[i], g.tap_dh[i] = dw, dh
    g.w, g.N_pad, g.N_valid = ptr(w), N_pad, N_valid
    g.bias, g.c, g.c_map, g.c_f32 = ptr(bias), ptr(c), c_map.c(), 1 if c_f32 else 0
    g.mask, g.flags, g.res = ptr(mask), flags, ptr(res)
    check(self.lib.v4l_tc_gemm(self.h, self.ctx.stream(), C.byref(g)))
    self.launches += 1
    self._note("v4l_tc_gemm", g, 2.0 * out_grid[0] * out_grid[1] * out_grid[2] * N_valid * len(taps) *
               min(a_shape[3], kchunks * 64))

  def tc_conv_flat(self, x, C_, P, Wg, Hout, Wout, taps, w, N_pad, N_valid, bias, c, c_map, n_img, x_idx=None,
                   flags=0, mode=0):
    """valid convolution on the flattened (image, position) grid; x fp16 [rows, C_]; taps [(dw, dh)]"""
    g = _lib.TcConvFlatArgs()
    g.x, g.x_rows, g.C = ptr(x), x.numel() // C_, C_
    g.P, g.Wg, g.Hout, g.Wout = P, Wg, Hout, Wout
    g.n_taps = len(taps)
    for i, (dw_, dh_) in enumerate(taps):
      g.tap_dw[i], g.tap_dh[i] = dw_, dh_
    g.w, g.N_pad, g.N_valid, g.bias = ptr(w), N_pad, N_valid, ptr(bias)
    g.c, g.c_map, g.n_img, g.x_idx, g.flags, g.mode = ptr(c), c_map.c(), n_img, ptr(x_idx), flags, mode
    check(self.lib.v4l_tc_conv_flat(self.h, self.ctx.stream(), C.byref(g)))
    self.launches += 1
    self._note("v4l_tc_conv_flat", g, 2.0 * n_img * Hout * Wout * N_valid * len(taps) * C_)

  def tc_wgrad(self, x, x_shape, dy, dy_C, out_grid, box, taps, N_valid, index, dw, x_idx=None,
               x_estride=1, subs=None, dy_strides=None, dy_off=0, x_strides=None, out_scale=1.0,
               dbias=None, defer=False, algo_flops=None):
    """subs: [(sub_dw, sub_dh, dy_channel_offset)] sub-iterations per tile (space-to-depth cells)"""
    g = TcWgradArgs()
    g.x = ptr(x)
    g.x_B, g.x_H, g.x_W, g.x_C = x_shape
    if x_strides:
      g.x_sW, g.x_sH, g.x_sB = x_strides
    g.x_estride, g.x_idx = x_estride, ptr(x_idx)
    g.dy, g.dy_C = ptr(dy) + 2 * dy_off, dy_C
    if dy_strides:
      g.dy_sW, g.dy_sH, g.dy_sB = dy_strides
    if subs:
      g.n_sub = len(subs)
      for i, (sw, sh, sc) in enumerate(subs):
        g.sub_dw[i], g.sub_dh[i], g.sub_dyc[i] = sw, sh, sc
    g.B, g.Hout, g.Wout = out_grid
    g.bw, g.bh, g.bb = box
    g.n_taps = len(taps)
    for i, (dw_, dh_) in enumerate(taps):
      g.tap_dw[i], g.tap_dh[i] = dw_, dh_
    g.N_valid, g.index, g.dw = N_valid, ptr(index), ptr(dw)
    g.out_scale = out_scale
    g.dbias, g.defer = ptr(dbias), 1 if defer else 0
    check(self.lib.v4l_tc_wgrad(self.h, self.ctx.stream(), C.byref(g)))
    self.launches += 1 if defer else 2
    if algo_flops is None:       # dW: 2 x rows x K x N (+ db); algo_flops overrides it (LayerNorm affine: diagonal only)
      algo_flops = 2.0 * out_grid[0] * out_grid[1] * out_grid[2] * N_valid * (len(taps) * x_shape[3] + 1)
    self._note("v4l_tc_wgrad", g, algo_flops)

  def tc_mlp_chain(self, x, M, x_cols, x_ld, layers):
    """up to three chained Linear layers in one launch (v4l_tc_mlp_chain).  layers: list of dicts with keys
    w, K, N_pad, N_valid, bias, relu, mask, mask_ld, out, out_f32, out_map (RM)"""
    g = _lib.TcMlpChainArgs()
    g.x, g.M, g.x_cols, g.x_ld, g.n_layers = ptr(x), M, x_cols, x_ld, len(layers)
    flops = 0.0
    for i, L in enumerate(layers):
      e = g.layer[i]
      e.w, e.K, e.N_pad, e.N_valid = ptr(L["w"]), L["K"], L["N_pad"], L["N_valid"]
      e.bias, e.relu = ptr(L.get("bias")), 1 if L.get("relu") else 0
      e.mask, e.mask_ld = ptr(L.get("mask")), L.get("mask_ld", 0)
      e.out, e.out_f32 = ptr(L.get("out")), 1 if L.get("out_f32") else 0
      e.out_map = (L.get("out_map") or RM()).c()
      flops += 2.0 * M * L["N_valid"] * L.get("K_true", L["K"])
    check(self.lib.v4l_tc_mlp_chain(self.h, self.ctx.stream(), C.byref(g)))
    self.launches += 1
    self._note("v4l_tc_mlp_chain", g, flops, tuple(layers))

  def tc_wgrad_conv1(self, x_s2d, x_idx, dy_cells, B, index, dw, dbias, out_scale=1.0, defer=True, accumulate=False):
    """conv1 weight + bias gradient on the space-to-depth image / cell layouts (v4l_tc_wgrad_conv1)"""
    check(self.lib.v4l_tc_wgrad_conv1(self.h, self.ctx.stream(), ptr(x_s2d), x_s2d.shape[0], ptr(x_idx), ptr(dy_cells),
                                      B, ptr(index), ptr(dw), ptr(dbias), out_scale, 1 if defer else 0,
                                      1 if accumulate else 0))
    self.launches += 1 if defer else 2
    if self._rec is not None:          # replayed through a closure (plain-argument entry point)
      args = (ptr(x_s2d), x_s2d.shape[0], ptr(x_idx), ptr(dy_cells), B, ptr(index), ptr(dw), ptr(dbias), out_scale,
              1 if defer else 0, 1 if accumulate else 0)
      self._rec.append(("v4l_tc_wgrad_conv1", args, 2.0 * B * 225 * 32 * 257, ()))

  def tc_wgrad_flush(self):
    check(self.lib.v4l_tc_wgrad_flush(self.h, self.ctx.stream()))
    self.launches += 1

  def colsum_f16(self, dy, dy_map, M, N, out, fold=1, out_scale=1.0):
    m = dy_map.c()
    check(self.lib.v4l_colsum_f16(self.h, self.ctx.stream(), ptr(dy), C.byref(m), M, N, fold, out_scale,
                                  ptr(out)))
    self.launches += 2

  def ingest_img(self, img_f32, out_s2d, n, idx=None):
    check(self.lib.v4l_ingest_img(self.h, self.ctx.stream(), ptr(img_f32), ptr(out_s2d), n, ptr(idx)))
    self.launches += 1

  def ingest_img_f16(self, img_f16_ptr, out_s2d, n, idx=None):
    check(self.lib.v4l_ingest_img_f16(self.h, self.ctx.stream(), img_f16_ptr, ptr(out_s2d), n, ptr(idx)))
    self.launches += 1

  def ingest_rows(self, obs_ptr, row_stride, S, idx, n_rows, state_out, img_out, s2d_out):
    check(self.lib.v4l_ingest_rows(self.h, self.ctx.stream(), obs_ptr, row_stride, S, ptr(idx), n_rows,
                                   ptr(state_out), ptr(img_out), ptr(s2d_out)))
    self.launches += 1

  def gather_rows_f16(self, src, src_is_f32, idx, dst, rows, src_cols, src_stride, dst_cols, scale=1.0):
    check(self.lib.v4l_gather_rows_f16(self.h, self.ctx.stream(), ptr(src), 1 if src_is_f32 else 0, ptr(idx),
                                       ptr(dst), rows, src_cols, src_stride, dst_cols, scale))
    self.launches += 1

  def relu_bwd_f16(self, dy, dy_map, act, act_map, out, out_map, M, N):
    a, b, c = dy_map.c(), act_map.c(), out_map.c()
    check(self.lib.v4l_relu_bwd_f16(self.h, self.ctx.stream(), ptr(dy), C.byref(a), ptr(act), C.byref(b),
                                     ptr(out), C.byref(c), M, N))
    self.launches += 1

  def attn_fwd_f16(self, qkv, o, p, B, T, d, nh):
    if nh == 1 and d == 64:          # tensor-core path: block-diagonal tcgen05 MMAs
      check(self.lib.v4l_tc_attn_fwd(self.h, self.ctx.stream(), ptr(qkv), ptr(o), ptr(p), B, T))
      self.launches += 1
      return
    check(self.lib.v4l_attn_fwd_f16(self.h, self.ctx.stream(), ptr(qkv), ptr(o), ptr(p), B, T, d, nh))
    self.launches += 1

  def tc_block_fwd(self, x, B, T, w, par, out, eps=1e-5):
    """One fused TransformerEncoderLayer forward (v4l_tc_block_fwd). w: dict of fp16 weight tensors
    (w_in, w_o, w_1, w_2); par: dict of fp32 vectors (b_in, b_o, g1, be1, b1, b2, g2, be2);
    out: dict of output/saved tensors (qkv, o, h, f1, y, p, z1, st1, z2, st2)."""
    a = _lib.TcBlockArgs()
    a.x = ptr(x); a.B = B; a.T = T; a.eps = eps
    for k in ("w_in", "w_o", "w_1", "w_2"):
      setattr(a, k, ptr(w[k]))
    for k in ("b_in", "b_o", "g1", "be1", "b1", "b2", "g2", "be2"):
      setattr(a, k, ptr(par[k]))
    for k in ("qkv", "o", "h", "f1", "y", "p", "z1", "st1", "z2", "st2", "xh1", "xh2"):
      setattr(a, k, ptr(out.get(k)))
    check(self.lib.v4l_tc_block_fwd(self.h, self.ctx.stream(), C.byref(a)))
    self.launches += 1
    self._note("v4l_tc_block_fwd", a, B * (2.0 * T * 64 * (192 + 64 + 256 + 256) + 4.0 * T * T * 64))

  def tc_block_bwd(self, dy, B, T, saved, w, g1, g2, out):
    """Fused data-gradient pass of one encoder layer (v4l_tc_block_bwd). saved: qkv, xh1, xh2, f1, p,
    st1, st2 from tc_block_fwd; w: fp16 dgrad-orientation weights (w2d, w1d, wod, wind);
    out: dz2, df1, dh, dz1, dqkv, dx."""
    a = _lib.TcBlockBwdArgs()
    a.dy = ptr(dy); a.B = B; a.T = T
    for k in ("qkv", "xh1", "xh2", "f1", "p", "st1", "st2"):
      setattr(a, k, ptr(saved[k]))
    a.g1, a.g2 = ptr(g1), ptr(g2)
    for k in ("w2d", "w1d", "wod", "wind"):
      setattr(a, k, ptr(w[k]))
    for k in ("dz2", "df1", "dh", "dz1", "dqkv", "dx"):
      setattr(a, k, ptr(out[k]))
    check(self.lib.v4l_tc_block_bwd(self.h, self.ctx.stream(), C.byref(a)))
    self.launches += 1
    # data gradients only: the four projections + dP, dQ, dK, dV (the weight gradients are separate launches)
    self._note("v4l_tc_block_bwd", a, B * (2.0 * T * 64 * (192 + 64 + 256 + 256) + 8.0 * T * T * 64))

  def attn_bwd_f16(self, qkv, p, d_o, d_qkv, B, T, d, nh):
    if nh == 1 and d == 64:
      check(self.lib.v4l_tc_attn_bwd(self.h, self.ctx.stream(), ptr(qkv), ptr(p), ptr(d_o), ptr(d_qkv), B, T))
      self.launches += 1
      return
    check(self.lib.v4l_attn_bwd_f16(self.h, self.ctx.stream(), ptr(qkv), ptr(p), ptr(d_o), ptr(d_qkv),
                                     B, T, d, nh))
    self.launches += 1

  def ln_fwd_f16(self, a, res, gamma, beta, y, z, stats, rows, d, eps=1e-5):
    check(self.lib.v4l_ln_fwd_f16(self.h, self.ctx.stream(), ptr(a), ptr(res), ptr(gamma), ptr(beta),
                                   ptr(y), ptr(z), ptr(stats), rows, d, eps))
    self.launches += 1

  def ln_bwd_f16(self, dy, z, stats, gamma, dz, dgamma, dbeta, rows, d, out_scale=1.0):
    check(self.lib.v4l_ln_bwd_f16(self.h, self.ctx.stream(), ptr(dy), ptr(z), ptr(stats), ptr(gamma),
                                  ptr(dz), ptr(dgamma), ptr(dbeta), rows, d, out_scale))
    self.launches += 2

  def pool_fwd_f16(self, tok, out, B, T, d, mode):
    check(self.lib.v4l_pool_fwd_f16(self.h, self.ctx.stream(), ptr(tok), ptr(out), B, T, d, mode))
    self.launches += 1

  def pool_bwd_f16(self, dout, dtok, B, T, d, mode):
    check(self.lib.v4l_pool_bwd_f16(self.h, self.ctx.stream(), ptr(dout), ptr(dtok), B, T, d, mode))
    self.launches += 1

  def pack_f16(self, src, index, dst, n):
    check(self.lib.v4l_pack_f16(self.h, self.ctx.stream(), ptr(src), ptr(index), ptr(dst), n))
    self.launches += 1

  # ---- transformer pieces
  def attn_fwd(self, qkv, o, p, B, T, d, nh):
    check(self.lib.v4l_attn_fwd(self.h, self.ctx.stream(), ptr(qkv), ptr(o), ptr(p), B, T, d, nh))
    self.launches += 1

  def attn_bwd(self, qkv, p, d_o, d_qkv, B, T, d, nh):
    check(self.lib.v4l_attn_bwd(self.h, self.ctx.stream(), ptr(qkv), ptr(p), ptr(d_o), ptr(d_qkv),
                                B, T, d, nh))
    self.launches += 1

  def ln_fwd(self, a, res, gamma, beta, y, z, stats, rows, d, eps=1e-5):
    check(self.lib.v4l_ln_fwd(self.h, self.ctx.stream(), ptr(a), ptr(res), ptr(gamma), ptr(beta),
                              ptr(y), ptr(z), ptr(stats), rows, d, eps))
    self.launches += 1

  def ln_bwd(self, dy, z, stats, gamma, dz, dgamma, dbeta, rows, d):
    check(self.lib.v4l_ln_bwd(self.h, self.ctx.stream(), ptr(dy), ptr(z), ptr(stats), ptr(gamma),
                              ptr(dz), ptr(dgamma), ptr(dbeta), rows, d))
    self.launches += 2

  def pool_fwd(self, tok, out, B, T, d, mode):
    check(self.lib.v4l_pool_fwd(self.h, self.ctx.stream(), ptr(tok), ptr(out), B, T, d, mode))
    self.launches += 1

  def pool_bwd(self, dout, dtok, B, T, d, mode):
    check(self.lib.v4l_pool_bwd(self.h, self.ctx.stream(), ptr(dout), ptr(dtok), B, T, d, mode))
    self.launches += 1

  # ---- PPO pieces
  def gae(self, rewards, values, terminals, time_limits, tl_st, tl_se, last_value, advs, rets, T, E,
          gamma, tau, time_limit_filter, mode=0):
    check(self.lib.v4l_gae(self.h, self.ctx.stream(), ptr(rewards), ptr(values), ptr(terminals),
                           ptr(time_limits), tl_st, tl_se, ptr(last_value), ptr(advs), ptr(rets), T, E,
                           gamma, tau, 1 if time_limit_filter else 0, mode))
    self.launches += 3

  def select_rows(self, flat_idx, slot, cur_idx, n):
    check(self.lib.v4l_select_rows(self.h, self.ctx.stream(), ptr(flat_idx), ptr(slot), ptr(cur_idx), n))
    self.launches += 1

  def slot_advance(self, slot, wrap=0):
    check(self.lib.v4l_slot_advance(self.h, self.ctx.stream(), ptr(slot), wrap))
    self.launches += 1

  def adv_stats(self, adv, idx, n, stats):
    check(self.lib.v4l_adv_stats(self.h, self.ctx.stream(), ptr(adv), ptr(idx), n, ptr(stats)))
    self.launches += 1

  def vf_loss(self, values, returns, old_values, idx, d_values, n, inv_global, inv_local, clipped,
              clip_para, info, slot, d_f16=None, scale_f16=1.0):
    """one launch: loss, gradient (fp32 and, optionally, loss-scaled fp16 [n,16]) and the logged value"""
    check(self.lib.v4l_vf_loss(self.h, self.ctx.stream(), ptr(values), ptr(returns), ptr(old_values),
                               ptr(idx), ptr(d_values), n, inv_global, inv_local, 1 if clipped else 0,
                               clip_para, ptr(info), ptr(slot), ptr(d_f16), scale_f16))
    self.launches += 1

  def pf_loss(self, mean, logstd, tmean, tlogstd, acts, adv, idx, stats, d_mean, d_logstd, n, A,
              inv_global, inv_local, clip_para, entropy_coeff, info, slot, target_indexed=False,
              d_f16=None, scale_f16=1.0, stats_per_slot=False):
    check(self.lib.v4l_pf_loss(self.h, self.ctx.stream(), ptr(mean), ptr(logstd), ptr(tmean),
                               ptr(tlogstd), ptr(acts), ptr(adv), ptr(idx), ptr(stats), ptr(d_mean),
                               ptr(d_logstd), n, A, inv_global, inv_local, clip_para, entropy_coeff,
                               ptr(info), ptr(slot), 1 if target_indexed else 0, ptr(d_f16), scale_f16,
                               1 if stats_per_slot else 0))
    self.launches += 1

  def adv_stats_epoch(self, flat_idx, n_mb, n, adv, stats):
    check(self.lib.v4l_adv_stats_epoch(self.h, self.ctx.stream(), ptr(flat_idx), n_mb, n, ptr(adv), ptr(stats)))
    self.launches += 1

  def depth_frame(self, zbuf, ring, reset, E, n_slots, head, near=0.01, far=1000.0):
    check(self.lib.v4l_depth_frame(self.h, self.ctx.stream(), ptr(zbuf), ptr(ring), ptr(reset), E, n_slots, head,
                                   near, far))
    self.launches += 1

  def stack_frames(self, ring, slots, E, n_slots, normalise, out_s2d=None, out_chw=None, chw_stride=0):
    check(self.lib.v4l_stack_frames(self.h, self.ctx.stream(), ptr(ring), ptr(slots), E, n_slots, int(normalise),
                                    ptr(out_s2d), ptr(out_chw), chw_stride))
    self.launches += 1

  def normalizer(self, x, n, S, mean, var, count, update, clip, out=None):
    check(self.lib.v4l_normalizer(self.h, self.ctx.stream(), ptr(x), n, S, ptr(mean), ptr(var), float(count),
                                  int(update), float(clip), ptr(out)))
    self.launches += 1

  def mb_begin(self, flat_idx, slot, cur_idx, n, adv, stats, state=None, S=0, state_f16=None, Sp=0):
    """minibatch prologue: row selection + advantage statistics (+ proprio rows -> fp16), one launch"""
    check(self.lib.v4l_mb_begin(self.h, self.ctx.stream(), ptr(flat_idx), ptr(slot), ptr(cur_idx), n, ptr(adv),
                                ptr(stats), ptr(state), S, ptr(state_f16), Sp))
    self.launches += 1

  def opt_tail(self, phases, param=None, grad=None, m=None, v=None, n=0, hyper=None, info=None, slot=None,
               norm_slot=-1, extra=(0, 0), scatter=None, packed_self=None, packed_other=None, slot_advance=None):
    """fused optimiser tail (v4l_opt_tail): phases bit 0 = split-K reduction of the deferred weight-
    gradient partials, bit 1 = clip + Adam + fp16 operand copies + step / slot counters"""
    a = _lib.OptTailArgs()
    a.phases = phases
    a.param, a.grad, a.m, a.v, a.n = ptr(param), ptr(grad), ptr(m), ptr(v), n
    a.hyper, a.info, a.slot, a.norm_slot = ptr(hyper), ptr(info), ptr(slot), norm_slot
    a.extra_lo, a.extra_n = extra
    a.scatter, a.packed_self, a.packed_other = ptr(scatter), ptr(packed_self), ptr(packed_other)
    a.slot_advance = ptr(slot_advance)
    check(self.lib.v4l_opt_tail(self.h, self.ctx.stream(), C.byref(a)))
    self.launches += 2 if phases & 2 else 1

  def opt_tail_error(self):
    return self.lib.v4l_opt_tail_error(self.h)

  def clip_adam(self, param, grad, m, v, n, hyper, info, slot, norm_slot):
    check(self.lib.v4l_clip_adam(self.h, self.ctx.stream(), ptr(param), ptr(grad), ptr(m), ptr(v), n,
                                 ptr(hyper), ptr(info), ptr(slot), norm_slot))
    self.launches += 3

  def h2d_raw(self, dst_ptr, src_ptr, nbytes):
    """contiguous pinned-host -> device copy on the current stream (copy engine)"""
    check(self.lib.v4l_h2d_2d(self.ctx.stream(), dst_ptr, nbytes, src_ptr, nbytes, nbytes, 1))

  def h2d_rows(self, dst_ptr, src_ptr, rows, row_bytes):
    """rows: contiguous int32 numpy array of row numbers; copies those rows pinned-host -> device"""
    check(self.lib.v4l_h2d_rows(self.ctx.stream(), dst_ptr, src_ptr, rows.ctypes.data, len(rows), row_bytes))

  def h2d_2d(self, dst, dpitch, src_ptr, spitch, width, height):
    check(self.lib.v4l_h2d_2d(self.ctx.stream(), ptr(dst), dpitch, src_ptr, spitch, width, height))


_ops = {}


def ops_for(device):
  device = torch.device(device)
  index = device.index if device.index is not None else torch.cuda.current_device()
  o = _ops.get(index)
  if o is None:
    o = _ops[index] = Ops(torch.device("cuda", index))
  return o


# =================================================================================================
# gather tables (im2col expressed as row-origin + per-k offsets)
# =================================================================================================

def _i32(a, device):
  return torch.tensor(np.ascontiguousarray(a, dtype=np.int32), device=device)


def conv_tables(device, Hin, Win, Cin, KH, KW, stride, chw_input):
  """Returns (pos_off [Hout*Wout], k_off [Cin*KH*KW], Hout, Wout) for a VALID conv whose k
  index runs (c, kh, kw) — torch's OIHW weight order — over an input stored CHW
  (chw_input=True: the observation image) or HWC (our activations)."""
  Hout, Wout = (Hin - KH) // stride + 1, (Win - KW) // stride + 1
  oh, ow = np.meshgrid(np.arange(Hout), np.arange(Wout), indexing="ij")
  c, kh, kw = np.meshgrid(np.arange(Cin), np.arange(KH), np.arange(KW), indexing="ij")
  if chw_input:
    pos = (oh * stride) * Win + ow * stride
    k = c * (Hin * Win) + kh * Win + kw
  else:
    pos = ((oh * stride) * Win + ow * stride) * Cin
    k = (kh * Win + kw) * Cin + c
  return _i32(pos.ravel(), device), _i32(k.ravel(), device), Hout, Wout


class _Plan:
  """Common machinery: workspace cache and dense Linear-stack helpers."""

  def __init__(self, ops, out_dim):
    self.ops = ops
    self.device = ops.device
    self.out_dim = out_dim
    self._ws = {}

  def buf(self, name, *shape):
    key = (name,) + tuple(shape)
    t = self._ws.get(key)
    if t is None:
      t = self._ws[key] = torch.empty(shape, device=self.device, dtype=torch.float32)
    return t

  def release(self):
    self._ws.clear()

  def _stack_fwd(self, P, keys, x, x_map, koff, B, din, tag, last_relu, out=None, out_map=None):
    """Linear stack: every layer has ReLU except (optionally) the last.  The last layer may
    write into a caller-provided view (out, out_map).  Returns [(act, map, width)]."""
    acts = []
    a, a_map, ko, K = x, x_map, koff, din
    for i, (wk, bk) in enumerate(keys):
      w = P[wk]
      N = w.shape[0]
      last = i + 1 == len(keys)
      if last and out is not None:
        y, y_map = out, out_map
      else:
        y = self.buf("%s%d" % (tag, i), B, N)
        y_map = RM.dense(N)
      self.ops.linear_fwd(a, a_map, ko, w, P[bk], y, y_map, B, N, K, relu=(not last) or last_relu)
      acts.append((y, y_map, N))
      a, a_map, ko, K = y, y_map, None, N
    return acts

  def _stack_bwd(self, P, G, keys, x, x_map, koff, B, din, acts, dy, dy_map, tag, dx=None,
                 dx_map=None, dx_mask=None, dx_mask_map=None):
    """dy: gradient w.r.t. the last layer's PRE-activation.  If dx is given, the gradient
    w.r.t. the stack input is written there (masked by dx_mask > 0 when the input is itself a
    ReLU output)."""
    for i in reversed(range(len(keys))):
      wk, bk = keys[i]
      w = P[wk]
      N, K = w.shape
      if i > 0:
        a, a_map, _ = acts[i - 1]
        ko = None
      else:
        a, a_map, ko = x, x_map, koff
      self.ops.linear_wgrad(dy, dy_map, a, a_map, ko, G[wk], G[bk], B, N, K)
      if i > 0:
        dprev = self.buf("%s_d%d" % (tag, i - 1), B, K)
        self.ops.linear_dgrad(dy, dy_map, w, dprev, RM.dense(K), B, N, K, mask=a, mask_map=a_map)
        dy, dy_map = dprev, RM.dense(K)
      elif dx is not None:
        self.ops.linear_dgrad(dy, dy_map, w, dx, dx_map, B, N, K, mask=dx_mask, mask_map=dx_mask_map)


def _head_keys(P, prefix):
  idx = sorted(int(k[len(prefix):].split(".")[0]) for k in P
               if k.startswith(prefix) and k.endswith(".weight"))
  return [(prefix + "%d.weight" % i, prefix + "%d.bias" % i) for i in idx]


# =================================================================================================
# proprio-only MLP (reference starter/ppo_state.py:90-104; nets.py:16-55; base.py:8-44)
# =================================================================================================
class MLPPlan(_Plan):
  family = "mlp"

  def __init__(self, ops, S, out_dim):
    super().__init__(ops, out_dim)
    self.S = S

  def forward(self, P, inp, out):
    B = inp.B
    self._inp = inp
    x_map = RM(1, inp.state_stride, 0, inp.state_base, idx=inp.idx)
    self._bk = _head_keys(P, "base.seq_fcs.")
    self._hk = _head_keys(P, "seq_append_fcs.")
    self._bacts = self._stack_fwd(P, self._bk, inp.state, x_map, None, B, self.S, "b", True)
    h, h_map, hd = self._bacts[-1]
    self._hacts = self._stack_fwd(P, self._hk, h, h_map, None, B, hd, "h", False, out,
                                  RM.dense(self.out_dim))
    return out

  def backward(self, P, G, d_out):
    inp = self._inp
    B = inp.B
    h, h_map, hd = self._bacts[-1]
    dh = self.buf("dh", B, hd)
    self._stack_bwd(P, G, self._hk, h, h_map, None, B, hd, self._hacts, d_out,
                    RM.dense(self.out_dim), "h", dx=dh, dx_map=RM.dense(hd), dx_mask=h,
                    dx_mask_map=h_map)
    x_map = RM(1, inp.state_stride, 0, inp.state_base, idx=inp.idx)
    self._stack_bwd(P, G, self._bk, inp.state, x_map, None, B, self.S, self._bacts, dh,
                    RM.dense(hd), "b")


# =================================================================================================
# NatureCNN trunk shared by NaturePlan / LocoPlan (reference base.py:304-342)
# =================================================================================================
class _ConvTrunk:
  """conv 8x8/4 -> 4x4/2 -> 3x3/1 (+ReLU each) on a [4,64,64] CHW image; activations NHWC."""

  def __init__(self, plan, prefix):
    self.plan = plan
    self.prefix = prefix
    dev = plan.device
    self.pos1, self.k1, h1, w1 = conv_tables(dev, 64, 64, IMG_C, 8, 8, 4, True)     # 15x15
    self.pos2, self.k2, h2, w2 = conv_tables(dev, h1, w1, 32, 4, 4, 2, False)       # 6x6
    self.pos3, self.k3, h3, w3 = conv_tables(dev, h2, w2, 64, 3, 3, 1, False)       # 4x4
    assert (h1, h2, h3) == (15, 6, 4)

  def keys(self, i):
    return self.prefix + "layers.%d.weight" % (2 * i), self.prefix + "layers.%d.bias" % (2 * i)

  def forward(self, P, inp):
    pl, ops, B = self.plan, self.plan.ops, inp.B
    self.inp = inp
    self.img_map = RM(225, inp.img_stride, 0, inp.img_base, idx=inp.idx, pos_off=self.pos1)
    w, b = self.keys(0)
    self.a1 = pl.buf("a1", B, 225, 32)
    ops.linear_fwd(inp.img, self.img_map, self.k1, P[w], P[b], self.a1, RM.dense(32), B * 225, 32, 256, True)
    self.a1_map = RM(36, 225 * 32, 0, 0, pos_off=self.pos2)
    w, b = self.keys(1)
    self.a2 = pl.buf("a2", B, 36, 64)
    ops.linear_fwd(self.a1, self.a1_map, self.k2, P[w], P[b], self.a2, RM.dense(64), B * 36, 64, 512, True)
    self.a2_map = RM(16, 36 * 64, 0, 0, pos_off=self.pos3)
    w, b = self.keys(2)
    self.a3 = pl.buf("a3", B, 16, 64)
    ops.linear_fwd(self.a2, self.a2_map, self.k3, P[w], P[b], self.a3, RM.dense(64), B * 16, 64, 576, True)
    return self.a3

  def backward(self, P, G, da3):
    """da3 [B,16,64]: gradient w.r.t. conv3's PRE-activation (already ReLU-masked)."""
    pl, ops, B = self.plan, self.plan.ops, self.inp.B
    inp = self.inp
    w3, b3 = self.keys(2)
    ops.linear_wgrad(da3, RM.dense(64), self.a2, self.a2_map, self.k3, G[w3], G[b3], B * 16, 64, 576)
    dcol = pl.buf("dcol", B * 36 * 512)             # shared by conv3 (B*16*576) and conv2 (B*36*512)
    ops.linear_dgrad(da3, RM.dense(64), P[w3], dcol, RM.dense(576), B * 16, 64, 576)
    da2 = pl.buf("da2", B, 36, 64)
    ops.col2im(dcol, self.a2, da2, B, 6, 6, 64, 3, 3, 1, 4, 4)
    w2, b2 = self.keys(1)
    ops.linear_wgrad(da2, RM.dense(64), self.a1, self.a1_map, self.k2, G[w2], G[b2], B * 36, 64, 512)
    ops.linear_dgrad(da2, RM.dense(64), P[w2], dcol, RM.dense(512), B * 36, 64, 512)
    da1 = pl.buf("da1", B, 225, 32)
    ops.col2im(dcol, self.a1, da1, B, 15, 15, 32, 4, 4, 2, 6, 6)
    w1, b1 = self.keys(0)
    ops.linear_wgrad(da1, RM.dense(32), inp.img, self.img_map, self.k1, G[w1], G[b1], B * 225, 32, 256)


# =================================================================================================
# NatureCNN + concat MLP  (reference nets.py:194-262, base.py:345-385; starter/ppo_nature_cnn.py)
# =================================================================================================
class NaturePlan(_Plan):
  family = "nature"

  def __init__(self, ops, S, out_dim):
    super().__init__(ops, out_dim)
    self.S = S
    self.trunk = _ConvTrunk(self, "encoder.visual_base.")
    # flatten of [B,64,4,4] is (c, p); our a3 is (p, c)
    c, p = np.meshgrid(np.arange(64), np.arange(16), indexing="ij")
    self.kflat = _i32((p * 64 + c).ravel(), self.device)

  def forward(self, P, inp, out):
    ops, B = self.ops, inp.B
    self._inp = inp
    a3 = self.trunk.forward(P, inp)
    wv, bv = "encoder.visual_projector.projection.0.weight", "encoder.visual_projector.projection.0.bias"
    self.vd = P[wv].shape[0]
    self._bk = _head_keys(P, "encoder.base.seq_fcs.")
    self._hk = _head_keys(P, "seq_append_fcs.")
    self.sd = P[self._bk[-1][0]].shape[0]
    W = self.vd + self.sd
    self.cat = self.buf("cat", B, W)
    ops.linear_fwd(a3, RM.dense(1024), self.kflat, P[wv], P[bv], self.cat, RM(1, W, 0, 0), B, self.vd, 1024, True)
    x_map = RM(1, inp.state_stride, 0, inp.state_base, idx=inp.idx)
    self._bacts = self._stack_fwd(P, self._bk, inp.state, x_map, None, B, self.S, "s", True,
                                  self.cat, RM(1, W, 0, self.vd))
    self._hacts = self._stack_fwd(P, self._hk, self.cat, RM.dense(W), None, B, W, "h", False, out,
                                  RM.dense(self.out_dim))
    return out

  def backward(self, P, G, d_out):
    ops, inp = self.ops, self._inp
    B, W = inp.B, self.vd + self.sd
    dcat = self.buf("dcat", B, W)
    self._stack_bwd(P, G, self._hk, self.cat, RM.dense(W), None, B, W, self._hacts, d_out,
                    RM.dense(self.out_dim), "h", dx=dcat, dx_map=RM.dense(W), dx_mask=self.cat,
                    dx_mask_map=RM.dense(W))
    x_map = RM(1, inp.state_stride, 0, inp.state_base, idx=inp.idx)
    self._stack_bwd(P, G, self._bk, inp.state, x_map, None, B, self.S, self._bacts, dcat,
                    RM(1, W, 0, self.vd), "s")
    wv, bv = "encoder.visual_projector.projection.0.weight", "encoder.visual_projector.projection.0.bias"
    dv_map = RM(1, W, 0, 0)
    ops.linear_wgrad(dcat, dv_map, self.trunk.a3, RM.dense(1024), self.kflat, G[wv], G[bv], B, self.vd, 1024)
    da3 = self.buf("da3", B, 16, 64)
    ops.linear_dgrad(dcat, dv_map, P[wv], da3, RM.dense(1024), B, self.vd, 1024, mask=self.trunk.a3,
                     mask_map=RM.dense(1024), dx_koff=self.kflat)
    self.trunk.backward(P, G, da3)


# =================================================================================================
# LocoTransformer (reference nets.py:909-1038, base.py:497-626) and the vision-only
# Transformer (nets.py:784-906, base.py:388-494; has_state=False)
# =================================================================================================
class LocoPlan(_Plan):
  family = "loco"

  def __init__(self, ops, S, out_dim, n_heads=(1, 1), has_state=True, token_dim=64):
    super().__init__(ops, out_dim)
    self.S = S
    self.has_state = has_state
    self.d = token_dim
    self.n_heads = list(n_heads)
    self.T = 16 + (1 if has_state else 0)
    self.trunk = _ConvTrunk(self, "encoder.depth_visual_base.")

  def _layer_keys(self, l):
    p = "visual_append_layers.%d." % l
    return {k: p + v for k, v in {
      "win": "self_attn.in_proj_weight", "bin": "self_attn.in_proj_bias",
      "wo": "self_attn.out_proj.weight", "bo": "self_attn.out_proj.bias",
      "w1": "linear1.weight", "b1": "linear1.bias", "w2": "linear2.weight", "b2": "linear2.bias",
      "g1": "norm1.weight", "be1": "norm1.bias", "g2": "norm2.weight", "be2": "norm2.bias"}.items()}

  def forward(self, P, inp, out):
    ops, B, T, d = self.ops, inp.B, self.T, self.d
    self._inp = inp
    first = 1 if self.has_state else 0
    a3 = self.trunk.forward(P, inp)
    tok = self.buf("tok0", B, T, d)
    self.vis_map = RM.slots(16, T, d, first)
    ops.linear_fwd(a3, RM.dense(64), None, P["encoder.depth_up_conv.weight"],
                   P["encoder.depth_up_conv.bias"], tok, self.vis_map, B * 16, d, 64, False)
    if self.has_state:
      self._sk = _head_keys(P, "encoder.base.seq_fcs.") + [
        ("encoder.state_projector.projection.0.weight", "encoder.state_projector.projection.0.bias")]
      x_map = RM(1, inp.state_stride, 0, inp.state_base, idx=inp.idx)
      self._sacts = self._stack_fwd(P, self._sk, inp.state, x_map, None, B, self.S, "s", True, tok,
                                    RM.slots(1, T, d, 0))
    R = B * T
    x = tok
    self._layers = []
    for l, nh in enumerate(self.n_heads):
      k = self._layer_keys(l)
      ff = P[k["w1"]].shape[0]
      qkv = self.buf("qkv%d" % l, R, 3 * d)
      ops.linear_fwd(x, RM.dense(d), None, P[k["win"]], P[k["bin"]], qkv, RM.dense(3 * d), R, 3 * d, d, False)
      o = self.buf("o%d" % l, R, d)
      p = self.buf("p%d" % l, B, nh, T, T)
      ops.attn_fwd(qkv, o, p, B, T, d, nh)
      proj = self.buf("proj", R, d)
      ops.linear_fwd(o, RM.dense(d), None, P[k["wo"]], P[k["bo"]], proj, RM.dense(d), R, d, d, False)
      h = self.buf("h%d" % l, R, d)
      z1 = self.buf("z1_%d" % l, R, d)
      st1 = self.buf("st1_%d" % l, R, 2)
      ops.ln_fwd(proj, x, P[k["g1"]], P[k["be1"]], h, z1, st1, R, d)
      f1 = self.buf("f1_%d" % l, R, ff)
      ops.linear_fwd(h, RM.dense(d), None, P[k["w1"]], P[k["b1"]], f1, RM.dense(ff), R, ff, d, True)
      f2 = self.buf("f2", R, d)
      ops.linear_fwd(f1, RM.dense(ff), None, P[k["w2"]], P[k["b2"]], f2, RM.dense(d), R, d, ff, False)
      y = self.buf("y%d" % l, R, d)
      z2 = self.buf("z2_%d" % l, R, d)
      st2 = self.buf("st2_%d" % l, R, 2)
      ops.ln_fwd(f2, h, P[k["g2"]], P[k["be2"]], y, z2, st2, R, d)
      self._layers.append(dict(k=k, nh=nh, ff=ff, x=x, qkv=qkv, o=o, p=p, h=h, z1=z1, st1=st1, f1=f1,
                               z2=z2, st2=st2))
      x = y
    self._tok0 = tok
    pd = 2 * d if self.has_state else d
    self.pooled = self.buf("pooled", B, pd)
    ops.pool_fwd(x, self.pooled, B, T, d, 0 if self.has_state else 1)
    self._hk = _head_keys(P, "visual_seq_append_fcs.")
    self._hacts = self._stack_fwd(P, self._hk, self.pooled, RM.dense(pd), None, B, pd, "h", False, out,
                                  RM.dense(self.out_dim))
    return out

  def backward(self, P, G, d_out):
    ops, inp = self.ops, self._inp
    B, T, d = inp.B, self.T, self.d
    R = B * T
    pd = 2 * d if self.has_state else d
    dpool = self.buf("dpool", B, pd)
    self._stack_bwd(P, G, self._hk, self.pooled, RM.dense(pd), None, B, pd, self._hacts, d_out,
                    RM.dense(self.out_dim), "h", dx=dpool, dx_map=RM.dense(pd))
    dx = self.buf("dx", R, d)
    ops.pool_bwd(dpool, dx, B, T, d, 0 if self.has_state else 1)
    for L in reversed(self._layers):
      k, nh, ff = L["k"], L["nh"], L["ff"]
      dz2 = self.buf("dz2", R, d)
      ops.ln_bwd(dx, L["z2"], L["st2"], P[k["g2"]], dz2, G[k["g2"]], G[k["be2"]], R, d)
      ops.linear_wgrad(dz2, RM.dense(d), L["f1"], RM.dense(ff), None, G[k["w2"]], G[k["b2"]], R, d, ff)
      df1 = self.buf("df1", R, ff)
      ops.linear_dgrad(dz2, RM.dense(d), P[k["w2"]], df1, RM.dense(ff), R, d, ff, mask=L["f1"],
                       mask_map=RM.dense(ff))
      ops.linear_wgrad(df1, RM.dense(ff), L["h"], RM.dense(d), None, G[k["w1"]], G[k["b1"]], R, ff, d)
      ops.linear_dgrad(df1, RM.dense(ff), P[k["w1"]], dz2, RM.dense(d), R, ff, d, accum=True)   # dh
      dz1 = self.buf("dz1", R, d)
      ops.ln_bwd(dz2, L["z1"], L["st1"], P[k["g1"]], dz1, G[k["g1"]], G[k["be1"]], R, d)
      ops.linear_wgrad(dz1, RM.dense(d), L["o"], RM.dense(d), None, G[k["wo"]], G[k["bo"]], R, d, d)
      do = self.buf("do", R, d)
      ops.linear_dgrad(dz1, RM.dense(d), P[k["wo"]], do, RM.dense(d), R, d, d)
      dqkv = self.buf("dqkv", R, 3 * d)
      ops.attn_bwd(L["qkv"], L["p"], do, dqkv, B, T, d, nh)
      ops.linear_wgrad(dqkv, RM.dense(3 * d), L["x"], RM.dense(d), None, G[k["win"]], G[k["bin"]], R, 3 * d, d)
      ops.linear_dgrad(dqkv, RM.dense(3 * d), P[k["win"]], dz1, RM.dense(d), R, 3 * d, d, accum=True)
      # dx for the next (earlier) layer lives in dz1; swap so the buffers are not clobbered
      self._ws[("dx", R, d)], self._ws[("dz1", R, d)] = dz1, dx
      dx = dz1
    tok = self._tok0
    if self.has_state:
      ds = self.buf("ds", B, d)
      smap = RM.slots(1, T, d, 0)
      ops.relu_bwd(dx, smap, tok, smap, ds, RM.dense(d), B, d)
      x_map = RM(1, inp.state_stride, 0, inp.state_base, idx=inp.idx)
      self._stack_bwd(P, G, self._sk, inp.state, x_map, None, B, self.S, self._sacts, ds, RM.dense(d), "s")
    a3 = self.trunk.a3
    ops.linear_wgrad(dx, self.vis_map, a3, RM.dense(64), None, G["encoder.depth_up_conv.weight"],
                     G["encoder.depth_up_conv.bias"], B * 16, d, 64)
    da3 = self.buf("da3", B, 16, 64)
    ops.linear_dgrad(dx, self.vis_map, P["encoder.depth_up_conv.weight"], da3, RM.dense(64), B * 16, d, 64,
                     mask=a3, mask_map=RM.dense(64))
    self.trunk.backward(P, G, da3)


class NatureVOPlan(_Plan):
  """Vision-only NatureCNN: flatten(conv trunk) -> MLP head (reference nets.py:133-191 with a
  flattening NatureEncoder, starter/ppo_nature_cnn_vision_only.py)."""
  family = "nvo"

  def __init__(self, ops, S, out_dim):
    super().__init__(ops, out_dim)
    self.S = 0
    self.trunk = _ConvTrunk(self, "encoder.")
    c, p = np.meshgrid(np.arange(64), np.arange(16), indexing="ij")
    self.kflat = _i32((p * 64 + c).ravel(), self.device)      # torch flatten (c,p) over our (p,c)

  def forward(self, P, inp, out):
    B = inp.B
    self._inp = inp
    a3 = self.trunk.forward(P, inp)
    self._hk = _head_keys(P, "seq_append_fcs.")
    self._hacts = self._stack_fwd(P, self._hk, a3, RM.dense(1024), self.kflat, B, 1024, "h", False, out,
                                  RM.dense(self.out_dim))
    return out

  def backward(self, P, G, d_out):
    B = self._inp.B
    a3 = self.trunk.a3
    da3 = self.buf("da3", B, 16, 64)
    # first head layer: dgrad scattered back to (p,c) order and masked by the trunk's ReLU
    keys = self._hk
    acts = self._hacts
    dy, dy_map = d_out, RM.dense(self.out_dim)
    for i in reversed(range(len(keys))):
      wk, bk = keys[i]
      w = P[wk]
      N, K = w.shape
      if i > 0:
        a, a_map, _ = acts[i - 1]
        self.ops.linear_wgrad(dy, dy_map, a, a_map, None, G[wk], G[bk], B, N, K)
        dprev = self.buf("h_d%d" % (i - 1), B, K)
        self.ops.linear_dgrad(dy, dy_map, w, dprev, RM.dense(K), B, N, K, mask=a, mask_map=a_map)
        dy, dy_map = dprev, RM.dense(K)
      else:
        self.ops.linear_wgrad(dy, dy_map, a3, RM.dense(1024), self.kflat, G[wk], G[bk], B, N, K)
        self.ops.linear_dgrad(dy, dy_map, w, da3, RM.dense(1024), B, N, K, mask=a3, mask_map=RM.dense(1024),
                              dx_koff=self.kflat)
    self.trunk.backward(P, G, da3)


def make_plan(family, ops, S, out_dim, **kw):
  if family == "mlp":
    return MLPPlan(ops, S, out_dim)
  if family == "nature":
    return NaturePlan(ops, S, out_dim)
  if family == "loco":
    return LocoPlan(ops, S, out_dim, **kw)
  if family == "vit":
    return LocoPlan(ops, 0, out_dim, has_state=False, **kw)
  if family == "nvo":
    return NatureVOPlan(ops, 0, out_dim)
  raise ValueError("unknown family %r" % (family,))
