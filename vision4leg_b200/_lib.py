"""ctypes binding of libv4l_b200.so (include/v4l_b200.h).

The library is the ONLY compute backend of this package: there is no CPU or PyTorch fallback
for the kernels it exports.  Importing this module never touches the GPU; `ctx()` does and
raises if the library or a B200-class device is missing.
"""
import ctypes as C
import os
import re

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libv4l_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "v4l_b200.h")

RELU, ACCUM = 1, 2
INFO_STRIDE, INFO_COUNT = 32, 18
INFO_KEYS = ["advs/mean", "advs/std", "advs/max", "advs/min", "Training/vf_loss", "grad_norm/vf",
             "Training/policy_loss", "logprob/mean", "logprob/std", "logprob/max", "logprob/min",
             "log_std/mean", "log_std/std", "log_std/max", "log_std/min", "ratio/max", "ratio/min",
             "grad_norm/pf"]
INFO_GRAD_NORM_VF, INFO_GRAD_NORM_PF = 5, 17


class V4LError(RuntimeError):
  pass


class RowMap(C.Structure):
  _fields_ = [("P", C.c_int32), ("item_stride", C.c_int64), ("pos_stride", C.c_int64),
              ("base", C.c_int64), ("idx", C.c_void_p), ("pos_off", C.c_void_p)]


class GemmArgs(C.Structure):
  _fields_ = [("a", C.c_void_p), ("a_map", RowMap), ("a_koff", C.c_void_p),
              ("b", C.c_void_p), ("b_sk", C.c_int64), ("b_sn", C.c_int64),
              ("bias", C.c_void_p),
              ("c", C.c_void_p), ("c_map", RowMap), ("c_koff", C.c_void_p),
              ("mask", C.c_void_p), ("mask_map", RowMap),
              ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("flags", C.c_int32)]


class WgradArgs(C.Structure):
  _fields_ = [("dy", C.c_void_p), ("dy_map", RowMap),
              ("a", C.c_void_p), ("a_map", RowMap), ("a_koff", C.c_void_p),
              ("dw", C.c_void_p), ("ldw", C.c_int64), ("dbias", C.c_void_p),
              ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32)]


class TcGemmArgs(C.Structure):
  _fields_ = [("a", C.c_void_p), ("a_B", C.c_int32), ("a_H", C.c_int32), ("a_W", C.c_int32), ("a_C", C.c_int32),
              ("a_sW", C.c_int64), ("a_sH", C.c_int64), ("a_sB", C.c_int64), ("a_idx", C.c_void_p),
              ("B", C.c_int32), ("Hout", C.c_int32), ("Wout", C.c_int32),
              ("bw", C.c_int32), ("bh", C.c_int32), ("bb", C.c_int32),
              ("n_taps", C.c_int32), ("kchunks", C.c_int32),
              ("tap_dw", C.c_int32 * 16), ("tap_dh", C.c_int32 * 16),
              ("w", C.c_void_p), ("N_pad", C.c_int32), ("N_valid", C.c_int32),
              ("bias", C.c_void_p),
              ("c", C.c_void_p), ("c_map", RowMap), ("c_f32", C.c_int32),
              ("mask", C.c_void_p), ("flags", C.c_int32), ("res", C.c_void_p)]


class TcWgradArgs(C.Structure):
  _fields_ = [("x", C.c_void_p), ("x_B", C.c_int32), ("x_H", C.c_int32), ("x_W", C.c_int32), ("x_C", C.c_int32),
              ("x_sW", C.c_int64), ("x_sH", C.c_int64), ("x_sB", C.c_int64), ("x_estride", C.c_int32),
              ("x_idx", C.c_void_p),
              ("dy", C.c_void_p), ("dy_C", C.c_int32),
              ("dy_sW", C.c_int64), ("dy_sH", C.c_int64), ("dy_sB", C.c_int64),
              ("n_sub", C.c_int32), ("sub_dw", C.c_int32 * 4), ("sub_dh", C.c_int32 * 4),
              ("sub_dyc", C.c_int32 * 4),
              ("B", C.c_int32), ("Hout", C.c_int32), ("Wout", C.c_int32),
              ("bw", C.c_int32), ("bh", C.c_int32), ("bb", C.c_int32),
              ("n_taps", C.c_int32), ("tap_dw", C.c_int32 * 16), ("tap_dh", C.c_int32 * 16),
              ("N_valid", C.c_int32), ("index", C.c_void_p), ("dw", C.c_void_p), ("out_scale", C.c_float),
              ("dbias", C.c_void_p), ("defer", C.c_int32), ("accumulate", C.c_int32)]


class TcMlpLayer(C.Structure):
  _fields_ = [("w", C.c_void_p), ("K", C.c_int32), ("N_pad", C.c_int32), ("N_valid", C.c_int32), ("bias", C.c_void_p),
              ("relu", C.c_int32), ("mask", C.c_void_p), ("mask_ld", C.c_int64), ("out", C.c_void_p),
              ("out_f32", C.c_int32), ("out_map", RowMap)]


class TcMlpChainArgs(C.Structure):
  _fields_ = [("x", C.c_void_p), ("M", C.c_int32), ("x_cols", C.c_int32), ("x_ld", C.c_int64), ("n_layers", C.c_int32),
              ("layer", TcMlpLayer * 3)]


class OptTailArgs(C.Structure):
  _fields_ = [("phases", C.c_int32), ("param", C.c_void_p), ("grad", C.c_void_p), ("m", C.c_void_p),
              ("v", C.c_void_p), ("n", C.c_int64), ("hyper", C.c_void_p), ("info", C.c_void_p),
              ("slot", C.c_void_p), ("norm_slot", C.c_int32), ("extra_lo", C.c_int64), ("extra_n", C.c_int64),
              ("scatter", C.c_void_p), ("packed_self", C.c_void_p), ("packed_other", C.c_void_p),
              ("slot_advance", C.c_void_p)]


class TcConvFlatArgs(C.Structure):
  _fields_ = [("x", C.c_void_p), ("x_rows", C.c_int64), ("C", C.c_int32), ("P", C.c_int32), ("Wg", C.c_int32),
              ("Hout", C.c_int32), ("Wout", C.c_int32), ("n_taps", C.c_int32), ("tap_dw", C.c_int32 * 16),
              ("tap_dh", C.c_int32 * 16), ("w", C.c_void_p), ("N_pad", C.c_int32), ("N_valid", C.c_int32),
              ("bias", C.c_void_p), ("c", C.c_void_p), ("c_map", RowMap), ("n_img", C.c_int64),
              ("x_idx", C.c_void_p), ("flags", C.c_int32), ("mode", C.c_int32)]


class TcBlockArgs(C.Structure):
  _fields_ = [("x", C.c_void_p), ("B", C.c_int32), ("T", C.c_int32), ("eps", C.c_float)] + \
    [(n, C.c_void_p) for n in ("w_in", "w_o", "w_1", "w_2", "b_in", "b_o", "g1", "be1", "b1", "b2", "g2", "be2",
                               "qkv", "o", "h", "f1", "y", "p", "z1", "st1", "z2", "st2", "xh1", "xh2")]


class TcBlockBwdArgs(C.Structure):
  _fields_ = [("dy", C.c_void_p), ("B", C.c_int32), ("T", C.c_int32), ("pad_", C.c_int32)] + \
    [(n, C.c_void_p) for n in ("qkv", "xh1", "xh2", "f1", "p", "st1", "st2", "g1", "g2", "w2d", "w1d", "wod", "wind",
                               "dz2", "df1", "dh", "dz1", "dqkv", "dx")]


_vp, _i, _i64, _f, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t

# name -> argtypes (restype is int unless listed in _RESTYPE)
SIGNATURES = {
  "v4l_version": [],
  "v4l_last_error": [],
  "v4l_ctx_create": [C.POINTER(_vp), _i, _sz],
  "v4l_ctx_destroy": [_vp],
  "v4l_ctx_sm_count": [_vp],
  "v4l_ctx_early_flushes": [_vp],
  "v4l_gemm_rows": [_vp, _vp, C.POINTER(GemmArgs)],
  "v4l_gemm_wgrad": [_vp, _vp, C.POINTER(WgradArgs)],
  "v4l_relu_bwd": [_vp, _vp, _vp, C.POINTER(RowMap), _vp, C.POINTER(RowMap), _vp, C.POINTER(RowMap),
                   _i, _i],
  "v4l_col2im": [_vp, _vp, _vp, _vp, _vp] + [_i] * 9,
  "v4l_attn_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i],
  "v4l_attn_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i],
  "v4l_ln_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f],
  "v4l_ln_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i],
  "v4l_pool_fwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i],
  "v4l_pool_bwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i],
  "v4l_attn_fwd_f16": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i],
  "v4l_attn_bwd_f16": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i],
  "v4l_ln_fwd_f16": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f],
  "v4l_ln_bwd_f16": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f],
  "v4l_pool_fwd_f16": [_vp, _vp, _vp, _vp, _i, _i, _i, _i],
  "v4l_pool_bwd_f16": [_vp, _vp, _vp, _vp, _i, _i, _i, _i],
  "v4l_tc_attn_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i],
  "v4l_tc_conv_flat": [_vp, _vp, C.POINTER(TcConvFlatArgs)],
  "v4l_tc_block_fwd": [_vp, _vp, C.POINTER(TcBlockArgs)],
  "v4l_tc_block_bwd": [_vp, _vp, C.POINTER(TcBlockBwdArgs)],
  "v4l_tc_block_timeline": [C.POINTER(C.c_uint64)],
  "v4l_tc_attn_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i],
  "v4l_gae": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _i, _i, _d, _d, _i, _i],
  "v4l_select_rows": [_vp, _vp, _vp, _vp, _vp, _i],
  "v4l_slot_advance": [_vp, _vp, _vp, C.c_int32],
  "v4l_adv_stats": [_vp, _vp, _vp, _vp, _i, _vp],
  "v4l_vf_loss": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _f, _f, _i, _f, _vp, _vp, _vp, _f],
  "v4l_pf_loss": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _f,
                  _f, _vp, _vp, _i, _vp, _f, _i],
  "v4l_adv_stats_epoch": [_vp, _vp, _vp, _i, _i, _vp, _vp],
  "v4l_clip_adam": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i],
  "v4l_tc_gemm": [_vp, _vp, C.POINTER(TcGemmArgs)],
  "v4l_tc_wgrad": [_vp, _vp, C.POINTER(TcWgradArgs)],
  "v4l_tc_wgrad_flush": [_vp, _vp],
  "v4l_tc_wgrad_conv1": [_vp, _vp, _vp, _i64, _vp, _vp, _i, _vp, _vp, _vp, _f, _i, _i],
  "v4l_tc_mlp_chain": [_vp, _vp, C.POINTER(TcMlpChainArgs)],
  "v4l_depth_frame": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _f],
  "v4l_stack_frames": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _i64],
  "v4l_normalizer": [_vp, _vp, _vp, _i, _i, _vp, _vp, C.c_double, _i, _f, _vp],
  "v4l_opt_tail": [_vp, _vp, C.POINTER(OptTailArgs)],
  "v4l_opt_tail_error": [_vp],
  "v4l_mb_begin": [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _i],
  "v4l_colsum_f16": [_vp, _vp, _vp, C.POINTER(RowMap), _i, _i, _i, _f, _vp],
  "v4l_ingest_img": [_vp, _vp, _vp, _vp, _i64, _vp],
  "v4l_ingest_img_f16": [_vp, _vp, _vp, _vp, _i64, _vp],
  "v4l_ingest_rows": [_vp, _vp, _vp, _i64, _i, _vp, _i64, _vp, _vp, _vp],
  "v4l_gather_rows_f16": [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i64, _i, _f],
  "v4l_relu_bwd_f16": [_vp, _vp, _vp, C.POINTER(RowMap), _vp, C.POINTER(RowMap), _vp, C.POINTER(RowMap),
                        _i, _i],
  "v4l_pack_f16": [_vp, _vp, _vp, _vp, _vp, _i64],
  "v4l_h2d_2d": [_vp, _vp, _sz, _vp, _sz, _sz, _sz],
  "v4l_h2d_rows": [_vp, _vp, _vp, _vp, _i, _sz],
}
_RESTYPE = {"v4l_last_error": C.c_char_p}


def header_symbols():
  """Every function the public header declares (used by the ABI test)."""
  text = open(HEADER).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(v4l_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load():
  """dlopen the library and bind every symbol (no GPU needed)."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise V4LError("%s is missing: build it with `python -m vision4leg_b200.build` "
                   "(there is no CPU fallback for this path)" % LIB_PATH)
  lib = C.CDLL(LIB_PATH)
  for name, argtypes in SIGNATURES.items():
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = _RESTYPE.get(name, C.c_int)
  if lib.v4l_version() != 1:
    raise V4LError("ABI version mismatch: library %d, binding 1" % lib.v4l_version())
  _lib = lib
  return lib


def check(rc):
  if rc != 0:
    raise V4LError("libv4l_b200: %s (rc=%d)" % (load().v4l_last_error().decode(), rc))


class Context:
  """One per device per process (v4l_ctx)."""

  def __init__(self, device, scratch_bytes=0):
    self.lib = load()
    self.device = torch.device(device)
    if self.device.type != "cuda":
      raise V4LError("vision4leg_b200 runs on CUDA devices only (got %s); there is no CPU path"
                     % (self.device,))
    index = self.device.index if self.device.index is not None else torch.cuda.current_device()
    self.device = torch.device("cuda", index)
    h = _vp()
    with torch.cuda.device(index):
      torch.cuda.current_stream()          # make sure torch has initialised the context
      check(self.lib.v4l_ctx_create(C.byref(h), index, scratch_bytes))
    self.handle = h
    self.sm_count = self.lib.v4l_ctx_sm_count(h)

  def stream(self):
    return _vp(torch.cuda.current_stream(self.device).cuda_stream)

  def __del__(self):
    try:
      if getattr(self, "handle", None):
        self.lib.v4l_ctx_destroy(self.handle)
        self.handle = None
    except Exception:
      pass


_contexts = {}


def ctx(device):
  device = torch.device(device)
  if device.type != "cuda":
    raise V4LError("vision4leg_b200 runs on CUDA devices only (got %s); there is no CPU path"
                   % (device,))
  index = device.index if device.index is not None else torch.cuda.current_device()
  c = _contexts.get(index)
  if c is None:
    c = _contexts[index] = Context(torch.device("cuda", index))
  return c


def ptr(t):
  return None if t is None else t.data_ptr()
