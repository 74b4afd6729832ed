"""CPU checks of the boundary: the C-ABI library loads, exports every symbol the public header
declares, and the ctypes structures match the header's layout.  No compute calls (no GPU)."""
import ctypes as C
import os
import re

import pytest

from vision4leg_b200 import _lib


def test_library_exports_every_declared_symbol():
  lib = _lib.load()
  declared = _lib.header_symbols()
  assert len(declared) >= 20
  missing = [s for s in declared if not hasattr(lib, s)]
  assert not missing, missing
  assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)
  assert lib.v4l_version() == 1


def test_struct_layouts_match_header():
  # v4l_rowmap: int32 (+pad), 3 x int64, 2 pointers = 48 bytes on LP64
  assert C.sizeof(_lib.RowMap) == 48
  assert _lib.GemmArgs.c_map.offset - _lib.GemmArgs.c.offset == 8
  assert C.sizeof(_lib.GemmArgs) == 8 + 48 + 8 + 8 + 16 + 8 + 8 + 48 + 8 + 8 + 48 + 16
  assert C.sizeof(_lib.WgradArgs) == 8 + 48 + 8 + 48 + 8 + 8 + 8 + 8 + 16


def test_info_layout_matches_header():
  text = open(_lib.HEADER).read()
  enum = re.search(r"V4L_INFO_ADV_MEAN = 0,(.*?)V4L_INFO_COUNT = (\d+), V4L_INFO_STRIDE = (\d+)", text, re.S)
  names = ["V4L_INFO_ADV_MEAN"] + re.findall(r"(V4L_INFO_[A-Z_]+)", enum.group(1))
  assert len(names) == int(enum.group(2)) == len(_lib.INFO_KEYS) == _lib.INFO_COUNT
  assert int(enum.group(3)) == _lib.INFO_STRIDE
  assert names.index("V4L_INFO_GRAD_NORM_VF") == _lib.INFO_GRAD_NORM_VF
  assert names.index("V4L_INFO_GRAD_NORM_PF") == _lib.INFO_GRAD_NORM_PF


def test_no_cpu_fallback():
  """The product must fail loudly without a CUDA device instead of computing on the CPU."""
  import torch
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  with pytest.raises(_lib.V4LError):
    _lib.ctx("cpu")
  from tests._harness import build_nets
  pf, vf = build_nets("mlp", 8, 2)
  with pytest.raises(_lib.V4LError):
    vf(torch.zeros(3, 8))


def test_product_never_imports_oracle():
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  bad = []
  for d, _, files in os.walk(os.path.join(root, "vision4leg_b200")):
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".h")):
        src = open(os.path.join(d, f)).read()
        if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M) or "oracle/" in src:
          bad.append(os.path.join(d, f))
  assert not bad, bad
