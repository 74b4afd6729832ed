import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
  try:
    import torch
    # the torch restatements the kernels are checked against must be true fp32 (cuDNN convs
    # default to TF32, ~1e-3 relative error — the size of the tolerance being tested)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
  except Exception:
    pass


def pytest_collection_modifyitems(config, items):
  try:
    import torch
    has_gpu = torch.cuda.is_available()
  except Exception:
    has_gpu = False
  if has_gpu:
    return
  skip = pytest.mark.skip(reason="no CUDA device")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)
