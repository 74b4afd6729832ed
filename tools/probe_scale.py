"""Experiment: gradient error of the fp16 tier vs the static loss scale (B=1024)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import test_gpu_tc as T
from vision4leg_b200 import engine_tc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for scale in (None, 2 ** 16, 2 ** 20, 2 ** 8):
  engine_tc.LOSS_SCALE_OVERRIDE = scale
  agent, orc, batch, pf, vf = T._tc_agent(B)
  ref = orc.update(batch)
  info = agent.update(batch)
  eng = agent.engine
  vc = min(1.0, 0.5 / (ref["grad_norm/vf"] + 1e-6))
  errs = {k: T.nrm_err(eng.G_vf[k] * vc, gr) for k, gr in orc._last["vgrads"].items()}
  top = sorted(errs.items(), key=lambda kv: -kv[1])
  print("scale", scale, "grad_norm/vf ours %.5f ref %.5f" % (info["grad_norm/vf"], ref["grad_norm/vf"]))
  for k, e in top[:4] + top[-4:]:
    print("   %-60s %.3e" % (k, e))
  if scale is None:
    plan = eng.plan_vf
    for name in ("dout16", "dh2", "dpool", "dxa", "dxb", "dz2", "df1", "dqkv", "da3", "da2", "da1c", "ds"):
      for key, t in plan._ws.items():
        if key[0] == name:
          a = t.float().abs()
          nz = a[a > 0]
          print("      buf %-8s max %.3e median(nz) %.3e frac<6e-5 %.3f" % (name, a.max().item(), nz.median().item() if nz.numel() else 0, (nz < 6e-5).float().mean().item() if nz.numel() else 0))
