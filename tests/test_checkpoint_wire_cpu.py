"""Checkpoint wire format: what RLAlgo.snapshot writes must stay loadable by the untouched
reference (viewers, TensorRT export): `model_{pf,vf}_{epoch}.pth` = plain state_dicts with the reference's keys,
`_obs_normalizer_{epoch}.pkl` = a pickled torchrl.env.base_wrapper.Normalizer (reference rl_algo.py:84-95).

The fixture tests/golden/obs_normalizer_ref.pkl was written by the live reference (oracle/make_golden_obs.py).
Two tests load OUR files into the LIVE reference in a subprocess; they need /root/reference and are skipped
where it does not exist (the GPU box)."""
import io
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests import _golden as g
from vision4leg_b200 import obs_pipeline as op

REF = os.environ.get("V4L_REFERENCE_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "torchrl")), reason="no reference checkout here")


def _globals(stream):
  """(module, name) of every class / function a pickle names, as the unpickler resolves them"""
  seen = []

  class Rec(op._WireUnpickler):
    def find_class(self, module, name):
      seen.append((module, name))
      return super().find_class(module, name)
  Rec(io.BytesIO(stream)).load()
  return seen


class _DeviceStandIn:
  """the device Normalizer's pickling hook without a device: same __reduce__, statistics from the fixture"""
  __reduce__ = op.Normalizer.__reduce__

  def __init__(self, st):
    self.st = st

  def to_reference(self):
    return op.reference_normalizer_object(**{k.lstrip("_") if k in ("_mean", "_var", "_count") else k: v
                                             for k, v in self.st.items()})


def test_reference_pickle_fixture_loads_without_the_reference():
  G = np.load(os.path.join(g.GOLDEN_DIR, "obs_normalizer.npz"))
  with open(os.path.join(g.GOLDEN_DIR, "obs_normalizer_ref.pkl"), "rb") as f:
    st = op.load_reference_normalizer(f)
  assert list(st.keys()) == ["shape", "_mean", "_var", "_count", "clip", "should_estimate"]
  assert np.array_equal(st["_mean"], G["mean"]) and np.array_equal(st["_var"], G["var"])
  assert st["_count"] == float(G["count"]) and st["clip"] == 10.0 and st["should_estimate"] is False


def test_our_pickle_names_only_the_reference_class():
  with open(os.path.join(g.GOLDEN_DIR, "obs_normalizer_ref.pkl"), "rb") as f:
    raw = f.read()
  st = op.load_reference_normalizer(io.BytesIO(raw))
  for obj in (op.reference_normalizer_object(st["shape"], st["_mean"], st["_var"], st["_count"], st["clip"],
                                             st["should_estimate"]), _DeviceStandIn(st)):
    mine = pickle.dumps(obj)
    assert b"vision4leg_b200" not in mine
    names = _globals(mine)
    assert op.REFERENCE_NORMALIZER in names
    assert {m.split(".")[0] for m, _ in names} <= {"torchrl", "numpy", "copyreg", "builtins"}, names
    back = op.load_reference_normalizer(io.BytesIO(mine))
    assert list(back.keys()) == list(st.keys())
    for k in st:
      assert np.array_equal(back[k], st[k]), k
  # the plain object pickles to the very bytes the reference wrote (same class path, attribute order, protocol)
  assert pickle.dumps(op.reference_normalizer_object(st["shape"], st["_mean"], st["_var"], st["_count"], st["clip"],
                                                     st["should_estimate"])) == raw


def _run_in_reference(code, *argv):
  env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=ROOT)
  r = subprocess.run([sys.executable, "-c", code] + list(argv), env=env, capture_output=True, text=True, timeout=300)
  assert r.returncode == 0, r.stdout + r.stderr
  return r.stdout


@needs_reference
def test_live_reference_loads_our_normalizer_pickle(tmp_path):
  rng = np.random.RandomState(3)
  mean, var = rng.randn(37), rng.rand(37) + 0.5
  path = str(tmp_path / "_obs_normalizer_7.pkl")
  with open(path, "wb") as f:
    pickle.dump(_DeviceStandIn(dict(shape=(37,), _mean=mean, _var=var, _count=123.0001, clip=10., should_estimate=True)), f)
  x = rng.randn(4, 37)
  np.save(str(tmp_path / "x.npy"), x)
  out = _run_in_reference("""
import sys, pickle, types, importlib.util, numpy as np
gym = types.ModuleType("gym")
for n in ("Wrapper", "RewardWrapper", "ObservationWrapper"):
  setattr(gym, n, type(n, (), {}))
sys.modules["gym"] = gym
spec = importlib.util.spec_from_file_location("torchrl.env.base_wrapper", sys.argv[1] + "/torchrl/env/base_wrapper.py")
mod = importlib.util.module_from_spec(spec); sys.modules["torchrl.env.base_wrapper"] = mod; spec.loader.exec_module(mod)
nz = pickle.load(open(sys.argv[2], "rb"))
assert type(nz) is mod.Normalizer, type(nz)
np.save(sys.argv[3], nz.filt(np.load(sys.argv[4])))
print(nz._count)
""", REF, path, str(tmp_path / "y.npy"), str(tmp_path / "x.npy"))
  assert float(out.strip()) == 123.0001
  want = np.clip((x - mean) / (np.sqrt(var) + 1e-4), -10, 10)
  assert np.array_equal(np.load(str(tmp_path / "y.npy")), want)


@needs_reference
@pytest.mark.parametrize("family", ["loco", "mlp"])
def test_live_reference_loads_our_state_dicts(tmp_path, family):
  """model_{pf,vf}_{epoch}.pth written the way RLAlgo.snapshot writes them load into the reference's own modules
  with strict=True and reproduce the reference's forward fixture"""
  from benchutil.harness import build_nets, load_np_sd
  S, A = g.FAMILIES[family]
  pf, vf = build_nets(family, S, A)
  pf_np, vf_np = g.family_weights(family)
  load_np_sd(pf, pf_np); load_np_sd(vf, vf_np)
  for name, net in (("pf", pf), ("vf", vf)):
    torch.save({k: v.detach().clone() for k, v in net.state_dict().items()}, str(tmp_path / ("model_%s_3.pth" % name)))
  _run_in_reference("""
import sys, numpy as np, torch
from oracle import make_golden as mg
from tests import _golden as g
family, d = sys.argv[1], sys.argv[2]
obs, acts = g.fwd_inputs(family)
G = g.load(family)
networks, policies, PPO, Buffer, Box = mg.import_reference()
S, A = mg.FAMILIES[family]
pf, vf = mg.build_reference_nets(networks, policies, family, S, A)
pf.load_state_dict(torch.load(d + "/model_pf_3.pth", map_location="cpu"), strict=True)
vf.load_state_dict(torch.load(d + "/model_vf_3.pth", map_location="cpu"), strict=True)
with torch.no_grad():
  assert np.allclose(vf(torch.tensor(obs)).numpy(), G["fwd/value"], rtol=1e-5, atol=1e-6)
  assert np.allclose(pf.update(torch.tensor(obs), torch.tensor(acts))["mean"].numpy(), G["fwd/mean"], rtol=1e-5, atol=1e-6)
""", family, str(tmp_path))
