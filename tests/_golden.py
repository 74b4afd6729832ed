"""Helpers shared by the oracle-vs-golden (CPU) and product-vs-golden/oracle (GPU) tests.

The golden fixtures hold OUTPUTS of the live reference (oracle/make_golden.py); inputs and
weights are regenerated from the same seeds through oracle/synth.py.
"""
import os

import numpy as np

from oracle import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FAMILIES = {"loco": (93, 12), "nature": (84, 6), "mlp": (84, 6), "vit": (0, 6), "nvo": (0, 6)}

INFO_KEYS = ["advs/mean", "advs/std", "advs/max", "advs/min", "Training/vf_loss", "grad_norm/vf",
             "Training/policy_loss", "logprob/mean", "logprob/std", "logprob/max", "logprob/min",
             "log_std/mean", "log_std/std", "log_std/max", "log_std/min", "ratio/max", "ratio/min",
             "grad_norm/pf"]


def load(name):
  return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def family_weights(family):
  S, A = FAMILIES[family]
  return synth.make_family_weights(1000, family, S, A)


def fwd_inputs(family):
  S, A = FAMILIES[family]
  roll = synth.make_rollout(2000, 4, 2, S, A, with_img=family != "mlp")
  return roll["obs"].reshape(8, -1), roll["acts"].reshape(8, -1)


def update_inputs(family):
  S, A = FAMILIES[family]
  T, E, Bm = 4, 4, 16
  roll = synth.make_rollout(3000, T, E, S, A, with_img=family != "mlp", p_term=0.2)
  rng = np.random.default_rng(11)
  return {"obs": roll["obs"].reshape(Bm, -1), "acts": roll["acts"].reshape(Bm, -1),
          "advs": rng.standard_normal((Bm, 1)), "estimate_returns": rng.standard_normal((Bm, 1)),
          "values": roll["values"].reshape(Bm, -1)}


def epoch_inputs(family):
  S, A = FAMILIES[family]
  return synth.make_rollout(4000, 8, 4, S, A, with_img=family != "mlp", p_term=0.15,
                            time_limit_p=0.1)


def gae_case(name, cfg):
  T, E, p_term, p_tl, tlf = cfg
  T, E = int(T), int(E)
  roll = synth.make_rollout(100 + T, T, E, 5, 2, with_img=False, p_term=p_term, time_limit_p=p_tl)
  last_value = np.random.default_rng(7).standard_normal((E, 1))
  return roll, last_value, bool(tlf)


def rel_err(a, b):
  a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
  return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-12))


def check_summary(G, prefix, named, rtol, what=""):
  """named: iterable of (key, numpy array) compared with the fingerprint stored by
  make_golden.summarize (sum / abs-sum / strided sample). Returns the worst relative error."""
  worst = 0.0
  for k, a in named:
    a = np.asarray(a, np.float64).ravel()
    step = max(1, a.size // 61)
    samp = a[::step][:64]
    g_samp = G["%s/%s/sample" % (prefix, k)]
    g_abs = float(G["%s/%s/abs" % (prefix, k)])
    scale = max(g_abs / a.size, 1e-12)
    e1 = float(np.max(np.abs(samp - g_samp))) / max(float(np.max(np.abs(g_samp))), scale)
    e2 = abs(float(np.abs(a).sum()) - g_abs) / max(g_abs, 1e-12)
    worst = max(worst, e1, e2)
    assert e1 <= rtol and e2 <= rtol, "%s %s/%s: sample err %.3e abs-sum err %.3e (tol %.1e)" % (
      what, prefix, k, e1, e2, rtol)
  return worst


def check_info(G, prefix, info, rtol, atol=1e-5, skip=()):
  for k in INFO_KEYS:
    if k in skip:
      continue
    g = float(G["%s/%s" % (prefix, k)])
    v = float(info[k])
    assert abs(v - g) <= atol + rtol * abs(g), "%s/%s: got %.8g want %.8g" % (prefix, k, v, g)
