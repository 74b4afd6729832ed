"""Drop-in shim: the reference's `torchrl` import paths, with the PPO-update hot path served by
vision4leg_b200 (CUDA, sm_100a) and everything else falling through to a reference checkout.

Put this repo BEFORE the reference on sys.path (PYTHONPATH=/path/to/this/repo:/path/to/vision4leg);
`starter/ppo_*.py` then run unchanged.  The reference root is found through $V4L_REFERENCE_ROOT or
any later sys.path entry that contains a `torchrl/` directory; sub-packages we do not re-author
(torchrl.collector, torchrl.env, torchrl.utils, torchrl.algo.off_policy, ...) resolve there.
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))


def _reference_roots():
  roots = []
  env = os.environ.get("V4L_REFERENCE_ROOT")
  if env:
    roots.append(env)
  roots += [p for p in sys.path if p]
  out = []
  for r in roots:
    cand = os.path.join(os.path.abspath(r), "torchrl")
    if os.path.isdir(cand) and os.path.abspath(cand) != _here and cand not in out:
      out.append(cand)
  return out


def _extend(path_list, *sub):
  for root in _reference_roots():
    cand = os.path.join(root, *sub)
    if os.path.isdir(cand) and cand not in path_list:
      path_list.append(cand)


_extend(__path__)
