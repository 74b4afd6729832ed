"""Deterministic synthetic rollouts and weights now live in benchutil/synth.py (a neutral package: bench.py's GPU
arm and the tools use them without importing anything under oracle/); this module re-exports them for the
oracle-side tests and the golden-vector generator."""
from benchutil.synth import *          # noqa: F401,F403
from benchutil.synth import (make_obs, make_rollout, make_weights, make_family_weights, IMG_ELEMS,   # noqa: F401
                             loco_encoder_spec, loco_net_spec, nature_encoder_spec, nature_net_spec,
                             mlp_base_spec, mlp_net_spec, vit_encoder_spec, vit_net_spec)
