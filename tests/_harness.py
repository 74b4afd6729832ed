"""Builders shared by the GPU tests, smoke() and bench.py live in benchutil/harness.py; re-exported here."""
from benchutil.harness import *        # noqa: F401,F403
from benchutil.harness import build_nets, load_np_sd, fill_buffer, make_ppo, Obj, Box, ListLogger   # noqa: F401
