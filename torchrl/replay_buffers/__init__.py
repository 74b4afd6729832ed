from torchrl import _extend as _ext
_ext(__path__, "replay_buffers")
from vision4leg_b200.replay_buffers import BaseReplayBuffer, OnPolicyReplayBuffer   # noqa: E402,F401
