from vision4leg_b200.replay_buffers.on_policy import BaseReplayBuffer   # noqa: F401
