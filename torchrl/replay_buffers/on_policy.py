from vision4leg_b200.replay_buffers.on_policy import (OnPolicyReplayBuffer,   # noqa: F401
                                                      OnPolicyReplayBufferBase, BaseReplayBuffer)
