from .on_policy import OnPolicyReplayBuffer, OnPolicyReplayBufferBase, BaseReplayBuffer  # noqa: F401
