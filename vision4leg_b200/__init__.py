"""vision4leg_b200 — the PPO-update hot path of Mehooz/vision4leg on B200 (sm_100a).

Host-side mirror of the reference's torchrl.{algo,networks,policies,replay_buffers} classes
over a C-ABI CUDA library (include/v4l_b200.h).  See DESIGN.md."""
__version__ = "0.1.0"
