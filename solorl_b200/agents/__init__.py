"""Trainer-side callers of the env-step hot path (PPO), kept in PyTorch: the policy is a
19k-parameter MLP, not a kernel target.  What changes against the reference's agents/ppo is the
data path: the rollout never leaves the device, GAE is one reverse-scan kernel, and gradients
are all-reduced across GPUs (one process per GPU, independent env shards)."""
