"""Device-resident rollout buffer, the replacement of ``OPBuffer`` (agents/ppo/storage.py:4-71).

Same tensors and shapes — obs [T+1,N,D], rewards/log-probs [T,N,1], values/returns/masks
[T+1,N,1], actions [T,N,A] — but allocated once on the GPU, filled by device-to-device copies,
and ``compute_returns`` is one launch of the reverse-scan GAE kernel (``solo_gae``) instead of a
Python loop of T tiny torch ops (storage.py:41-50)."""
import torch

from ..sim import gae as gae_kernel


class OPBuffer:
    def __init__(self, num_steps, num_agents, obs_shape, action_dim, device):
        f = dict(dtype=torch.float32, device=device)
        self.obs = torch.zeros(num_steps + 1, num_agents, *obs_shape, **f)
        self.rewards = torch.zeros(num_steps, num_agents, 1, **f)
        self.value_preds = torch.zeros(num_steps + 1, num_agents, 1, **f)
        self.returns = torch.zeros(num_steps + 1, num_agents, 1, **f)
        self.action_log_probs = torch.zeros(num_steps, num_agents, 1, **f)
        self.actions = torch.zeros(num_steps, num_agents, action_dim, **f)
        self.masks = torch.ones(num_steps + 1, num_agents, 1, **f)
        self.num_samples = num_steps * num_agents
        self.num_steps = num_steps
        self.device = torch.device(device)
        self.step = 0

    def append(self, obs, actions, action_log_probs, value_preds, rewards, masks):
        t = self.step
        self.obs[t + 1].copy_(obs)
        self.actions[t].copy_(actions)
        self.action_log_probs[t].copy_(action_log_probs)
        self.value_preds[t].copy_(value_preds)
        self.rewards[t].copy_(rewards)
        self.masks[t + 1].copy_(masks)
        self.step = (t + 1) % self.num_steps

    def reset(self):
        """Carry the last observation / mask into slot 0 (storage.py:31-33)."""
        self.obs[0].copy_(self.obs[-1])
        self.masks[0].copy_(self.masks[-1])

    def compute_returns(self, next_value, use_gae=True, gamma=0.99, gae_lambda=0.95):
        if self.device.type != "cuda":
            raise RuntimeError("OPBuffer.compute_returns runs the CUDA GAE kernel (no CPU fallback)")
        if use_gae:
            self.value_preds[-1].copy_(next_value)
        else:
            self.returns[-1].copy_(next_value)
        gae_kernel(self.rewards, self.value_preds, self.masks, self.returns, gamma, gae_lambda, use_gae)

    def batch_generator(self, advantages, mini_batch_size, generator=None):
        """Random mini-batches without replacement, last partial batch dropped (storage.py:57-71)."""
        perm = torch.randperm(self.num_samples, device=self.device, generator=generator)
        obs = self.obs[:-1].reshape(self.num_samples, *self.obs.shape[2:])
        actions = self.actions.reshape(self.num_samples, -1)
        values = self.value_preds[:-1].reshape(-1, 1)
        returns = self.returns[:-1].reshape(-1, 1)
        masks = self.masks[:-1].reshape(-1, 1)
        logp = self.action_log_probs.reshape(-1, 1)
        adv = advantages.reshape(-1, 1)
        for start in range(0, self.num_samples - mini_batch_size + 1, mini_batch_size):
            idx = perm[start:start + mini_batch_size]
            yield obs[idx], actions[idx], values[idx], returns[idx], masks[idx], logp[idx], adv[idx]
