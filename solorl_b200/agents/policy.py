"""Actor-critic with the architecture and ``state_dict`` layout of the reference's
``agents/ppo/policy.py`` for 1-D observations and Box actions (``Policy`` :10-58, ``MLP`` :60-81,
``DiagGaussian`` :139-149): separate tanh MLPs for actor features and critic, a linear mean head
and a state-independent log-std.  Checkpoints written by the reference
(``{'update','state_dict','ob_rms'}``, agents/ppo/train.py:124-131) load unchanged; parameter
counts are pinned against the reference in tests/golden/policy_shapes.json."""
import math

import torch
import torch.nn as nn

from .utils import init_layer

_LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)


class MLP(nn.Module):
    def __init__(self, num_inputs, hidden_size=64):
        super().__init__()
        self.output_size = hidden_size
        self.features = nn.Sequential(
            init_layer(nn.Linear(num_inputs, hidden_size)), nn.Tanh(),
            init_layer(nn.Linear(hidden_size, hidden_size)), nn.Tanh())
        self.critic = nn.Sequential(
            init_layer(nn.Linear(num_inputs, hidden_size)), nn.Tanh(),
            init_layer(nn.Linear(hidden_size, hidden_size)), nn.Tanh(),
            init_layer(nn.Linear(hidden_size, 1)))

    def forward(self, x):
        return self.critic(x), self.features(x)


class DiagGaussian(nn.Module):
    """Mean = linear(features), std = exp(logstd) independent of the state."""

    def __init__(self, num_inputs, num_outputs):
        super().__init__()
        self.mean = init_layer(nn.Linear(num_inputs, num_outputs))
        self.logstd = nn.Parameter(torch.zeros(num_outputs))

    def forward(self, x):
        return self.mean(x), self.logstd


class Policy(nn.Module):
    def __init__(self, obs_shape, action_space, base=None, base_kwargs=None):
        super().__init__()
        if len(obs_shape) != 1:
            raise NotImplementedError("only 1-D observations (the transformer base serves the gait envs)")
        if action_space.__class__.__name__ != "Box":
            raise NotImplementedError("only Box action spaces (SoloBaseEnv)")
        self.base = MLP(obs_shape[0], **(base_kwargs or {}))
        if base is not None:
            self.base.load_state_dict(base)              # --base-checkpoint, agents/ppo/train.py:44
        self.pi_dist = DiagGaussian(self.base.output_size, action_space.shape[0])

    @staticmethod
    def _log_prob(action, mean, logstd):
        z = (action - mean) * torch.exp(-logstd)
        return (-0.5 * z * z - logstd - _LOG_SQRT_2PI).sum(-1, keepdim=True)

    def act(self, inputs, deterministic=False):
        value, feat = self.base(inputs)
        mean, logstd = self.pi_dist(feat)
        action = mean if deterministic else mean + torch.randn_like(mean) * torch.exp(logstd)
        return value, action, self._log_prob(action, mean, logstd)

    def get_value(self, inputs):
        return self.base(inputs)[0]

    def evaluate_actions(self, inputs, action):
        value, feat = self.base(inputs)
        mean, logstd = self.pi_dist(feat)
        # Normal.entropy() per dimension, averaged over batch AND action dims (policy.py:55)
        entropy = (0.5 + _LOG_SQRT_2PI + logstd).expand_as(mean).mean()
        return value, self._log_prob(action, mean, logstd), entropy
