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


class CategoricalHead(nn.Module):
    """The head ``agents/ppo/policy.py:22-23`` names for ``Discrete`` action spaces but never defines:
    written as the reference writes its ``BernoulliHead`` (:158-164), a single ``init_layer``-ed linear map to the
    logits of the ``ModCategorical`` distribution that IS defined there (:178-187)."""

    def __init__(self, num_inputs, num_outputs):
        super().__init__()
        self.linear = init_layer(nn.Linear(num_inputs, num_outputs))

    def forward(self, x):
        return torch.log_softmax(self.linear(x), dim=-1)


class Policy(nn.Module):
    def __init__(self, obs_shape, action_space, base=None, base_kwargs=None):
        super().__init__()
        if len(obs_shape) != 1:
            raise NotImplementedError("only 1-D observations (the transformer base serves the timing envs)")
        kind = action_space.__class__.__name__
        if kind not in ("Box", "Discrete"):
            raise NotImplementedError("Box (SoloBaseEnv) and Discrete (SoloGaitEnvContact) action spaces are built")
        self.base = MLP(obs_shape[0], **(base_kwargs or {}))
        if base is not None:
            self.base.load_state_dict(base)              # --base-checkpoint, agents/ppo/train.py:44
        self.discrete = kind == "Discrete"
        if self.discrete:
            self.pi_dist = CategoricalHead(self.base.output_size, action_space.n)
        else:
            self.pi_dist = DiagGaussian(self.base.output_size, action_space.shape[0])

    @staticmethod
    def _log_prob(action, mean, logstd):
        z = (action - mean) * torch.exp(-logstd)
        return (-0.5 * z * z - logstd - _LOG_SQRT_2PI).sum(-1, keepdim=True)

    # ---- Discrete(n): ModCategorical semantics (policy.py:178-187): actions [N,1], log-probs [N,1] ------------
    def _act_discrete(self, value, feat, deterministic):
        logp = self.pi_dist(feat)
        if deterministic:
            action = logp.argmax(dim=-1, keepdim=True)                       # mode()
        else:                                                               # Gumbel-max: CUDA-graph capturable
            u = torch.rand_like(logp).clamp_(1e-20, 1.0)
            action = (logp - torch.log(-torch.log(u))).argmax(dim=-1, keepdim=True)
        return value, action.float(), logp.gather(-1, action)

    def _evaluate_discrete(self, value, feat, action):
        logp = self.pi_dist(feat)
        entropy = -(logp.exp() * logp).sum(-1).mean()
        return value, logp.gather(-1, action.long().reshape(-1, 1)), entropy

    def act(self, inputs, deterministic=False):
        value, feat = self.base(inputs)
        if self.discrete:
            return self._act_discrete(value, feat, deterministic)
        mean, logstd = self.pi_dist(feat)
        action = mean if deterministic else mean + torch.randn_like(mean) * torch.exp(logstd)
        return value, action, self._log_prob(action, mean, logstd)

    def get_value(self, inputs):
        return self.base(inputs)[0]

    def evaluate_actions(self, inputs, action):
        value, feat = self.base(inputs)
        if self.discrete:
            return self._evaluate_discrete(value, feat, action)
        mean, logstd = self.pi_dist(feat)
        # Normal.entropy() per dimension, averaged over batch AND action dims (policy.py:55)
        entropy = (0.5 + _LOG_SQRT_2PI + logstd).expand_as(mean).mean()
        return value, self._log_prob(action, mean, logstd), entropy
