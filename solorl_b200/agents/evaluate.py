"""Batched evaluation of a fixed policy — the roll-out loop of the reference's
``testing/test_ppo.py:88-151`` (load ``solo.pt``, run ``--num-runs`` stochastic episodes, report
mean length / reward / success) over N parallel envs instead of one GUI env.

Every env contributes its first ``ceil(num_runs / num_envs)`` episodes, like the reference's single env
contributes its first ``num_runs``.  Returned
records are numpy arrays (one entry per episode): ``episode_return`` (sum of rewards),
``episode_reward`` (last-step reward, what the reference prints), ``episode_length``, ``success``."""
from __future__ import annotations

import numpy as np
import torch

from ..envs import SoloBaseEnv, make_vec_envs
from .policy import Policy


def load_policy(checkpoint, obs_shape, action_space, hidden_size=64, device="cuda"):
    """checkpoint: path or dict with the reference layout {'update','state_dict','ob_rms'}
    (agents/ppo/train.py:124-131)."""
    ckpt = torch.load(checkpoint, map_location=device, weights_only=False) if isinstance(checkpoint, str) else checkpoint
    sd = ckpt["state_dict"]
    hidden = sd["base.features.0.weight"].shape[0] if "base.features.0.weight" in sd else hidden_size
    policy = Policy(obs_shape, action_space, None, {"hidden_size": hidden}).to(device)
    policy.load_state_dict(sd)
    policy.eval()
    return policy, ckpt


def evaluate(policy, config, num_runs=10, num_envs=None, deterministic=False, seed=0, device=None,
             max_steps=None, torch_seed=0):
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    num_envs = int(num_envs or min(num_runs, 4096))
    env = make_vec_envs(config, num_envs, SoloBaseEnv, device=device, training=False, seed=seed)
    sim = env.envs.venv.sim
    g = torch.Generator(device=device).manual_seed(torch_seed)
    # every env contributes its first `quota` episodes (taking episodes in order of completion would
    # over-represent short, i.e. failed, episodes)
    quota = -(-num_runs // num_envs)
    count = np.zeros(num_envs, dtype=np.int64)
    slots = [[] for _ in range(num_envs)]
    obs = env.reset()
    steps = 0
    limit = max_steps or (int(config["episode_length"]) + 1) * quota + 1
    while (count < quota).any() and steps < limit:
        with torch.no_grad():
            if hasattr(policy, "pi_dist"):                     # PPO actor-critic (agents/ppo/policy.py)
                value, feat = policy.base(obs)
                mean, logstd = policy.pi_dist(feat)
                action = mean if deterministic else mean + torch.randn(mean.shape, device=device, generator=g) * logstd.exp()
            else:                                              # a deterministic actor, e.g. TD3 (agents/td3/models.py)
                action = policy(obs)
        obs, reward, done, infos = env.step(action)
        steps += 1
        # one small D2H per step (the done vector); records only when something finished
        d = done.cpu().numpy() > 0.5
        if d.any():
            rec = sim.episode_stats()
            for i in np.nonzero(d & (count < quota))[0]:
                slots[i].append(rec[i].copy())
                count[i] += 1
    env.close()
    flat = [r for k in range(quota) for i in range(num_envs) if len(slots[i]) > k for r in (slots[i][k],)]
    if flat:
        r = np.stack(flat)[:num_runs]
    else:
        r = np.zeros(0, dtype=sim.episode_stats().dtype)
    return {"episode_return": r["episode_return"].astype(np.float64), "episode_reward": r["episode_reward"].astype(np.float64),
            "episode_length": r["episode_length"].astype(np.int64), "success": r["success"].astype(np.int64),
            "steps": steps}


def summarize(res):
    n = max(len(res["episode_return"]), 1)
    return {"episodes": len(res["episode_return"]), "mean_length": float(res["episode_length"].sum() / n),
            "mean_reward": float(res["episode_reward"].sum() / n), "mean_return": float(res["episode_return"].sum() / n),
            "std_return": float(res["episode_return"].std()) if n > 1 else 0.0,
            "mean_success": float(res["success"].sum() / n)}
