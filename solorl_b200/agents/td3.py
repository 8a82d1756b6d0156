"""TD3 against the GPU vec-env — the caller named by the north_star next to PPO
(reference ``agents/td3/{models,td3,buffer,train}.py``, ``training/train_td3.py``).

The learner is Fujimoto's TD3 exactly as the reference has it (256-256 ReLU actor with tanh head,
twin 256-256 critics, target smoothing, delayed actor/target updates, Adam 3e-4), with the layer names
``l1..l6`` kept so reference checkpoints (``{'update','state_dict','critic_state_dict'}``,
agents/td3/train.py:142-154) load.  What changes is the data path: the replay buffer is a
device-resident ring filled by ONE batched write per vec-env step instead of a Python loop over
transitions on CPU tensors (buffer.py:25-39), sampling indexes it on the device, and episode statistics
come from the device-side accumulators instead of ``done[i].item()`` per env (train.py:108-115)."""
from __future__ import annotations

import copy
import os
import time
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..envs import SoloBaseEnv, make_vec_envs
from . import utils
from .train import EpisodeTracker, _DR


class Actor(nn.Module):
    def __init__(self, state_dim, action_dim):
        super().__init__()
        self.l1 = nn.Linear(state_dim, 256)
        self.l2 = nn.Linear(256, 256)
        self.l3 = nn.Linear(256, action_dim)

    def forward(self, state):
        return torch.tanh(self.l3(F.relu(self.l2(F.relu(self.l1(state))))))


class Critic(nn.Module):
    def __init__(self, state_dim, action_dim):
        super().__init__()
        self.l1 = nn.Linear(state_dim + action_dim, 256)
        self.l2 = nn.Linear(256, 256)
        self.l3 = nn.Linear(256, 1)
        self.l4 = nn.Linear(state_dim + action_dim, 256)
        self.l5 = nn.Linear(256, 256)
        self.l6 = nn.Linear(256, 1)

    def forward(self, state, action):
        sa = torch.cat([state, action], 1)
        q1 = self.l3(F.relu(self.l2(F.relu(self.l1(sa)))))
        q2 = self.l6(F.relu(self.l5(F.relu(self.l4(sa)))))
        return q1, q2

    def Q1(self, state, action):
        sa = torch.cat([state, action], 1)
        return self.l3(F.relu(self.l2(F.relu(self.l1(sa)))))


class ReplayBuffer:
    """Ring of transitions on the device (agents/td3/buffer.py:10-52 kept the tensors on the CPU)."""

    def __init__(self, max_size, observation_dim, action_dim, device):
        self.device = torch.device(device)
        self._max_size = int(max_size)
        f = dict(dtype=torch.float32, device=self.device)
        self._observations = torch.zeros(self._max_size, observation_dim, **f)
        self._actions = torch.zeros(self._max_size, action_dim, **f)
        self._rewards = torch.zeros(self._max_size, 1, **f)
        self._next_observations = torch.zeros(self._max_size, observation_dim, **f)
        self._not_terminal = torch.ones(self._max_size, 1, **f)
        self._size = 0
        self._top = 0

    def __len__(self):
        return self._size

    def append_batch(self, obs, action, reward, next_obs, not_terminal):
        """N transitions in one write (the ring wraps with a second slice, no per-transition loop)."""
        n = obs.shape[0]
        if n > self._max_size:
            raise ValueError("batch larger than the replay buffer")
        first = min(n, self._max_size - self._top)
        for dst, src in ((self._observations, obs), (self._actions, action), (self._rewards, reward.reshape(n, 1)),
                         (self._next_observations, next_obs), (self._not_terminal, not_terminal.reshape(n, 1))):
            dst[self._top:self._top + first].copy_(src[:first])
            if first < n:
                dst[:n - first].copy_(src[first:])
        self._top = (self._top + n) % self._max_size
        self._size = min(self._size + n, self._max_size)

    def append(self, transitions):
        """The reference's per-transition interface (buffer.py:25-27), kept for drop-in use."""
        for s, a, r, s2, nt in transitions:
            self.append_batch(torch.as_tensor(s, device=self.device).reshape(1, -1),
                              torch.as_tensor(a, device=self.device).reshape(1, -1),
                              torch.as_tensor(r, device=self.device).reshape(1, 1),
                              torch.as_tensor(s2, device=self.device).reshape(1, -1),
                              torch.as_tensor(nt, device=self.device).reshape(1, 1))

    def sample(self, mini_batch_size, generator=None):
        idx = torch.randint(0, self._size, (mini_batch_size,), device=self.device, generator=generator)
        return self.gather(idx)

    def gather(self, idx):
        return (self._observations[idx], self._actions[idx], self._rewards[idx], self._next_observations[idx],
                self._not_terminal[idx])


class TD3:
    def __init__(self, obs_dim, action_dim, gamma=0.99, tau=0.005, policy_noise=0.2, noise_clip=0.5, policy_freq=2,
                 device="cuda", use_graph=True):
        self.actor = Actor(obs_dim, action_dim).to(device)
        self.actor_target = copy.deepcopy(self.actor)
        self.critic = Critic(obs_dim, action_dim).to(device)
        self.critic_target = copy.deepcopy(self.critic)
        cuda = torch.device(device).type == "cuda"
        kw = dict(fused=True, capturable=True) if cuda else {}
        self.actor_optimizer = torch.optim.Adam(self.actor.parameters(), lr=3e-4, **kw)
        self.critic_optimizer = torch.optim.Adam(self.critic.parameters(), lr=3e-4, **kw)
        # the update is ~150 small launches; on CUDA it is captured once per kind of step (critic only /
        # critic + delayed actor and targets) and replayed with a fresh index tensor
        self.use_graph = bool(use_graph and cuda)
        self._graphs, self._warm, self._idx = {}, {}, None
        self.gamma, self.tau = float(gamma), float(tau)
        self.policy_noise, self.noise_clip, self.policy_freq = float(policy_noise), float(noise_clip), int(policy_freq)

    def select_action(self, obs):
        return self.actor(obs)

    def train(self, replay_buffer, step, batch_size=100):
        """One TD3 update (td3.py:41-91); returns loss TENSORS (no host sync here)."""
        if not self.use_graph:
            return self._update(replay_buffer.sample(batch_size), step % self.policy_freq == 0)
        with_actor = (step % self.policy_freq == 0)
        if self._idx is None or self._idx.numel() != batch_size:
            self._idx = torch.zeros(batch_size, dtype=torch.long, device=replay_buffer.device)
            self._out = {k: [torch.zeros((), device=replay_buffer.device), torch.zeros((), device=replay_buffer.device)]
                         for k in (False, True)}
            self._graphs, self._warm = {}, {}
        self._idx.copy_(torch.randint(0, len(replay_buffer), (batch_size,), device=replay_buffer.device))

        def body():
            q, a = self._update(replay_buffer.gather(self._idx), with_actor)
            self._out[with_actor][0].copy_(q)
            if a is not None:
                self._out[with_actor][1].copy_(a)

        g = self._graphs.get(with_actor)
        if g is not None:
            g.replay()
        elif self._warm.get(with_actor, 0) < 3:             # eager warm-up (real updates) on a side stream
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                body()
            torch.cuda.current_stream().wait_stream(s)
            self._warm[with_actor] = self._warm.get(with_actor, 0) + 1
        else:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    body()
                self._graphs[with_actor] = g
                g.replay()
            except Exception as e:                           # pragma: no cover - capture is best effort
                print(f"[td3] CUDA graph capture failed ({e}); continuing eagerly", flush=True)
                self.use_graph = False
                body()
        out = self._out[with_actor]
        return out[0], (out[1] if with_actor else None)

    def _update(self, batch, with_actor):
        obs, action, reward, next_obs, not_done = batch
        with torch.no_grad():
            noise = (torch.randn_like(action) * self.policy_noise).clamp(-self.noise_clip, self.noise_clip)
            next_action = self.actor_target(next_obs) + noise
            tq1, tq2 = self.critic_target(next_obs, next_action)
            target_q = reward + not_done * self.gamma * torch.min(tq1, tq2)
        q1, q2 = self.critic(obs, action)
        critic_loss = F.mse_loss(q1, target_q) + F.mse_loss(q2, target_q)
        self.critic_optimizer.zero_grad(set_to_none=False)
        critic_loss.backward()
        self.critic_optimizer.step()
        actor_loss = None
        if with_actor:
            actor_loss = -self.critic.Q1(obs, self.actor(obs)).mean()
            self.actor_optimizer.zero_grad(set_to_none=False)
            actor_loss.backward()
            self.actor_optimizer.step()
            with torch.no_grad():
                for p, tp in zip(self.critic.parameters(), self.critic_target.parameters()):
                    tp.mul_(1 - self.tau).add_(p, alpha=self.tau)
                for p, tp in zip(self.actor.parameters(), self.actor_target.parameters()):
                    tp.mul_(1 - self.tau).add_(p, alpha=self.tau)
        return critic_loss.detach(), (actor_loss.detach() if actor_loss is not None else None)


def default_args(**kw):
    """argparse defaults of training/train_td3.py:10-39."""
    a = dict(env_name="base", seed=0, start_timesteps=int(25e3), eval_freq=int(5e3), num_env_steps=int(1e6),
             expl_noise=0.1, batch_size=256, gamma=0.99, tau=0.005, policy_noise=0.2, noise_clip=0.5, policy_freq=2,
             load_model="", max_replay_size=1000000, num_agents=32, no_cuda=False, logdir=None, timestamp=None,
             log_interval=1000, save_interval=2000, task=None, max_seconds=None)
    a.update(kw)
    return SimpleNamespace(**a)


def train(args, config, env_constructor=SoloBaseEnv, writer=None):
    """The loop of agents/td3/train.py:35-154: one vec-env step, one batched replay append and one TD3
    update per iteration; random (unclipped normal) actions for the first ``start_timesteps``."""
    torch.manual_seed(args.seed)
    torch.cuda.manual_seed_all(args.seed)
    device = torch.device("cuda", torch.cuda.current_device())
    N = int(args.num_agents)
    envs = make_vec_envs(config, N, env_constructor, args.gamma, device, seed=args.seed)
    sim = envs.envs.venv.sim
    obs_dim, action_dim = envs.observation_space.shape[0], envs.action_space.shape[0]
    policy = TD3(obs_dim, action_dim, args.gamma, args.tau, args.policy_noise, args.noise_clip, args.policy_freq, device)
    replay = ReplayBuffer(args.max_replay_size, obs_dim, action_dim, device)
    tracker = EpisodeTracker(device)
    obs = envs.reset().clone()
    step, start = 0, time.time()
    q_loss = ac_loss = None
    history = []
    t = 0
    for t in range(0, int(args.num_env_steps), N):
        if t < args.start_timesteps:
            action = torch.randn(N, action_dim, device=device)                      # train.py:98
        else:
            with torch.no_grad():
                action = policy.select_action(obs) + float(args.expl_noise) * torch.randn(N, action_dim, device=device)
        next_obs, rewards, dones, _infos = envs.step_inplace(action)      # copied into the replay ring right below
        tracker.update(sim, dones)
        replay.append_batch(obs, action, rewards, next_obs, 1.0 - dones)
        obs.copy_(next_obs)
        if t >= args.start_timesteps:
            q, a = policy.train(replay, step, args.batch_size)
            q_loss = q
            ac_loss = a if a is not None else ac_loss
            if step % args.log_interval == 0:
                st = tracker.fetch()
                st.update(frames=t, fps=t / max(time.time() - start, 1e-9), q_loss=float(q_loss),
                          actor_loss=float(ac_loss) if ac_loss is not None else float("nan"))
                history.append(st)
                print("Num env frames {}, FPS {}: {} episodes, mean/max return {:.2f}/{:.2f}, critics loss {:.2f}, rl loss {:.2f}".format(
                    t, int(st["fps"]), st["episodes"], st["episode_return"], st["return_max"], st["q_loss"], st["actor_loss"]),
                    flush=True)
                if writer is not None:
                    utils.log(writer, st["actor_loss"], "Loss/action", t)
                    utils.log(writer, st["q_loss"], "Loss/Qval", t)
                    utils.log(writer, st["episode_reward"], "Episode/reward", t)
                    utils.log(writer, st["episode_length"], "Episode/length", t)
                    utils.log(writer, st["success"], "Episode/success_mean", t)
                    for k in _DR:
                        utils.log(writer, st[k], k, t)
            if step % args.save_interval == 0 and args.logdir is not None:
                save_checkpoint(args.logdir, t, policy, "ckpt_{}.pth".format(t))
        step += 1
        if args.max_seconds is not None and time.time() - start > args.max_seconds:
            break
    if args.logdir is not None:
        save_checkpoint(args.logdir, args.num_env_steps, policy, "ckpt_final.pth")
    envs.close()
    return {"history": history, "policy": policy, "replay": replay, "frames": t + N, "seconds": time.time() - start}


def save_checkpoint(logdir, update, policy, name):
    os.makedirs(logdir, exist_ok=True)
    torch.save({"update": update, "state_dict": policy.actor.state_dict(),
                "critic_state_dict": policy.critic.state_dict()}, os.path.join(logdir, name))
