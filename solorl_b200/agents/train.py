"""PPO driver with the behaviour of the reference's ``agents/ppo/train.py:21-162`` on the GPU vec-env.

Same loop — rollout of ``num_steps`` vec-env steps, GAE, clipped-surrogate update, linear LR decay,
curriculum hook, checkpoints ``solo_{steps}.pt`` / ``solo.pt`` holding ``{'update','state_dict',
'ob_rms'}`` (train.py:121-131) — with three changes to the data path:

* nothing leaves the device during a rollout: the reference reads ``done[i].item()`` and an info dict
  for every env and step (train.py:90-100); here finished episodes are folded into device-side
  accumulators from ``solo_episode_stats`` and read once per log interval;
* the whole rollout (policy act -> env step -> buffer append, ``num_steps`` times) is captured in ONE
  CUDA graph and replayed, so the ~25 small launches per step cost no CPU time;
* under ``torchrun`` every rank owns an env shard and the PPO update all-reduces gradients
  (``agents/ppo.py``); rank 0 logs and checkpoints.
"""
from __future__ import annotations

import os
import time
from types import SimpleNamespace

import torch

from ..envs import SoloBaseEnv, make_vec_envs
from . import utils
from .policy import Policy
from .ppo import PPO, broadcast_parameters, dist_ready
from .storage import OPBuffer

_DR = ("dr/stand_rew", "dr/joint_pose_rew", "dr/torque_rew", "dr/roll_pitch_balance_rew", "dr/progress_rew")


class EpisodeTracker:
    """Device-side statistics of finished episodes (what train.py:90-100 appends to its deques): one
    fused launch per step (``solo_accumulate_episode_stats``) into a 13-double accumulator that is read
    once per log interval."""

    def __init__(self, device):
        self.device = torch.device(device)
        self._init = torch.tensor([0.0] * 10 + [float("inf"), float("-inf"), 0.0], dtype=torch.float64, device=device)
        self.acc = self._init.clone()

    def clear(self):
        """In place: the tensor is baked into the captured rollout graph."""
        self.acc.copy_(self._init)

    def update_from_tensors(self, done, episode_return, episode_length, success, last_reward):
        """Envs without a C-ABI episode record (the gait-env shell keeps its episode sums as tensors): the same
        accumulator filled with torch ops (no host sync)."""
        m = (done > 0.5).double()
        ret, ln = episode_return.double(), episode_length.double()
        self.acc[0] += m.sum()
        self.acc[1] += (m * last_reward.double()).sum()
        self.acc[2] += (m * ret).sum()
        self.acc[3] += (m * ln).sum()
        self.acc[4] += (m * success.double()).sum()
        big = torch.full_like(ret, float("inf"))
        self.acc[10] = torch.minimum(self.acc[10], torch.where(m > 0, ret, big).min())
        self.acc[11] = torch.maximum(self.acc[11], torch.where(m > 0, ret, -big).max())
        self.acc[12] = torch.maximum(self.acc[12], (m * ln).max())

    def update(self, sim, done):
        if self.acc.is_cuda:
            sim.accumulate_episode_stats(done, self.acc)
        else:                                    # host-side logic tests: same arithmetic with torch ops
            f, i = sim.episode_stats_device()
            m = done > 0.5
            if bool(m.any()):
                cols = [torch.ones_like(f[:, 0]), f[:, 0], f[:, 1], i[:, 2].float(), i[:, 3].float(), f[:, 6],
                        f[:, 7], f[:, 8], f[:, 9], f[:, 10]]
                self.acc[:10] += torch.stack([c[m].double().sum() for c in cols])
                self.acc[10] = torch.minimum(self.acc[10], f[m, 1].double().min())
                self.acc[11] = torch.maximum(self.acc[11], f[m, 1].double().max())
                self.acc[12] = torch.maximum(self.acc[12], i[m, 2].double().max())

    def fetch(self, clear=True):
        """One host sync; totals are summed over ranks when torch.distributed is up."""
        acc = self.acc.clone()
        if dist_ready():
            sums, mn, mx = acc[:10].clone(), acc[10:11].clone(), acc[11:].clone()
            torch.distributed.all_reduce(sums)
            torch.distributed.all_reduce(mn, op=torch.distributed.ReduceOp.MIN)
            torch.distributed.all_reduce(mx, op=torch.distributed.ReduceOp.MAX)
            acc = torch.cat([sums, mn, mx])
        a = acc.tolist()
        n = max(a[0], 1.0)
        out = {"episodes": int(a[0]), "episode_reward": a[1] / n, "episode_return": a[2] / n,
               "episode_length": a[3] / n, "success": a[4] / n,
               "return_min": a[10], "return_max": a[11], "length_max": a[12]}
        for k, v in zip(_DR, a[5:10]):
            out[k] = v / n
        if clear:
            self.clear()
        return out


class Rollout:
    """``num_steps`` x (act, env.step, append), eager or as one replayed CUDA graph."""

    def __init__(self, envs, actor_critic, buf, tracker, num_steps, use_graph=True):
        self.envs, self.ac, self.buf, self.tracker, self.T = envs, actor_critic, buf, tracker, num_steps
        self.venv = envs.envs.venv
        self.sim = getattr(self.venv, "sim", None)       # None: the gait-env shell (its own tick-loop graph inside)
        self.graph = None
        self.use_graph = use_graph and self.sim is not None

    def _run(self):
        for t in range(self.T):
            with torch.no_grad():
                value, action, logp = self.ac.act(self.buf.obs[t])
                obs, reward, done, _infos = self.envs.step_inplace(action)    # copied into buf right below
                if self.sim is not None:
                    self.tracker.update(self.sim, done)
                else:
                    li = self.venv.last_info
                    self.tracker.update_from_tensors(done, li["episode_reward"], li["episode_length"], li["success"],
                                                     reward.reshape(-1))
                self.buf.append(obs, action, logp, value, reward, (1.0 - done).unsqueeze(-1))

    def __call__(self):
        if not self.use_graph:
            return self._run()
        if self.graph is None:
            try:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._run()                        # warm-up on a side stream (allocator, lazy init)
                torch.cuda.current_stream().wait_stream(s)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run()
                self.graph = g
                # capture does not execute: the rollout of THIS call is the eager warm-up above, which
                # was a regular rollout from buf.obs[0]; later calls replay the graph
            except Exception as e:                     # pragma: no cover - capture is best effort
                print(f"[ppo] CUDA graph capture failed ({e}); running the rollout eagerly", flush=True)
                self.use_graph = False
                self.graph = None
                return self._run()
            return
        self.graph.replay()


def default_args(**kw):
    """The argparse defaults of training/train_ppo.py:9-45 as a namespace (for programmatic use)."""
    a = dict(num_agents=32, output_size=64, hidden_size=64, no_cuda=False, env_name="base", gamma=0.99, tau=0.95,
             clip_param=0.1, ppo_epoch=10, mini_batch_size=32, lr=1e-3, l2_coef=0.0, value_loss_coef=0.5,
             entropy_coef=0.01, max_grad_norm=0.5, clip_value_loss=False, use_linear_lr_decay=False, use_gae=False,
             num_env_steps=int(1e6), seed=2301, curriculum_schedule=0, log_interval=10, logdir=None,
             base_checkpoint=None, timestamp=None, save_interval=20, task=None, num_steps=None, cuda_graph=True,
             max_seconds=None, target_return=None)
    a.update(kw)
    return SimpleNamespace(**a)


def train(args, config, env_constructor=SoloBaseEnv, writer=None):
    """Returns a dict with the final statistics (the reference calls sys.exit() instead, train.py:162)."""
    rank = torch.distributed.get_rank() if dist_ready() else 0
    world = torch.distributed.get_world_size() if dist_ready() else 1
    torch.manual_seed(args.seed + rank)
    torch.cuda.manual_seed_all(args.seed + rank)
    device = torch.device("cuda", torch.cuda.current_device())
    num_steps = int(args.num_steps or config["episode_length"])        # train_ppo.py:62-63
    N = int(args.num_agents)

    envs = make_vec_envs(config, N, env_constructor, args.gamma, device, seed=args.seed,
                         env_id_offset=rank * N)
    discrete = envs.action_space.__class__.__name__ == "Discrete"
    action_dim = 1 if discrete else envs.action_space.shape[0]        # storage_dim of agents/ppo/train.py:34-39
    base = torch.load(args.base_checkpoint, map_location=device) if args.base_checkpoint is not None else None
    actor_critic = Policy(envs.observation_space.shape, envs.action_space, base, {"hidden_size": args.hidden_size})
    actor_critic.to(device)
    broadcast_parameters(actor_critic)
    agent = PPO(actor_critic, args.clip_param, args.ppo_epoch, args.mini_batch_size, args.value_loss_coef,
                args.entropy_coef, lr=args.lr, l2_coef=args.l2_coef, max_grad_norm=args.max_grad_norm)
    buf = OPBuffer(num_steps, N, envs.observation_space.shape, action_dim, device)
    buf.obs[0].copy_(envs.reset_inplace())
    tracker = EpisodeTracker(device)
    rollout = Rollout(envs, actor_critic, buf, tracker, num_steps, use_graph=getattr(args, "cuda_graph", True))

    num_updates = max(int(args.num_env_steps) // num_steps // N // world, 1)
    start = time.time()
    last = {"episodes": 0}
    history = []
    j = 0
    for j in range(num_updates):
        if args.use_linear_lr_decay:
            utils.update_linear_schedule(agent.optimizer, j, num_updates, args.lr)
        rollout()
        with torch.no_grad():
            next_value = actor_critic.get_value(buf.obs[-1]).detach()
        buf.compute_returns(next_value, args.use_gae, args.gamma, args.tau)
        step_reward = buf.rewards.mean()           # mean reward per env step of this rollout (read only at log time)
        value_loss, action_loss, dist_entropy = agent.update(buf)
        buf.reset()
        if args.curriculum_schedule and (j + 1) % args.curriculum_schedule == 0:
            envs.increment_curriculum()
        total_num_steps = (j + 1) * N * num_steps * world
        last_update = j == num_updates - 1
        if (j % args.save_interval == 0 or last_update) and args.logdir is not None and rank == 0:
            save_checkpoint(args.logdir, j, actor_critic, envs, total_num_steps)
        stop = False
        if j % args.log_interval == 0 or last_update:
            st = tracker.fetch()
            elapsed = time.time() - start
            st.update(update=j, steps=total_num_steps, seconds=elapsed, fps=total_num_steps / max(elapsed, 1e-9),
                      value_loss=value_loss, action_loss=action_loss, entropy=dist_entropy,
                      mean_step_reward=float(step_reward))
            history.append(st)
            last = st
            if rank == 0:
                print("Updates {}, num timesteps {}, FPS {}\n {} training episodes: mean return {:.2f} (min {:.2f} max {:.2f}), "
                      "mean last-step reward {:.3f}, mean/max length {:.0f}/{:.0f}, mean success {:.2f}\n"
                      " entropy {:.2f} value loss {:.3f} action loss {:.4f}".format(
                          j, total_num_steps, int(st["fps"]), st["episodes"], st["episode_return"], st["return_min"],
                          st["return_max"], st["episode_reward"], st["episode_length"], st["length_max"], st["success"],
                          dist_entropy, value_loss, action_loss), flush=True)
                if writer is not None:
                    utils.log(writer, value_loss, "Loss/value", total_num_steps)
                    utils.log(writer, action_loss, "Loss/action", total_num_steps)
                    utils.log(writer, dist_entropy, "Loss/entropy", total_num_steps)
                    utils.log(writer, st["episode_reward"], "Episode/reward", total_num_steps)
                    utils.log(writer, st["episode_return"], "Episode/return", total_num_steps)
                    utils.log(writer, st["success"], "Episode/success_mean", total_num_steps)
                    utils.log(writer, st["episode_length"], "Episode/length", total_num_steps)
                    for k in _DR:
                        utils.log(writer, st[k], k, total_num_steps)
            if args.target_return is not None and st["episodes"] > 0 and st["episode_return"] >= args.target_return:
                stop = True
            if args.max_seconds is not None and elapsed >= args.max_seconds:
                stop = True
            if dist_ready():
                # every rank must leave the loop at the same update: `elapsed` is a per-rank clock, and a rank
                # that stopped alone would leave the others waiting in the next gradient all-reduce
                flag = torch.tensor([1.0 if stop else 0.0], device=device)
                torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MAX)
                stop = bool(flag.item() > 0.5)
        if stop:
            break
    if args.logdir is not None and rank == 0:
        save_checkpoint(args.logdir, j, actor_critic, envs, (j + 1) * N * num_steps * world)
    envs.close()
    return {"last": last, "history": history, "actor_critic": actor_critic, "updates": j + 1,
            "seconds": time.time() - start}


def save_checkpoint(logdir, update, actor_critic, envs, total_num_steps):
    os.makedirs(logdir, exist_ok=True)
    blob = {"update": update, "state_dict": actor_critic.state_dict(), "ob_rms": getattr(envs.envs, "ob_rms", None)}
    torch.save(blob, os.path.join(logdir, "solo_{}.pt".format(total_num_steps)))
    torch.save(blob, os.path.join(logdir, "solo.pt"))
