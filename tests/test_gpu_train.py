"""GPU tests of the trainer-side callers: PPO loop on the device vec-env (CUDA-graph rollout and eager),
checkpoint format, evaluation harness."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import make_config

pytestmark = pytest.mark.gpu


def _args(**kw):
    from solorl_b200.agents.train import default_args
    base = dict(num_agents=256, num_steps=16, mini_batch_size=1024, ppo_epoch=2, lr=3e-4, use_gae=True,
                num_env_steps=256 * 16 * 3, log_interval=1, save_interval=1, seed=3)
    base.update(kw)
    return default_args(**base)


@pytest.mark.parametrize("graph", [True, False])
def test_ppo_train_runs_and_checkpoints(tmp_path, graph):
    from solorl_b200.agents import train as ppo
    from solorl_b200.agents.evaluate import evaluate, load_policy, summarize
    from solorl_b200.envs import Box
    cfg = make_config("solo12", "stand", "torque", 1, episode_length=40)
    out = ppo.train(_args(logdir=str(tmp_path), cuda_graph=graph), cfg)
    assert out["updates"] == 3
    for st in out["history"]:
        assert np.isfinite([st["value_loss"], st["action_loss"], st["entropy"]]).all()
    assert sum(st["episodes"] for st in out["history"]) > 0
    ck = torch.load(os.path.join(tmp_path, "solo.pt"), weights_only=False)
    assert set(ck) == {"update", "state_dict", "ob_rms"} and ck["ob_rms"] is None      # agents/ppo/train.py:124-131
    assert any(f.startswith("solo_") for f in os.listdir(tmp_path))
    pol, _ = load_policy(os.path.join(tmp_path, "solo.pt"), (76,), Box(-np.ones(12), np.ones(12)))
    res = evaluate(pol, cfg, num_runs=50, num_envs=32)
    s = summarize(res)
    assert s["episodes"] == 50 and 1 <= s["mean_length"] <= 40 and np.isfinite(s["mean_return"])


def test_graph_rollout_fills_the_buffer_like_eager():
    """The captured rollout must leave every slot of the rollout buffer written and the episode
    accumulators advancing on each replay."""
    from solorl_b200.agents.policy import Policy
    from solorl_b200.agents.storage import OPBuffer
    from solorl_b200.agents.train import EpisodeTracker, Rollout
    from solorl_b200.envs import make_vec_envs
    cfg = make_config("solo8", "walk", "torque", 1, episode_length=10)
    N, T = 128, 12
    envs = make_vec_envs(cfg, N, seed=2)
    ac = Policy(envs.observation_space.shape, envs.action_space, None, {"hidden_size": 64}).cuda()
    buf = OPBuffer(T, N, envs.observation_space.shape, 8, "cuda")
    buf.obs[0].copy_(envs.reset())
    tr = EpisodeTracker(torch.device("cuda"))
    ro = Rollout(envs, ac, buf, tr, T, use_graph=True)
    counts = []
    for it in range(3):
        buf.obs[1:].fill_(float("nan")); buf.rewards.fill_(float("nan"))
        ro()
        torch.cuda.synchronize()
        assert ro.graph is not None
        assert torch.isfinite(buf.obs).all() and torch.isfinite(buf.rewards).all()
        assert torch.isfinite(buf.action_log_probs).all() and torch.isfinite(buf.value_preds[:-1]).all()
        counts.append(tr.fetch(clear=False)["episodes"])
        buf.reset()
    # episode_length 10 and 12 steps per rollout: every env finishes at least once per rollout
    assert counts[0] >= N and counts[1] >= counts[0] + N and counts[2] >= counts[1] + N
    envs.close()


@pytest.mark.parametrize("task,fixture,floor", [("stand", "ppo_stand_solo12.pt", 150.0), ("walk", "ppo_walk_solo12.pt", 150.0)])
def test_episode_return_statistics_of_trained_policy_within_5pct(task, fixture, floor):
    """north_star: episode-return statistics of a FIXED trained PPO policy over 1000 Stand/Walk episodes must
    fall within 5 % of the reference path's.  PyBullet is not installable here (parity unpinned), so the
    comparison is against the fp64 CPU oracle driven by the same checkpoint (tests/golden/ppo_*_solo12.pt,
    trained by training/train_ppo.py on this repo's GPU path: Stand after 2e8 env steps holds the pose for
    the whole episode, return 186; Walk after 1.3e8 env steps lunges forward for about a second before it
    falls, return 260 +- 61)."""
    from solorl_b200.agents.evaluate import evaluate, load_policy, summarize
    from solorl_b200.envs import Box
    from tests.helpers import GOLDEN, oracle_policy_episodes
    cfg = make_config("solo12", task, "torque", 1)
    space = Box(-np.ones(12), np.ones(12))
    pol, _ = load_policy(os.path.join(GOLDEN, fixture), (76,), space)
    gpu = evaluate(pol, cfg, num_runs=1000, num_envs=1000, seed=21)
    s = summarize(gpu)
    assert s["episodes"] == 1000
    cpu_pol, _ = load_policy(os.path.join(GOLDEN, fixture), (76,), space, device="cpu")
    ret, length, last = oracle_policy_episodes(cpu_pol, cfg, 1000, seed=22)
    assert s["mean_return"] > floor                                        # the policy does its task
    assert abs(s["mean_return"] - ret.mean()) <= 0.05 * abs(ret.mean())
    assert abs(s["mean_length"] - length.mean()) <= 0.05 * length.mean()
    assert abs(s["std_return"] - ret.std()) <= 0.05 * abs(ret.mean())
    for q in (10, 50, 90):
        assert abs(np.percentile(gpu["episode_return"], q) - np.percentile(ret, q)) <= 0.05 * abs(ret.mean())
    if task == "stand":
        # last-step reward (what the reference prints as 'episode_reward').  A fall ends an episode with
        # -10, so a handful of falls in 1000 episodes moves the plain mean by several hundredths (7 falls = 0.07):
        # the reward of the episodes that ran to the timeout is compared tightly, the fall rates separately
        g_len, g_last = np.asarray(gpu["episode_length"]), np.asarray(gpu["episode_reward"])
        full_g, full_c = g_len >= cfg["episode_length"], length >= cfg["episode_length"]
        assert abs(g_last[full_g].mean() - last[full_c].mean()) <= 0.02
        assert abs((~full_g).mean() - (~full_c).mean()) <= 0.015


def test_td3_train_runs_on_the_vec_env(tmp_path):
    from solorl_b200.agents import td3
    cfg = make_config("solo8", "stand", "torque", 1, episode_length=30)
    args = td3.default_args(num_agents=64, start_timesteps=64 * 20, num_env_steps=64 * 60, batch_size=128,
                            log_interval=10, save_interval=20, logdir=str(tmp_path), max_replay_size=4096)
    out = td3.train(args, cfg)
    assert len(out["replay"]) == 64 * 60 and out["frames"] == 64 * 60
    assert out["history"] and all(np.isfinite(h["q_loss"]) for h in out["history"])
    ck = torch.load(os.path.join(tmp_path, "ckpt_final.pth"), weights_only=False)
    assert set(ck) == {"update", "state_dict", "critic_state_dict"}          # agents/td3/train.py:142-154
    # transitions stored by the ring are the env's: obs rows are finite, not_terminal is 0/1
    assert torch.isfinite(out["replay"]._observations[:64 * 60]).all()
    nt = out["replay"]._not_terminal[:64 * 60]
    assert ((nt == 0) | (nt == 1)).all() and (nt == 0).any()


def test_fused_episode_accumulator_matches_host_records():
    """solo_accumulate_episode_stats against the same sums taken from the per-env records on the host."""
    from solorl_b200.agents.train import EpisodeTracker
    from solorl_b200.envs import SoloVecEnv
    cfg = make_config("solo12", "pointgoal", "torque", 1, episode_length=8)
    n = 512
    env = SoloVecEnv(cfg, n, device="cuda:0", seed=5)
    env.reset()
    tr = EpisodeTracker(torch.device("cuda"))
    g = torch.Generator(device="cuda").manual_seed(1)
    tot = dict(n=0, rew=0.0, ret=0.0, ln=0, succ=0, dr=np.zeros(5), mn=np.inf, mx=-np.inf, lmx=0)
    for t in range(20):
        a = torch.rand(n, 12, device="cuda", generator=g) * 3 - 1.5
        obs, rew, done, infos = env.step(a)
        tr.update(env.sim, done)
        r = infos.done_records()
        tot["n"] += len(r); tot["rew"] += float(r["episode_reward"].astype(np.float64).sum())
        tot["ret"] += float(r["episode_return"].astype(np.float64).sum()); tot["ln"] += int(r["episode_length"].sum())
        tot["succ"] += int(r["success"].sum())
        for k, f in enumerate(("dr_stand", "dr_joint_pose", "dr_torque", "dr_balance", "dr_progress")):
            tot["dr"][k] += float(r[f].astype(np.float64).sum())
        if len(r):
            tot["mn"] = min(tot["mn"], float(r["episode_return"].min())); tot["mx"] = max(tot["mx"], float(r["episode_return"].max()))
            tot["lmx"] = max(tot["lmx"], int(r["episode_length"].max()))
    st = tr.fetch()
    assert st["episodes"] == tot["n"] > n
    assert abs(st["episode_return"] - tot["ret"] / tot["n"]) < 1e-6 * max(1, abs(tot["ret"] / tot["n"]))
    assert abs(st["episode_reward"] - tot["rew"] / tot["n"]) < 1e-6 and abs(st["episode_length"] - tot["ln"] / tot["n"]) < 1e-9
    assert abs(st["success"] - tot["succ"] / tot["n"]) < 1e-9
    assert abs(st["dr/stand_rew"] - tot["dr"][0] / tot["n"]) < 1e-6 and abs(st["dr/progress_rew"] - tot["dr"][4] / tot["n"]) < 1e-5
    assert st["return_min"] == tot["mn"] and st["return_max"] == tot["mx"] and st["length_max"] == tot["lmx"]
    assert tr.fetch()["episodes"] == 0
    env.close()


def test_graphed_ppo_update_matches_eager():
    """The mini-batch step replayed as a CUDA graph takes the same optimisation steps as the eager loop
    (same random mini-batches, same Adam state): weights agree after three updates, the second and third of
    which run almost entirely from the graph."""
    import copy
    from solorl_b200.agents.policy import Policy
    from solorl_b200.agents.ppo import PPO
    from solorl_b200.agents.storage import OPBuffer
    from solorl_b200.agents.utils import update_linear_schedule
    from solorl_b200.envs import Box
    torch.manual_seed(0)
    T, N, D, A = 8, 64, 20, 6
    ac1 = Policy((D,), Box(-np.ones(A), np.ones(A)), None, {"hidden_size": 32}).cuda()
    ac2 = copy.deepcopy(ac1)
    agents = [PPO(ac1, 0.1, 3, 128, 0.5, 0.01, lr=1e-3, max_grad_norm=0.5, use_graph=True),
              PPO(ac2, 0.1, 3, 128, 0.5, 0.01, lr=1e-3, max_grad_norm=0.5, use_graph=False)]
    bufs = [OPBuffer(T, N, (D,), A, "cuda") for _ in range(2)]
    for it in range(3):
        g = torch.Generator(device="cuda").manual_seed(100 + it)
        data = {k: torch.randn(*getattr(bufs[0], k).shape, device="cuda", generator=g)
                for k in ("obs", "actions", "value_preds", "returns")}
        with torch.no_grad():
            _, lp, _ = ac2.evaluate_actions(data["obs"][:-1].reshape(-1, D), data["actions"].reshape(-1, A))
        for agent, buf in zip(agents, bufs):
            for k, v in data.items():
                getattr(buf, k).copy_(v)
            buf.action_log_probs.copy_(lp.reshape(T, N, 1))
            update_linear_schedule(agent.optimizer, it, 5, 1e-3)
            torch.manual_seed(7 + it)                      # same randperm sequence for both
            losses = agent.update(buf)
            assert np.isfinite(losses).all()
        assert agents[0]._graph is not None or it == 0
        for p1, p2 in zip(ac1.parameters(), ac2.parameters()):
            assert torch.allclose(p1, p2, atol=2e-6), it


def test_td3_and_ppo_evaluation_clis(tmp_path):
    """testing/test_ppo.py and testing/test_td3.py on freshly written checkpoints (reference layouts)."""
    import importlib.util
    import yaml
    from tests.helpers import ROOT
    from solorl_b200.agents import td3, train as ppo

    def load(name):
        spec = importlib.util.spec_from_file_location(name[:-3], os.path.join(ROOT, "testing", name))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    cfg = make_config("solo8", "stand", "torque", 1, episode_length=20)
    cfg_file = tmp_path / "cfg.yaml"
    cfg_file.write_text(yaml.safe_dump(cfg))
    pdir, tdir = tmp_path / "ppo", tmp_path / "td3"
    ppo.train(_args(logdir=str(pdir), num_agents=64, num_env_steps=64 * 16 * 2), cfg)
    td3.train(td3.default_args(num_agents=32, start_timesteps=32 * 5, num_env_steps=32 * 30, batch_size=64,
                               save_interval=10, logdir=str(tdir), max_replay_size=2048), cfg)
    s1 = load("test_ppo.py").main(["--checkpoint-dir", str(pdir), "--config-file", str(cfg_file), "--num-runs", "20"])
    s2 = load("test_td3.py").main(["--checkpoint-dir", str(tdir), "--config-file", str(cfg_file), "--num-runs", "20"])
    for s in (s1, s2):
        assert s["episodes"] == 20 and 1 <= s["mean_length"] <= 20 and np.isfinite(s["mean_return"])


def test_graphed_td3_update_matches_eager():
    import copy
    from solorl_b200.agents.td3 import TD3, ReplayBuffer
    torch.manual_seed(0)
    a = TD3(12, 4, device="cuda", use_graph=True)
    b = TD3(12, 4, device="cuda", use_graph=False)
    for src, dst in ((a.actor, b.actor), (a.critic, b.critic), (a.actor_target, b.actor_target),
                     (a.critic_target, b.critic_target)):
        dst.load_state_dict(copy.deepcopy(src.state_dict()))
    rb = ReplayBuffer(4096, 12, 4, "cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    rb.append_batch(torch.randn(3000, 12, device="cuda", generator=g), torch.randn(3000, 4, device="cuda", generator=g).tanh(),
                    torch.randn(3000, device="cuda", generator=g), torch.randn(3000, 12, device="cuda", generator=g),
                    (torch.rand(3000, device="cuda", generator=g) > 0.1).float())
    for step in range(12):                      # 3 eager warm-ups per kind of step, then capture, then replays
        for agent in (a, b):
            torch.manual_seed(50 + step)        # same batch indices and the same target-smoothing noise
            q, al = agent.train(rb, step, 256)
            assert torch.isfinite(q)
    assert a._graphs.get(True) is not None and a._graphs.get(False) is not None
    for m1, m2 in ((a.actor, b.actor), (a.critic, b.critic), (a.actor_target, b.actor_target)):
        for p1, p2 in zip(m1.parameters(), m2.parameters()):
            assert torch.allclose(p1, p2, atol=5e-6)


def test_reference_shaped_cli_invocation_configs0(tmp_path):
    """BASELINE configs[0]: `train_ppo.py --config-file configs/basic.yaml --num-agents 64` exactly as the
    reference README runs it (rollout length = episode_length = 400, README.md:34-36 hyper-parameters), for two
    updates; then the checkpoint is rolled out with testing/test_ppo.py."""
    import importlib.util
    from tests.helpers import ROOT

    def load(rel):
        spec = importlib.util.spec_from_file_location(os.path.basename(rel)[:-3], os.path.join(ROOT, rel))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    out = load("training/train_ppo.py").main([
        "--config-file", os.path.join(ROOT, "configs", "basic.yaml"), "--num-agents", "64", "--lr", "2.5e-4",
        "--clip-param", "0.1", "--ppo-epoch", "5", "--mini-batch-size", "512", "--use-gae", "--use-linear-lr-decay",
        "--num-env-steps", str(64 * 400 * 2), "--log-interval", "1", "--logdir", str(tmp_path), "--timestamp", "t"])
    assert out["updates"] == 2 and out["last"]["episodes"] > 0
    run = [d for d in os.listdir(tmp_path) if d.startswith("SoloBase_")]
    assert len(run) == 1 and os.path.exists(os.path.join(tmp_path, run[0], "solo.pt"))
    s = load("testing/test_ppo.py").main(["--checkpoint-dir", os.path.join(tmp_path, run[0]), "--config-file",
                                          os.path.join(ROOT, "configs", "basic.yaml"), "--num-runs", "10"])
    assert s["episodes"] == 10
