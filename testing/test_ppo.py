#!/usr/bin/env python
"""Roll out a PPO checkpoint — the CLI of the reference's ``testing/test_ppo.py:22-151`` for ``--env-name base``.

  python testing/test_ppo.py --checkpoint-dir runs/SoloBase_stand_... --config-file configs/basic12.yaml \\
      --task stand --num-runs 1000

Prints ``mean length / mean reward / mean success`` like the reference (mean reward = the LAST step's reward,
SURVEY F8) plus the mean and spread of the episode RETURN (sum of rewards)."""
import argparse
import json
import os
import sys

import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solorl_b200.agents.evaluate import evaluate, load_policy, summarize  # noqa: E402
from solorl_b200.envs import SoloBaseEnv, make_vec_envs  # noqa: E402


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--checkpoint-dir", type=str, required=True)
    p.add_argument("--config-file", type=str, required=True)
    p.add_argument("--mode", type=str, default="headless")
    p.add_argument("--env-name", type=str, default="base")
    p.add_argument("--task", type=str, default=None)
    p.add_argument("--num-runs", type=int, default=10)
    p.add_argument("--episode-length", type=int, default=None)
    p.add_argument("--num-envs", type=int, default=None)
    p.add_argument("--deterministic", action="store_true", default=False)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--json", action="store_true", default=False)
    args = p.parse_args(argv)
    if args.env_name != "base":
        raise NotImplementedError("Error Env {} not found!".format(args.env_name))
    with open(args.config_file, "r") as f:
        config = yaml.load(f, Loader=yaml.FullLoader)
    config["mode"] = "headless" if args.mode == "gui" else args.mode      # no GUI on a GPU box
    if args.episode_length is not None:
        config["episode_length"] = args.episode_length
    if args.task is not None:
        config["task"] = args.task
    probe = make_vec_envs(config, 1, SoloBaseEnv, training=False)
    obs_shape, action_space = probe.observation_space.shape, probe.action_space
    probe.close()
    policy, ckpt = load_policy(os.path.join(args.checkpoint_dir, "solo.pt"), obs_shape, action_space)
    res = evaluate(policy, config, num_runs=args.num_runs, num_envs=args.num_envs, deterministic=args.deterministic,
                   seed=args.seed)
    s = summarize(res)
    if args.json:
        print(json.dumps(s))
    else:
        print("mean length {} mean reward {} mean success {}".format(s["mean_length"], s["mean_reward"], s["mean_success"]))
        print("episodes {} mean return {:.3f} std {:.3f}".format(s["episodes"], s["mean_return"], s["std_return"]))
    return s


if __name__ == "__main__":
    main()
