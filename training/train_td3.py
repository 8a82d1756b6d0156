#!/usr/bin/env python
"""TD3 on the GPU vec-env — the CLI of the reference's ``training/train_td3.py:10-75``.

  python training/train_td3.py --config-file configs/basic.yaml --task stand --num-agents 1024 --logdir runs
"""
import argparse
import os
import sys
from datetime import datetime

import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solorl_b200.agents import td3  # noqa: E402
from solorl_b200.envs import SoloBaseEnv  # noqa: E402


def get_td3_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--env-name", type=str, default="base")
    p.add_argument("--seed", default=0, type=int)
    p.add_argument("--start-timesteps", default=25e3, type=float)
    p.add_argument("--eval-freq", default=5e3, type=float)
    p.add_argument("--num-env-steps", default=1e6, type=float)
    p.add_argument("--expl-noise", default=0.1, type=float)
    p.add_argument("--batch-size", default=256, type=int)
    p.add_argument("--gamma", default=0.99, type=float)
    p.add_argument("--tau", default=0.005, type=float)
    p.add_argument("--policy-noise", default=0.2, type=float)
    p.add_argument("--noise-clip", default=0.5, type=float)
    p.add_argument("--policy-freq", default=2, type=int)
    p.add_argument("--load-model", default="")
    p.add_argument("--max-replay-size", type=int, default=1000000)
    p.add_argument("--num-agents", type=int, default=32)
    p.add_argument("--no-cuda", action="store_true", default=False)
    p.add_argument("--logdir", type=str, default=None)
    p.add_argument("--timestamp", type=str, default=None)
    p.add_argument("--log-interval", type=int, default=1000)
    p.add_argument("--save-interval", type=int, default=2000)
    p.add_argument("--config-file", type=str,
                   default=os.path.join(os.path.dirname(__file__), "..", "configs", "basic.yaml"))
    p.add_argument("--task", type=str, default=None)
    p.add_argument("--max-seconds", type=float, default=None)
    return p.parse_args(argv)


def main(argv=None):
    args = get_td3_args(argv)
    if args.no_cuda or not torch.cuda.is_available():
        raise SystemExit("the env step runs on a CUDA device only (there is no CPU fallback)")
    with open(args.config_file, "r") as f:
        config = yaml.load(f, Loader=yaml.FullLoader)
    if args.task is not None:
        config["task"] = args.task
    args.episode_length = config["episode_length"]
    stamp = datetime.now().strftime("%Y%m%d-%H%M%S") if args.timestamp is None else \
        datetime.now().strftime("%Y%m%d-") + args.timestamp
    writer = None
    if args.logdir is not None:
        args.logdir = os.path.join(args.logdir, "Solo" + args.env_name.capitalize() + "_" + stamp)
        try:
            from torch.utils.tensorboard import SummaryWriter
            writer = SummaryWriter(args.logdir)
        except Exception:
            os.makedirs(args.logdir, exist_ok=True)
    if args.env_name != "base":
        raise NotImplementedError("Error Env {} not found!".format(args.env_name))
    return td3.train(args, config, SoloBaseEnv, writer)


if __name__ == "__main__":
    main()
