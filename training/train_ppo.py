#!/usr/bin/env python
"""PPO on the GPU vec-env — the CLI of the reference's ``training/train_ppo.py:9-101``.

Single GPU:   python training/train_ppo.py --config-file configs/basic12.yaml --task stand --num-agents 4096 \\
                  --num-steps 32 --mini-batch-size 16384 --ppo-epoch 5 --lr 3e-4 --use-gae --logdir runs
Multi GPU:    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
                  training/train_ppo.py ... (--num-agents is PER GPU; env shards are independent, gradients are
                  all-reduced over NCCL)

Flags and defaults are the reference's.  Additions: ``--num-steps`` (the reference forces it to
``episode_length``, train_ppo.py:62-63; at 4096 envs that is 1.6 M samples per update), ``--task`` is honoured
(commented out in the reference, :57-60), ``--max-seconds`` / ``--target-return`` stop early, ``--no-cuda-graph``.
"""
import argparse
import os
import sys
from datetime import datetime

import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solorl_b200.agents import train as ppo  # noqa: E402
from solorl_b200.envs import SoloBaseEnv  # noqa: E402


def get_ppo_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--num-agents", type=int, default=32)
    p.add_argument("--output-size", type=int, default=64)
    p.add_argument("--hidden-size", type=int, default=64)
    p.add_argument("--no-cuda", action="store_true", default=False)
    p.add_argument("--env-name", default="base")
    p.add_argument("--gamma", type=float, default=0.99)
    p.add_argument("--tau", type=float, default=0.95)
    p.add_argument("--clip-param", type=float, default=0.1)
    p.add_argument("--ppo-epoch", type=int, default=10)
    p.add_argument("--mini-batch-size", type=int, default=32)
    p.add_argument("--lr", type=float, default=1e-3)
    p.add_argument("--l2-coef", type=float, default=0.0)
    p.add_argument("--value-loss-coef", type=float, default=0.5)
    p.add_argument("--entropy-coef", type=float, default=0.01)
    p.add_argument("--max-grad-norm", type=float, default=0.5)
    p.add_argument("--clip-value-loss", action="store_true", default=False)
    p.add_argument("--use-linear-lr-decay", action="store_true", default=False)
    p.add_argument("--use-gae", action="store_true", default=False)
    p.add_argument("--num-env-steps", type=float, default=1e6)
    p.add_argument("--seed", type=int, default=2301)
    p.add_argument("--curriculum-schedule", type=int, default=0)
    p.add_argument("--log-interval", type=int, default=10)
    p.add_argument("--logdir", default=None)
    p.add_argument("--base-checkpoint", default=None)
    p.add_argument("--timestamp", default=None)
    p.add_argument("--save-interval", type=int, default=20)
    p.add_argument("--config-file", default=os.path.join(os.path.dirname(__file__), "..", "configs", "basic.yaml"))
    p.add_argument("--task", default=None)
    # additions
    p.add_argument("--num-steps", type=int, default=None, help="rollout length (default: episode_length)")
    p.add_argument("--max-seconds", type=float, default=None)
    p.add_argument("--target-return", type=float, default=None)
    p.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false", default=True)
    return p.parse_args(argv)


def parse_config(config_file):
    with open(config_file, "r") as f:
        return yaml.load(f, Loader=yaml.FullLoader)


def main(argv=None):
    args = get_ppo_args(argv)
    if args.no_cuda or not torch.cuda.is_available():
        raise SystemExit("the env step runs on a CUDA device only (there is no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    rank = int(os.environ.get("RANK", "0"))
    config = parse_config(args.config_file)
    if args.task is not None:
        config["task"] = args.task
    if rank == 0:
        print(config)
    args.episode_length = config["episode_length"]
    stamp = datetime.now().strftime("%Y%m%d-%H%M%S") if args.timestamp is None else \
        datetime.now().strftime("%Y%m%d-") + args.timestamp
    writer = None
    if args.logdir is not None:
        task = args.task + "_" if args.task is not None else ""
        args.logdir = os.path.join(args.logdir, "Solo" + args.env_name.capitalize() + "_" + task + stamp)
        if rank == 0:
            try:
                from torch.utils.tensorboard import SummaryWriter
                writer = SummaryWriter(args.logdir)
            except Exception:       # tensorboard is optional in this image
                os.makedirs(args.logdir, exist_ok=True)
    if args.env_name == "base":                                   # training/train_ppo.py:76-99
        env_constructor = SoloBaseEnv
    elif args.env_name == "contact":
        from solorl_b200.gait import SoloGaitEnvContact as env_constructor
    else:
        raise NotImplementedError(f"Error Env {args.env_name} not found! ('base' = SoloBaseEnv and 'contact' = "
                                  "SoloGaitEnvContact are built; the other gait / timing envs cannot be constructed "
                                  "from any shipped config)")
    out = ppo.train(args, config, env_constructor, writer)
    if world > 1:
        torch.distributed.destroy_process_group()
    return out


if __name__ == "__main__":
    main()
