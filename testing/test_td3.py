#!/usr/bin/env python
"""Roll out a TD3 actor checkpoint — the reference's ``testing/test_td3.py`` (an endless GUI roll-out of the newest
``ckpt_*`` file) as a bounded, batched evaluation: ``--num-runs`` episodes, then mean length / reward / success.

  python testing/test_td3.py --checkpoint-dir runs/SoloBase_... --config-file configs/basic.yaml --task stand
"""
import argparse
import glob
import json
import os
import re
import sys

import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solorl_b200.agents.evaluate import evaluate, summarize  # noqa: E402
from solorl_b200.agents.td3 import Actor  # noqa: E402
from solorl_b200.envs import SoloBaseEnv, make_vec_envs  # noqa: E402


def newest_checkpoint(directory):
    """ckpt_final.pth if present, else the ckpt_<frames>.pth with the largest number (test_td3.py:24-31)."""
    ckpts = [os.path.basename(p) for p in glob.glob(os.path.join(directory, "ckpt_*"))]
    if not ckpts:
        raise FileNotFoundError(f"no ckpt_* file in {directory}")
    if "ckpt_final.pth" in ckpts:
        return os.path.join(directory, "ckpt_final.pth")
    ckpts.sort(key=lambda n: int(re.sub("[^0-9]", "", n) or 0))
    return os.path.join(directory, ckpts[-1])


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--checkpoint-dir", type=str, required=True)
    p.add_argument("--config-file", type=str, default=os.path.join(os.path.dirname(__file__), "..", "configs", "basic.yaml"))
    p.add_argument("--mode", type=str, default="headless")
    p.add_argument("--task", type=str, default=None)
    p.add_argument("--num-runs", type=int, default=10)
    p.add_argument("--num-envs", type=int, default=None)
    p.add_argument("--json", action="store_true", default=False)
    args = p.parse_args(argv)
    with open(args.config_file, "r") as f:
        config = yaml.load(f, Loader=yaml.FullLoader)
    config["mode"] = "headless"
    if args.task is not None:
        config["task"] = args.task
    filename = newest_checkpoint(args.checkpoint_dir)
    ckpt = torch.load(filename, map_location="cuda", weights_only=False)
    probe = make_vec_envs(config, 1, SoloBaseEnv, training=False)
    obs_dim, act_dim = probe.observation_space.shape[0], probe.action_space.shape[0]
    probe.close()
    policy = Actor(obs_dim, act_dim).cuda()
    policy.load_state_dict(ckpt["state_dict"])
    policy.eval()
    s = summarize(evaluate(policy, config, num_runs=args.num_runs, num_envs=args.num_envs))
    if args.json:
        print(json.dumps(dict(s, checkpoint=os.path.basename(filename))))
    else:
        print(os.path.basename(filename))
        print("mean length {} mean reward {} mean success {}".format(s["mean_length"], s["mean_reward"], s["mean_success"]))
        print("episodes {} mean return {:.3f} std {:.3f}".format(s["episodes"], s["mean_return"], s["std_return"]))
    return s


if __name__ == "__main__":
    main()
