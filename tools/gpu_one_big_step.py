"""A few env steps at a machine-filling batch (throughput build), for ncu captures. Usage: python tools/gpu_one_big_step.py [n]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solorl_b200.envs import SoloVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cfg = {"model_urdf": "solo12", "mode": "headless", "episode_length": 400, "frame_skip": 4, "control": "torque",
       "task": "walk", "num_history_stack": 1}
env = SoloVecEnv(cfg, n, device="cuda:0", seed=3)
env.reset()
g = torch.Generator(device="cuda").manual_seed(7)
for i in range(40):
    env.sim.step(torch.rand(n, 12, device="cuda", generator=g) * 2 - 1)
torch.cuda.synchronize()
print("done", n)
