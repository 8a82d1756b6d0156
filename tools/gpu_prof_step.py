"""Profiling target: N env steps of the headline workload with one build of the step kernel (run under ncu).
Usage: python tools/gpu_prof_step.py [variant=auto] [n=4096] [steps=24] [body_contacts=0]"""
import os
import sys

variant = sys.argv[1] if len(sys.argv) > 1 else "auto"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 24
body = int(sys.argv[4]) if len(sys.argv) > 4 else 0
if variant != "auto":
    os.environ["SOLO_STEP_VARIANT"] = variant
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solorl_b200.envs import SoloVecEnv  # noqa: E402

cfg = {"model_urdf": "solo12", "mode": "headless", "episode_length": 400, "frame_skip": 4, "control": "torque",
       "task": "walk", "num_history_stack": 1, "body_contacts": body}
env = SoloVecEnv(cfg, n, device="cuda:0", seed=1)
env.reset()
g = torch.Generator(device="cuda").manual_seed(5)
acts = [torch.rand(n, env.sim.act_dim, device="cuda", generator=g) * 2 - 1 for _ in range(8)]
for i in range(steps):
    env.sim.step(acts[i % 8])
torch.cuda.synchronize()
print("ok", env.sim.step_variant, float(env.sim.get_state().double().sum()))
env.close()
