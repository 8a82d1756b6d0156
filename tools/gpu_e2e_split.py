"""Where the host-buffer step's time goes (solo_step_host against the device-resident step), 4096 envs.
Usage: python tools/gpu_e2e_split.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solorl_b200.envs import SoloVecEnv  # noqa: E402

cfg = {"model_urdf": "solo12", "mode": "headless", "episode_length": 400, "frame_skip": 4, "control": "torque",
       "task": "walk", "num_history_stack": 1}
n, K = 4096, 300
env = SoloVecEnv(cfg, n, device="cuda:0", seed=1)
env.reset()
sim = env.sim
g = torch.Generator(device="cuda").manual_seed(5)
acts = [torch.rand(n, sim.act_dim, device="cuda", generator=g) * 2 - 1 for _ in range(4)]
for i in range(30):
    sim.step(acts[i % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(K):
    sim.step(acts[i % 4])
e1.record(); torch.cuda.synchronize()
print(f"device-resident, back to back, no host sync      : {e0.elapsed_time(e1) / K * 1e3:7.1f} us per step")
t0 = time.perf_counter()
for i in range(K):
    sim.step(acts[i % 4]); torch.cuda.synchronize()
print(f"device-resident + one stream synchronise per step: {(time.perf_counter() - t0) / K * 1e6:7.1f} us per step")
h_act = [a.cpu().pin_memory() for a in acts]
h_obs = torch.empty(n, sim.d).pin_memory(); h_rew = torch.empty(n).pin_memory(); h_done = torch.empty(n).pin_memory()
pa = [t.data_ptr() for t in h_act]
for mode in ("1", "0"):
    os.environ["SOLO_HOST_ZERO_COPY"] = mode
    e2 = SoloVecEnv(cfg, n, device="cuda:0", seed=1); e2.reset()
    for i in range(10):
        e2.sim.step_host_ptr(pa[i % 4], h_obs.data_ptr(), h_rew.data_ptr(), h_done.data_ptr())
    t0 = time.perf_counter()
    for i in range(K):
        e2.sim.step_host_ptr(pa[i % 4], h_obs.data_ptr(), h_rew.data_ptr(), h_done.data_ptr())
    dt = (time.perf_counter() - t0) / K * 1e6
    print(f"solo_step_host, pinned buffers, {'kernel reads / writes them directly' if mode == '1' else 'staged copies':36s}: {dt:7.1f} us per step")
    e2.close()
