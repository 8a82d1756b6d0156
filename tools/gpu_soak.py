"""Soak: long random-action rollouts, counting non-finite-state guard hits and checking ranges (GPU box)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solorl_b200.envs import SoloVecEnv

for robot, task, control, body in (("solo12", "walk", "torque", 0), ("solo8", "stand", "pd", 0), ("solo12", "pointgoal", "vpd", 0),
                                   ("solo12", "walk", "torque", 1), ("solo8", "stand", "pd", 1)):
    cfg = {"model_urdf": robot, "mode": "headless", "episode_length": 400, "frame_skip": 4, "control": control,
           "task": task, "num_history_stack": 1, "gains": [5., .2], "body_contacts": body}
    n = 4096
    env = SoloVecEnv(cfg, n, device="cuda:0", seed=11)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(3)
    nan = 0; eps = 0; mxq = 0.0; mxv = 0.0; lens = []; zmin = 1.0
    for t in range(1500):
        a = torch.randn(n, env.sim.act_dim, device="cuda", generator=g) * (3.0 if t % 7 == 0 else 1.0)
        obs, rew, done, infos = env.step(a)
        if t % 25 == 0:
            assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
            s = env.sim.get_state()
            nj = env.sim.nj
            mxq = max(mxq, float(s[:, 13:13 + nj].abs().max())); mxv = max(mxv, float(s[:, 13 + nj:].abs().max()))
            zmin = min(zmin, float(s[:, 2].min()))
            r = infos.done_records()
            nan += int(r["nan"].sum()); eps += len(r); lens += r["episode_length"].tolist()
    print(f"{robot}/{task}/{control}/body_contacts={body}: lowest base height {zmin:.3f} m, sampled {eps} finished episodes, nan-guard hits {nan}, mean length {np.mean(lens):.1f}, "
          f"max |q| {mxq:.2f} rad, max |qd| {mxv:.1f} rad/s", flush=True)
    env.close()
