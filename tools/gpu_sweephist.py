"""Distribution of PGS sweeps per contact substep (per env and per warp of 8 envs) on bench-like states."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solorl_b200.envs import SoloVecEnv

n = 4096
cfg = {"model_urdf": "solo12", "mode": "headless", "episode_length": 400, "frame_skip": 4,
       "control": "torque", "task": "walk", "num_history_stack": 1}
env = SoloVecEnv(cfg, n, device="cuda:0", seed=1)
env.reset()
g = torch.Generator(device="cuda").manual_seed(5)
allsw, allnc = [], []
for t in range(300):
    a = torch.rand(n, 12, device="cuda", generator=g) * 2 - 1
    env.sim.step(a)
    if t % 10 == 9:
        tau = torch.zeros(n, 12, device="cuda")
        env.sim.substep(tau)
        w = env.sim.get_work_counters().cpu().numpy()
        allnc.append(w[:, 0]); allsw.append(np.where(w[:, 0] > 0, w[:, 1] / np.maximum(w[:, 0], 1), 0))
nc = np.stack(allnc); sw = np.stack(allsw)
print("mean contacts", nc.mean(), "frac envs with contact", (nc > 0).mean())
print("per-env sweeps (contact envs): mean %.1f median %.1f p90 %.1f frac==50 %.3f" % (
    sw[nc > 0].mean(), np.median(sw[nc > 0]), np.percentile(sw[nc > 0], 90), (sw[nc > 0] >= 50).mean()))
hist = np.bincount(sw[nc > 0].astype(int), minlength=51)
print("hist", hist.tolist())
wm = sw.reshape(sw.shape[0], -1, 8).max(-1)
print("per-warp max sweeps: mean %.1f median %.1f frac==50 %.3f frac==0 %.3f" % (wm.mean(), np.median(wm), (wm >= 50).mean(), (wm == 0).mean()))
for k in (1, 2, 3, 4):
    s = sw[nc == k]
    if len(s): print(f"nc={k}: n={len(s)} mean {s.mean():.1f} frac50 {(s>=50).mean():.3f}")
