"""Phase timing of the wide step kernel from in-kernel cycle stamps (a -DSOLO_TRACE build of the library under
tools/_ab/).  Prints, per phase of the LAST substep of a step, the mean / max over blocks in cycles."""
import ctypes as C
import os
import sys

import numpy as np

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "tools", "_ab", "libsolo_trace.so")
os.environ["SOLO_B200_LIB"] = lib
os.environ["SOLO_STEP_VARIANT"] = "wide"
sys.path.insert(0, root)
import torch  # noqa: E402
from solorl_b200 import _lib  # noqa: E402
from solorl_b200.envs import SoloVecEnv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = {"model_urdf": "solo12", "mode": "headless", "episode_length": 400, "frame_skip": 4, "control": "torque",
       "task": "walk", "num_history_stack": 1}
env = SoloVecEnv(cfg, n, device="cuda:0", seed=1)
env.reset()
g = torch.Generator(device="cuda").manual_seed(5)
acts = [torch.rand(n, 12, device="cuda", generator=g) * 2 - 1 for _ in range(8)]
L = _lib.lib()
names = ["S0 publish -> bar1", "bar1 -> bar2 (P1 links)", "P2 recursion+base", "bar3 wait", "bar4 wait (P3 rows)",
         "load rows + solve + integrate"]
acc, hacc = [], []
for i in range(40):
    env.sim.step(acts[i % 8])
    if i >= 20:
        torch.cuda.synchronize()
        buf = np.zeros((1024, 8), dtype=np.int64)
        assert L.solo_debug_wide_trace(buf.ctypes.data_as(C.c_void_p)) == 0
        nb = (n + 31) // 32
        acc.append(np.diff(buf[:nb, :7], axis=1))
        hb = np.zeros((1024, 3, 8), dtype=np.int64)
        assert L.solo_debug_wide_trace_helpers(hb.ctypes.data_as(C.c_void_p)) == 0
        hacc.append(np.concatenate([np.diff(hb[:nb, :, :7], axis=2), (hb[:nb, :, 0:1] - buf[:nb, None, 0:1])], axis=2))
a = np.stack(acc)           # [steps, blocks, 6]
print(f"wide step kernel, {n} envs, last substep of each step, cycles (mean over blocks and steps / mean of per-step max over blocks)")
for k, nm in enumerate(names):
    print(f"  {nm:34s} {a[:, :, k].mean():9.0f} {a[:, :, k].max(axis=1).mean():9.0f}")
print(f"  {'substep total':34s} {a.sum(2).mean():9.0f} {a.sum(2).max(axis=1).mean():9.0f}")
raw7 = hb[:nb, :, 7] - hb[:nb, :, 1]
print("helper: stamp7 - stamp1 (first run of P1 incl. barrier wait) / stamp2 - stamp7 (second, warm run):",
      raw7.mean(0), (hb[:nb, :, 2] - hb[:nb, :, 7]).mean(0))
h = np.stack(hacc)          # [steps, blocks, role, 7]
hn = ["wait bar1", "P1 link work", "wait bar2", "wait bar3 (P2)", "P3 row work", "wait bar4", "(helper t0 - leg t0)"]
print("helper warpgroups (lane 0 of each), cycles, mean over blocks and steps: role 0 / 1 / 2")
for k, nm in enumerate(hn):
    print(f"  {nm:24s} " + " ".join(f"{h[:, :, r, k].mean():9.0f}" for r in range(3)))
