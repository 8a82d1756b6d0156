"""Per-substep phase timing of the narrow step kernel from in-kernel cycle stamps (-DSOLO_TRACE build under tools/_ab/:
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DSOLO_TRACE
     -o tools/_ab/libsolo_trace.so solorl_b200/csrc/solo_kernels.cu)."""
import ctypes as C, os, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["SOLO_B200_LIB"] = os.path.join(root, "tools", "_ab", "libsolo_trace.so")
os.environ["SOLO_STEP_VARIANT"] = sys.argv[2] if len(sys.argv) > 2 else "latency"
sys.path.insert(0, root)
import torch
from solorl_b200 import _lib
from solorl_b200.envs import SoloVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = {"model_urdf": "solo12", "mode": "headless", "episode_length": 400, "frame_skip": 4, "control": "torque", "task": "walk", "num_history_stack": 1}
env = SoloVecEnv(cfg, n, device="cuda:0", seed=1); env.reset()
g = torch.Generator(device="cuda").manual_seed(5)
acts = [torch.rand(n, 12, device="cuda", generator=g) * 2 - 1 for _ in range(8)]
L = _lib.lib(); acc = []; kacc = []
for i in range(40):
    env.sim.step(acts[i % 8])
    if i >= 20:
        torch.cuda.synchronize()
        buf = np.zeros((1024, 8, 8), dtype=np.int64)
        assert L.solo_debug_narrow_trace(buf.ctypes.data_as(C.c_void_p)) == 0
        nb = (n + 31) // 32
        acc.append(buf[:nb, :4, :6].copy())
        kacc.append(buf[:nb, 4, :4].copy())
a = np.stack(acc).astype(np.float64)       # [steps, blocks, substep, stamp]
d = np.diff(a, axis=3)
names = ["ABA (inward, base solve, outward)", "contact_setup", "limit select/setup + assembly + init", "PGS sweeps", "impulses + integrate"]
print(f"{os.environ['SOLO_STEP_VARIANT']} build, {n} envs: cycles per phase, mean over blocks and steps; substeps 0..3")
print("  (stamps 3 / 4 are only written when warp 0 of the block enters the solve: blocks whose first warp holds no contact and no\n   limit row keep stale values there, so read the per-phase rows as indicative and the kernel-level rows below as exact)")
for k, nm in enumerate(names):
    print(f"  {nm:40s} " + " ".join(f"{d[:, :, s, k].mean():9.0f}" for s in range(4)))
print(f"  {'substep total':40s} " + " ".join(f"{(a[:, :, s, 5] - a[:, :, s, 0]).mean():9.0f}" for s in range(4)))
print(f"  {'gap to the next substep':40s} " + " ".join(f"{(a[:, :, s + 1, 0] - a[:, :, s, 5]).mean():9.0f}" for s in range(3)))
print("  whole substep loop, mean / max over blocks:", (a[:, :, 3, 5] - a[:, :, 0, 0]).mean(), (a[:, :, 3, 5] - a[:, :, 0, 0]).max(axis=1).mean())
k = np.stack(kacc).astype(np.float64)      # [steps, blocks, stamp]: kernel entry, after step_load, before step_finish, end
print(f"  prologue (entry -> end of step_load)      mean {(k[:, :, 1] - k[:, :, 0]).mean():9.0f}  max {(k[:, :, 1] - k[:, :, 0]).max(axis=1).mean():9.0f}")
print(f"  substep loop                              mean {(k[:, :, 2] - k[:, :, 1]).mean():9.0f}  max {(k[:, :, 2] - k[:, :, 1]).max(axis=1).mean():9.0f}")
print(f"  epilogue (step_finish)                    mean {(k[:, :, 3] - k[:, :, 2]).mean():9.0f}  max {(k[:, :, 3] - k[:, :, 2]).max(axis=1).mean():9.0f}")
print(f"  whole block                               mean {(k[:, :, 3] - k[:, :, 0]).mean():9.0f}  max {(k[:, :, 3] - k[:, :, 0]).max(axis=1).mean():9.0f}")
slow = (k[:, :, 3] - k[:, :, 0]).argmax(axis=1)
idx = np.arange(k.shape[0])
print(f"  slowest block of each step: prologue {(k[idx, slow, 1] - k[idx, slow, 0]).mean():.0f}, loop {(k[idx, slow, 2] - k[idx, slow, 1]).mean():.0f}, "
      f"epilogue {(k[idx, slow, 3] - k[idx, slow, 2]).mean():.0f}")
