"""Per-sweep cost of the PGS loop: the headline workload with the residual exit off (every contact substep runs all
solver_iters sweeps), timed for two sweep counts; the difference / (4 substeps x delta sweeps) is the cycles per sweep.
Usage: python tools/gpu_fixed_sweeps.py [n=4096]   (SOLO_B200_LIB selects the build)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gpu_sweep import time_cfg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
res = {}
for iters in (10, 50):
    ms, nc, sw, _ = time_cfg("solo12", n, "latency", K=100, extra={"solver_residual_threshold": 0.0, "solver_iters": iters})
    res[iters] = ms
    print(f"solver_iters={iters}: {ms * 1e3:.1f} us/step (contacts/substep {nc:.2f}, sweeps {sw:.1f})", flush=True)
per_sweep_us = (res[50] - res[10]) * 1e3 / (4 * 40)
print(f"per sweep: {per_sweep_us * 1e3:.0f} ns = {per_sweep_us * 1965:.0f} cycles at 1965 MHz; fixed part {res[10] * 1e3 - 40 * per_sweep_us:.1f} us/step")
