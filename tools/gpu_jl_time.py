import os, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
from gpu_sweep import time_cfg
for n in (4096, 65536):
    for jl in (0, 1):
        ms, nc, sw, ssum = time_cfg("solo12", n, "auto", extra={"joint_limits": jl})
        print(f"n={n} joint_limits={jl}: {ms*1e3:.1f} us/step {n/ms*1e3:.3e} sweeps {sw:.1f}", flush=True)
