"""Profiling target: the GAE reverse scan (solo_gae -> gae_chunked_kernel) at T = 400, N = 4096, L2 flushed
between launches (run under ncu).  Usage: python tools/gpu_prof_gae.py [T=400] [N=4096] [reps=4]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solorl_b200.sim import gae  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 400
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
r = torch.randn(T, N, device=dev, generator=g)
v = torch.randn(T + 1, N, device=dev, generator=g)
m = (torch.rand(T + 1, N, device=dev, generator=g) > 0.02).float()
ret = torch.zeros(T + 1, N, device=dev)
flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
for i in range(reps):
    flush.fill_(float(i))
    gae(r, v, m, ret, 0.99, 0.95, True)
torch.cuda.synchronize()
print("ok", float(ret.double().sum()))
