"""A/B timing of the step-kernel register variants selected by SOLO_STEP_REGS (run on a GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_sweep import time_cfg
variants = sys.argv[1].split(",") if len(sys.argv) > 1 else ["255", "128"]
ns = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4096, 16384, 65536, 262144]
for n in ns:
    for v in variants:
        os.environ["SOLO_STEP_REGS"] = v
        ms, nc, sw, ssum = time_cfg("solo12", n, 8)
        print(f"solo12 n={n} variant={v}: {ms*1e3:8.1f} us/step {n/ms*1e3:.3e} env-steps/s contacts {nc:.2f} sweeps {sw:.1f} checksum {ssum:.6f}", flush=True)
