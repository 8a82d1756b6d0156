"""The configs of BASELINE.json / BASELINE.md §3 side by side: GPU env-steps/s (device resident, back to back) next
to the CPU restatement on all host threads and on one.  Run on a GPU box: python tests/tools/bench_configs.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.oracle import OracleVecEnv, lib  # noqa: E402  (CPU baseline leg)
from solorl_b200.abi import params_from_config  # noqa: E402
from solorl_b200.envs import SoloVecEnv  # noqa: E402
from solorl_b200.gait import SoloGaitVecEnv  # noqa: E402
from solorl_b200.model import SoloModel  # noqa: E402


def gpu_rate(cfg, n, steps=200):
    env = SoloVecEnv(cfg, n, device="cuda:0", seed=1)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = [torch.rand(n, env.sim.act_dim, device="cuda", generator=g) * 2 - 1 for _ in range(8)]
    for i in range(30):
        env.sim.step(acts[i % 8])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        env.sim.step(acts[i % 8])
    e1.record()
    torch.cuda.synchronize()
    env.close()
    return n * steps / (e0.elapsed_time(e1) * 1e-3)


def cpu_rate(cfg, nthreads, seconds=3.0):
    m = SoloModel.resolve(cfg["model_urdf"])
    p = params_from_config(cfg, m)
    n = 32 * nthreads
    v = OracleVecEnv(m, p, n, seed=1, nthreads=nthreads)
    v.reset()
    rng = np.random.default_rng(1)
    acts = [rng.uniform(-1, 1, size=(n, v.act_dim)).astype(np.float32) for _ in range(4)]
    v.step(acts[0])
    t0, k = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        v.step(acts[k % 4]); k += 1
    return n * k / (time.perf_counter() - t0)


def main():
    base = {"mode": "headless", "episode_length": 400, "frame_skip": 4, "flat_ground": True}
    rows = [("1 basic.yaml shape", dict(base, model_urdf="solo8", task="walk", control="torque", num_history_stack=1), 64),
            ("1b Solo8 stand", dict(base, model_urdf="solo8", task="stand", control="torque", num_history_stack=1), 64),
            ("2 Solo12 walk (bench.py)", dict(base, model_urdf="solo12", task="walk", control="torque", num_history_stack=1), 4096),
            ("2b Solo12 pointgoal (basic12.yaml)", dict(base, model_urdf="solo12", task="pointgoal", control="torque", num_history_stack=1), 4096),
            ("3 basic_pd.yaml", dict(base, model_urdf="solo8", task="stand", control="pd", gains=[5., .2], num_history_stack=0), 16384)]
    nthr = len(os.sched_getaffinity(0))
    print(f"# {torch.cuda.get_device_name(0)}; CPU restatement (fp64 oracle, not PyBullet) on {nthr} host threads and on 1")
    print(f"# {'config':38s} {'envs':>6s} {'GPU env-steps/s':>16s} {'CPU all threads':>16s} {'CPU 1 thread':>13s} {'GPU/CPU-all':>11s}")
    for name, cfg, n in rows:
        g = gpu_rate(cfg, n)
        c = cpu_rate(cfg, nthr)
        c1 = cpu_rate(cfg, 1, seconds=2.0)
        print(f"  {name:38s} {n:6d} {g:16.3e} {c:16.3e} {c1:13.3e} {g / c:11.0f}", flush=True)
    # config 4: the gait-env shell (one RL step = 80 controller ticks of 0.002 s, posture-hold stand-in controller)
    cfg = {"solo12": True, "episode_length": 50, "vel_switch": 1000, "mode": "headless", "num_history_stack": 1,
           "flat_ground": True, "auto_vel_switch": True}
    env = SoloGaitVecEnv(cfg, 4096, seed=1)
    env.reset()
    a = torch.randint(0, 9, (4096,), device="cuda")
    env.step(a)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    k = 5
    for _ in range(k):
        env.step(a)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"  {'4 basic_contact.yaml shell':38s} {4096:6d} {4096 * k / dt:16.3e} RL steps/s = {4096 * k * 80 / dt:.3e} simulator ticks/s "
          f"(controller in torch, one launch per tick; no CPU counterpart: the MPC is external)")
    env.close()


if __name__ == "__main__":
    main()
