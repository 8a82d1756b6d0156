"""Timing sweep of the step kernel over its two builds and the batch size (run on a GPU box).
Usage: python tools/gpu_sweep.py [n1,n2,...] [latency,throughput,auto]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solorl_b200.envs import SoloVecEnv  # noqa: E402


def time_cfg(name, n, variant, task="walk", K=200, extra=None):
    if variant and variant != "auto":
        os.environ["SOLO_STEP_VARIANT"] = str(variant)
    else:
        os.environ.pop("SOLO_STEP_VARIANT", None)
    cfg = {"model_urdf": name, "mode": "headless", "episode_length": 400, "frame_skip": 4,
           "control": "torque", "task": task, "num_history_stack": 1}
    cfg.update(extra or {})
    env = SoloVecEnv(cfg, n, device="cuda:0", seed=1)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    acts = [torch.rand(n, env.sim.act_dim, device="cuda", generator=g) * 2 - 1 for _ in range(8)]
    for i in range(30):
        env.sim.step(acts[i % 8])
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(K):
            env.sim.step(acts[i % 8])
        ev1.record()
        torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1) / K)
    w = env.sim.get_work_counters().float().mean(0)
    state_sum = float(env.sim.get_state().double().sum().item())
    env.close()
    return best, w[0].item() / 4, (w[1] / w[0].clamp_min(1e-9)).item(), state_sum


def main():
    ns = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4096, 16384, 65536]
    epws = sys.argv[2].split(",") if len(sys.argv) > 2 else ["latency", "throughput"]
    print(torch.cuda.get_device_name(0), flush=True)
    for name in (os.environ.get("SOLO_SWEEP_ROBOTS", "solo12,solo8").split(",")):
        for n in ns:
            for epw in epws:
                ms, nc, sw, ssum = time_cfg(name, n, epw)
                print(f"{name} n={n} build={epw}: {ms * 1e3:8.1f} us/step  {n / ms * 1e3:.3e} env-steps/s  "
                      f"contacts/substep {nc:.2f} sweeps {sw:.1f} state-checksum {ssum:.6f}", flush=True)


if __name__ == "__main__":
    main()
