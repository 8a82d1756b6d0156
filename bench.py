#!/usr/bin/env python
"""bench.py — env-steps/sec of the batched Solo12 env step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
  python bench.py --impl reference --gpus N --steps K ...   (CPU arm: the reference's path)

Workload (BASELINE.json configs[1]): Solo12, task walk, torque control, num_history_stack 1,
frame_skip 4 at 1/240 s, 4096 envs per GPU, random actions U(-1,1)^12, auto-reset on.
A "step" is one vectorised env step of all envs of a rank (= 4 physics substeps + observation +
reward + termination + auto-reset per env).  `value` times the step with actions already in
HBM; `e2e` times the same step through the host-buffer C-ABI call (solo_step_host: H2D actions,
kernel, D2H obs/reward/done, sync).  N > 1: one process per GPU (torchrun), independent env
shards, no data-path collective ("weak" scaling); time = max over ranks.

The reference's own arithmetic for this path is PyBullet, which is not installable in this
image, so the CPU arm (`--impl reference`, and the `cpu_baseline` object) times this repo's
double-precision CPU restatement (oracle/, kind "port") on all host threads.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 4096
CONFIG = {"model_urdf": "solo12", "mode": "headless", "episode_length": 400, "frame_skip": 4,
          "control": "torque", "task": "walk", "num_history_stack": 1, "flat_ground": True}
WORKLOAD = "configs/basic12.yaml-shaped: Solo12 walk, torque control, H=1, frame_skip 4, random actions"


def algorithmic_flops_per_env_step(nj, nc_sum, sweep_feet, frame_skip=4):
    """F_step = S (F_ABA + F_int) + sum_substeps (F_setup + F_PGS) + 400 with m = 3 nc
    rows per substep.  nc_sum = sum over the S substeps of feet in contact; sweep_feet = sum over
    substeps of feet in contact x PGS sweeps actually run (K = 50 only when the residual test never
    fires), both measured by the kernel (solo_get_work_counters)."""
    nb, ndof = nj + 1, nj + 6
    f_aba = 429 * nb - 502
    f_setup = 3.0 * nc_sum * (250 * nb + 2 * ndof)
    f_pgs = 3.0 * sweep_feet * (4 * ndof + 10)
    f_int = 80 * nb
    return frame_skip * (f_aba + f_int) + f_setup + f_pgs + 400


def algorithmic_bytes_per_env_step(nj, D):
    """read state(13+2nj)+action(nj)+contact(4)+scalars(8+3); write state+obs(D)+
    reward+done+contact+scalars, fp32."""
    rd = (16 + 2 * nj) + nj + 4 + 12
    wr = (16 + 2 * nj) + D + 2 + 4 + 12
    return 4 * (rd + wr)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for k, nme in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group("nccl")
        else:
            dist.init_process_group("gloo")
    return rank, world, local


def host_threads():
    """All host threads this process may run on.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to
    every rank, which would time the CPU arm on a single thread at N > 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(seconds=12.0, nthreads=None):
    """The oracle port on all host threads over a bounded sample of the same workload."""
    from oracle.oracle import OracleVecEnv, lib
    from solorl_b200.abi import params_from_config
    from solorl_b200.model import SoloModel
    m = SoloModel.resolve(CONFIG["model_urdf"])
    p = params_from_config(CONFIG, m)
    nthreads = nthreads or host_threads()
    n = 32 * nthreads
    v = OracleVecEnv(m, p, n, seed=1, nthreads=nthreads)
    v.reset()
    rng = np.random.default_rng(1)
    acts = [rng.uniform(-1, 1, size=(n, v.act_dim)).astype(np.float32) for _ in range(8)]
    for i in range(3):
        v.step(acts[i])
    t0, steps = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        v.step(acts[steps % 8])
        steps += 1
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": "env-steps/s", "cores": nthreads, "kind": "port",
            "sample": f"{n} envs x {steps} steps ({dt:.1f} s) of the same workload; fp64 CPU restatement "
                      f"(oracle/), not PyBullet (not installable in this image)"}


def run_reference(args):
    """CPU arm.  Rank 0 alone runs; a step = one vec step of a bounded sample (32 envs per thread)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle import OracleVecEnv, lib
    from solorl_b200.abi import dims, params_from_config
    from solorl_b200.model import SoloModel
    m = SoloModel.resolve(CONFIG["model_urdf"])
    p = params_from_config(CONFIG, m)
    nthreads = host_threads()
    n = 32 * nthreads
    v = OracleVecEnv(m, p, n, seed=1, nthreads=nthreads)
    v.reset()
    rng = np.random.default_rng(1)
    acts = [rng.uniform(-1, 1, size=(n, v.act_dim)).astype(np.float32) for _ in range(8)]
    for i in range(args.warmup):
        v.step(acts[i % 8])
    t0 = time.perf_counter()
    for i in range(args.steps):
        v.step(acts[i % 8])
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    sample = (f"{n} envs per step on {nthreads} host threads (bounded sample of the {ENVS_PER_GPU}-env workload); "
              f"fp64 CPU restatement (oracle/), PyBullet itself is not installable in this image")
    line = {"impl": "reference", "metric": "env-steps/sec", "value": val, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_step": n, "robot": "solo12", "task": "walk"},
            "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": nthreads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "pybullet_direct": None, "reason": "pybullet not installed"}
    print(json.dumps(line), flush=True)


def time_policy_rollout(cfg, n, dev, rank, world, barrier, T=32, reps=10):
    """env-steps/s of policy rollouts: Policy (agents/ppo/policy.py layout, init_layer orthogonal gain sqrt 2,
    log-std 0), torch.manual_seed(1), stochastic sampling, rollout buffer appends included."""
    import torch
    from solorl_b200.agents.policy import Policy
    from solorl_b200.agents.storage import OPBuffer
    from solorl_b200.agents.train import EpisodeTracker, Rollout
    from solorl_b200.envs import make_vec_envs
    torch.manual_seed(1)
    envs = make_vec_envs(cfg, n, device=dev, seed=2, env_id_offset=rank * n)
    ac = Policy(envs.observation_space.shape, envs.action_space, None, {"hidden_size": 64}).to(dev)
    buf = OPBuffer(T, n, envs.observation_space.shape, envs.action_space.shape[0], dev)
    buf.obs[0].copy_(envs.reset())
    ro = Rollout(envs, ac, buf, EpisodeTracker(dev), T, use_graph=True)
    for _ in range(3):
        ro(); buf.reset()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ro(); buf.reset()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    envs.close()
    return {"value": world * n * T * reps / (ms * 1e-3), "unit": "env-steps/s", "ms_per_step": ms / (T * reps),
            "rollout_steps": T, "cuda_graph": ro.graph is not None,
            "what": "policy act + env step + rollout-buffer append per step, device resident"}


def time_saturated(cfg, dev, n, steps=60):
    """Device-resident env-steps/s of the same workload at a batch that fills the GPU (several resident waves
    of the 16-warps-per-SM throughput build), with the work counters of that run."""
    import torch
    from solorl_b200.envs import SoloVecEnv
    env = SoloVecEnv(cfg, n, device=dev, seed=3)
    env.reset()
    g = torch.Generator(device=dev).manual_seed(7)
    acts = [torch.rand(n, env.sim.act_dim, device=dev, generator=g) * 2 - 1 for _ in range(4)]
    ncs, sws = [], []
    for i in range(12):
        env.sim.step(acts[i % 4])
        w = env.sim.get_work_counters().float().mean(0)
        ncs.append(w[0].item()); sws.append(w[1].item())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        env.sim.step(acts[i % 4])          # state (n x 0.8 KB = 52 MB at 65536 envs) + obs exceed what L2 keeps hot
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    nj = env.sim.nj
    env.close()
    return {"envs": n, "value": n / (ms * 1e-3), "unit": "env-steps/s", "ms_per_step": ms, "step_kernel_build": "throughput",
            "flops_per_env_step": algorithmic_flops_per_env_step(nj, float(np.mean(ncs)), float(np.mean(sws)))}


def run_ours(args):
    import torch
    from solorl_b200 import _lib, build
    from solorl_b200.envs import SoloVecEnv
    rank, world, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    _lib.lib()   # fail loudly if the extension is missing
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = args.envs
    cfg = dict(CONFIG)
    env = SoloVecEnv(cfg, n, device=dev, seed=args.seed, env_id_offset=rank * n)
    sim = env.sim
    nj, A, D = sim.nj, sim.act_dim, sim.d
    env.reset()
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    nact = 16
    acts = [torch.rand(n, A, device=dev, generator=g) * 2 - 1 for _ in range(nact)]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    ncs, sweeps, resets = [], [], []

    def sample_work(d):
        w = sim.get_work_counters().float().mean(0)
        ncs.append(w[0].item())
        sweeps.append(w[1].item())
        resets.append(d.mean().item())

    for i in range(max(args.warmup, 3)):
        _, _, d, _ = env.step(acts[i % nact])
        sample_work(d)

    # ---- device-resident timing: exactly K steps, L2 flushed between timed steps -------------
    sampler = ClockSampler(local)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.start()
    launches0 = sim.launch_count
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(float(i))
        evs[i][0].record()
        sim.step(acts[i % nact])
        evs[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = sim.launch_count - launches0
    clocks = sampler.stop()
    step_ms = np.array([a.elapsed_time(b) for a, b in evs])
    total_ms = float(step_ms.sum())
    for i in range(16):
        _, _, d, _ = env.step(acts[i % nact])
        sample_work(d)
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * n * args.steps / (total_ms * 1e-3)

    # ---- back-to-back (no flush) timing of the same K steps, for reference ---------------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        sim.step(acts[i % nact])
    e1.record()
    barrier()
    b2b_ms = e0.elapsed_time(e1) / args.steps

    # ---- end to end through the host-buffer C-ABI call -----------------------------------------
    h_act = [torch.empty(n, A, dtype=torch.float32).pin_memory() for _ in range(4)]
    for i, t in enumerate(h_act):
        t.copy_(acts[i].cpu())
    h_obs = torch.empty(n, D, dtype=torch.float32).pin_memory()
    h_rew = torch.empty(n, dtype=torch.float32).pin_memory()
    h_done = torch.empty(n, dtype=torch.float32).pin_memory()
    p_act = [t.data_ptr() for t in h_act]                 # pinned host addresses, reused every step
    p_obs, p_rew, p_done = h_obs.data_ptr(), h_rew.data_ptr(), h_done.data_ptr()
    for i in range(3):
        sim.step_host_ptr(p_act[i % 4], p_obs, p_rew, p_done)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        sim.step_host_ptr(p_act[i % 4], p_obs, p_rew, p_done)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * n * args.steps / e2e_s

    # ---- policy rollout (reference-shaped MLP 64-64 tanh policy, stochastic actions), one CUDA graph of
    #      T x (act -> env step -> buffer append); reported next to the random-action figure -------------
    policy_rollout = None
    if not args.no_policy_rollout:
        try:
            policy_rollout = time_policy_rollout(cfg, n, dev, rank, world, barrier)
        except Exception as e:      # the headline metric does not depend on it
            policy_rollout = {"error": repr(e)[:200]}

    # ---- the same kernel when the batch fills the machine (throughput build, 65536 envs): how far the code is
    #      from the FP32 peak once it is no longer limited by the 4096-env batch of the headline workload ------
    saturated = None
    if not args.no_saturated and world == 1:
        try:
            saturated = time_saturated(cfg, dev, args.saturated_envs)
        except Exception as e:
            saturated = {"error": repr(e)[:200]}

    if rank != 0:
        env.close()
        return

    # ---- roofline of the dominant kernel (step_kernel: one launch per step) ----------------------
    nc_sum, sweep_feet = float(np.mean(ncs)), float(np.mean(sweeps))
    flops = algorithmic_flops_per_env_step(nj, nc_sum, sweep_feet) * n
    kernel_ms = total_ms / args.steps if world == 1 else float(step_ms.mean())
    fma_tflops = C.c_double()
    fma_ms = C.c_double()
    bl = C.CDLL(build.BENCH_LIB)
    bl.solo_bench_fma_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    rc = bl.solo_bench_fma_peak(local, C.byref(fma_tflops), C.byref(fma_ms))
    peak = fma_tflops.value if rc == 0 else 74.5
    achieved = flops / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    abytes = algorithmic_bytes_per_env_step(nj, D) * n
    roofline = {
        "bound": "fp32", "kernel": "step_kernel<3>", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture of this workload
        # (profiles/r1j_step_kernel_ncu_full.txt); only valid for the default 4096-env Solo12 workload
        "traffic": 1298432 if (n == ENVS_PER_GPU) else None,
        "traffic_unit": "bytes per launch (algorithmic: %d)" % (algorithmic_bytes_per_env_step(nj, D) * n),
        "peak_source": "FP32 FMA microbenchmark measured live in this run (solo_bench_fma_peak)" if rc == 0
        else "fallback 148 SM x 128 lanes x 2 x 1.965 GHz",
        "algorithmic_flops_per_env_step": algorithmic_flops_per_env_step(nj, nc_sum, sweep_feet),
        "mean_contacts_per_substep": nc_sum / 4.0,
        "mean_pgs_sweeps_per_contact_substep": (sweep_feet / nc_sum) if nc_sum > 0 else 0.0,
        "resets_per_env_step": float(np.mean(resets)), "kernel_ms": kernel_ms,
        "hbm": {"bound": "hbm", "achieved": abytes / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": abytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                "algorithmic_bytes_per_env_step": algorithmic_bytes_per_env_step(nj, D),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"},
    }
    line = {
        "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": n, "robot": "solo12", "task": "walk", "control": "torque",
                   "num_history_stack": 1, "episode_length": 400, "solver_iters": 50,
                   "solver_residual_threshold": 1e-7, "reset_mode": "cached", "step_kernel_build": "latency" if n <= 8192 else "throughput",
                   "l2": "flushed (256 MiB write) between timed steps; per-step CUDA events summed",
                   "parallelism": f"env-shard x{world}"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": n * A * 4,
                "d2h_bytes_per_step": n * (D + 2) * 4, "ms_per_step": e2e_s / args.steps * 1e3,
                "api": "solo_step_host (pinned host buffers read and written by the step kernel through their mapped alias, one sync per step)"},
        "roofline": roofline,
        "back_to_back_ms_per_step": b2b_ms, "wall_ms_per_step_incl_flush": t_wall / args.steps * 1e3,
        "policy_rollout": policy_rollout,
        "saturated": saturated,
    }
    if saturated and "flops_per_env_step" in saturated:
        ach = saturated["flops_per_env_step"] * saturated["value"] / 1e12
        saturated.update({"achieved_tflops": ach, "peak_tflops": peak, "frac": ach / peak})
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.cpu_seconds)
        try:                                    # the reference itself, if it ever becomes runnable here
            import pybullet  # noqa: F401
            line["pybullet_direct"] = "installed but not wired: see BASELINE.md B0"
        except Exception:
            line["pybullet_direct"], line["reason"] = None, "pybullet not installed"
    print(json.dumps(line), flush=True)
    env.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU (BASELINE: 4096)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-policy-rollout", action="store_true")
    ap.add_argument("--no-saturated", action="store_true")
    ap.add_argument("--saturated-envs", type=int, default=65536)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        try:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
