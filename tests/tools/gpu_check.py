"""Quick GPU sanity run: parity of the CUDA path against the CPU oracle and a rough timing.
Usage (on a GPU box): python tests/tools/gpu_check.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.oracle import OracleEnv, default_params  # noqa: E402
from solorl_b200.abi import params_from_config  # noqa: E402
from solorl_b200.envs import SoloVecEnv  # noqa: E402
from solorl_b200.model import SoloModel  # noqa: E402
from solorl_b200.sim import SoloSim  # noqa: E402


def rand_states(rng, n, nj):
    s = np.zeros((n, 13 + 2 * nj))
    s[:, :3] = rng.normal(size=(n, 3)) * 0.3
    s[:, 2] += 1.0
    q = rng.normal(size=(n, 4))
    s[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    s[:, 7:10] = rng.normal(size=(n, 3))
    s[:, 10:13] = rng.normal(size=(n, 3)) * 2
    s[:, 13:13 + nj] = rng.uniform(-2, 2, size=(n, nj))
    s[:, 13 + nj:] = rng.normal(size=(n, nj)) * 5
    return s.astype(np.float32)


def main():
    rng = np.random.default_rng(0)
    print(torch.cuda.get_device_name(0))
    for name in ("solo8", "solo12"):
        model = SoloModel.builtin(name)
        p = default_params()
        n = 256
        sim = SoloSim(model, p, n, device=0)
        nj = sim.nj
        # forward dynamics
        s = rand_states(rng, n, nj)
        tau = rng.uniform(-3, 3, size=(n, nj)).astype(np.float32)
        qdd = sim.forward_dynamics(torch.from_numpy(s).cuda(), torch.from_numpy(tau).cuda()).cpu().numpy()
        o = OracleEnv(model, p)
        worst = 0
        for i in range(n):
            o.set_state(s[i].astype(np.float64))
            ref = o.forward_dynamics(tau[i].astype(np.float64))
            worst = max(worst, np.linalg.norm(ref - qdd[i]) / np.linalg.norm(ref))
        print(f"{name}: forward dynamics worst rel err {worst:.3e}")
        # contact substep from stance states
        st = np.zeros((n, 13 + 2 * nj), np.float32)
        st[:, 2] = 0.24
        st[:, 6] = 1
        njl = nj // 4
        for l in range(4):
            sg = 1 if l < 2 else -1
            st[:, 13 + l * njl + njl - 2] = 0.8 * sg + rng.normal(size=n) * 0.1
            st[:, 13 + l * njl + njl - 1] = -1.6 * sg + rng.normal(size=n) * 0.1
        sim.set_state(torch.from_numpy(st).cuda())
        errs = []
        cur = st.copy()
        for t in range(30):
            tau = (rng.uniform(-1, 1, size=(n, nj)) * 1.0).astype(np.float32)
            sim.set_state(torch.from_numpy(cur).cuda())
            sim.substep(torch.from_numpy(tau).cuda())
            nxt = sim.get_state().cpu().numpy()
            con = sim.get_contacts().cpu().numpy()
            for i in range(0, n, 16):
                o.set_state(cur[i].astype(np.float64))
                o.substep(tau[i].astype(np.float64))
                ref = o.get_state()
                errs.append((np.abs(ref - nxt[i]) / np.maximum(1, np.abs(ref))).max())
                assert (o.get_contacts()[:, 1] == con[i, :, 1]).all()
            cur = nxt
        print(f"{name}: contact substep err max {max(errs):.3e} median {np.median(errs):.3e}, "
              f"mean contacts {con[:, :, 1].sum(1).mean():.2f}")
        sim.close()

    # full env steps vs oracle (cached reset), all tasks
    for name, task, control, H in (("solo8", "walk", "torque", 1), ("solo12", "pointgoal", "torque", 1),
                                   ("solo8", "stand", "pd", 0), ("solo12", "walk", "torque", 2)):
        cfg = {"model_urdf": name, "mode": "headless", "episode_length": 30, "frame_skip": 4,
               "control": control, "task": task, "num_history_stack": H, "gains": [5., .2]}
        n = 64
        env = SoloVecEnv(cfg, n, device="cuda:0", seed=3)
        obs = env.reset().cpu().numpy()
        ors = [OracleEnv(env.model, env.params, seed=3, env_id=i) for i in range(8)]
        worst_obs = max(np.abs(ors[i].reset() - obs[i]).max() for i in range(8))
        worst_rew = 0
        ndone = 0
        mism = 0
        for t in range(45):
            a = rng.uniform(-1.5, 1.5, size=(n, env.sim.act_dim)).astype(np.float32)
            obs, rew, done, infos = env.step(torch.from_numpy(a).cuda())
            obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
            for i in range(8):
                oo, r, d, info = ors[i].step(a[i].astype(np.float64), auto_reset=True)
                if d != bool(done[i] > 0.5):
                    mism += 1
                    # resync
                    continue
                worst_obs = max(worst_obs, np.abs(oo - obs[i]).max())
                worst_rew = max(worst_rew, abs(r - rew[i]))
                ndone += d
                if d:
                    gi = infos[i]
                    assert gi["episode_length"] == info["episode_length"], (gi, info)
                    assert gi["success"] == bool(info["success"]) and gi["timeout"] == bool(info["timeout"])
        print(f"{name}/{task}/{control}/H{H}: obs err {worst_obs:.3e} reward err {worst_rew:.3e} "
              f"episodes {ndone} done-mismatch {mism}")
        env.close()

    # cached vs simulate reset: bitwise equal trajectories
    cfg = {"model_urdf": "solo12", "mode": "headless", "episode_length": 20, "frame_skip": 4,
           "control": "torque", "task": "pointgoal", "num_history_stack": 1}
    e1 = SoloVecEnv(cfg, 128, device="cuda:0", seed=5)
    e2 = SoloVecEnv(dict(cfg, reset_mode="simulate"), 128, device="cuda:0", seed=5)
    o1, o2 = e1.reset().clone(), e2.reset().clone()
    same = bool((o1 == o2).all())
    for t in range(50):
        a = torch.rand(128, 12, device="cuda") * 2 - 1
        r1 = e1.step(a)
        r2 = e2.step(a)
        same = same and bool((r1[0] == r2[0]).all()) and bool((r1[1] == r2[1]).all()) and bool((r1[2] == r2[2]).all())
    print("cached reset == simulated reset (bitwise):", same)
    e1.close(); e2.close()

    # rough timing
    for name, n, thr in (("solo12", 4096, 1e-7), ("solo12", 4096, 0.0), ("solo8", 4096, 1e-7), ("solo12", 65536, 1e-7),
                         ("solo12", 262144, 1e-7), ("solo12", 262144, 0.0)):
        cfg = {"model_urdf": name, "mode": "headless", "episode_length": 400, "frame_skip": 4,
               "control": "torque", "task": "walk", "num_history_stack": 1, "solver_residual_threshold": thr}
        env = SoloVecEnv(cfg, n, device="cuda:0", seed=1)
        env.reset()
        acts = [torch.rand(n, env.sim.act_dim, device="cuda") * 2 - 1 for _ in range(8)]
        for i in range(20):
            env.step(acts[i % 8])
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 200
        ev0.record()
        nd = 0
        for i in range(K):
            env.step(acts[i % 8])
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / K
        con = env.sim.get_contacts()[:, :, 1].sum(1).mean().item()
        print(f"{name} n={n} thr={thr}: {ms * 1e3:.1f} us/step -> {n / ms * 1e3:.3e} env-steps/s (mean contacts {con:.2f})")
        env.close()


if __name__ == "__main__":
    main()
