"""Shared helpers of the test-suite."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


# ---- shared helpers -----------------------------------------------------------------------
def make_config(robot="solo12", task="stand", control="torque", H=1, episode_length=400, **kw):
    cfg = {"model_urdf": robot, "mode": "headless", "episode_length": episode_length, "frame_skip": 4,
           "control": control, "task": task, "num_history_stack": H, "flat_ground": True}
    if control in ("pd", "fpd", "fixed_pd"):
        cfg["gains"] = [5., .2]     # configs/basic_pd.yaml:6
    cfg.update(kw)
    return cfg


def random_states(rng, n, nj, vel_scale=1.0):
    """Free-flight states: z ~ 1 m, random attitude, random velocities."""
    s = np.zeros((n, 13 + 2 * nj))
    s[:, :3] = rng.normal(size=(n, 3)) * 0.3
    s[:, 2] += 1.0
    q = rng.normal(size=(n, 4))
    s[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    s[:, 7:10] = rng.normal(size=(n, 3)) * vel_scale
    s[:, 10:13] = rng.normal(size=(n, 3)) * 2 * vel_scale
    s[:, 13:13 + nj] = rng.uniform(-2, 2, size=(n, nj))
    s[:, 13 + nj:] = rng.normal(size=(n, nj)) * 5 * vel_scale
    return s.astype(np.float32).astype(np.float64)   # exactly representable in fp32


def stance_states(rng, n, nj, z=0.24, noise=0.1):
    """Bent-leg stance near the SRDF nominal posture (srdf/solo.srdf:69-83: z 0.235, HFE +-0.8,
    KFE -+1.6): feet on or near the ground, well-conditioned contact problem."""
    s = np.zeros((n, 13 + 2 * nj))
    s[:, 2] = z + rng.normal(size=n) * 0.005
    s[:, 6] = 1.0
    njl = nj // 4
    for l in range(4):
        sg = 1.0 if l < 2 else -1.0
        if njl == 3:
            s[:, 13 + l * 3] = rng.normal(size=n) * noise * 0.5
        s[:, 13 + l * njl + njl - 2] = 0.8 * sg + rng.normal(size=n) * noise
        s[:, 13 + l * njl + njl - 1] = -1.6 * sg + rng.normal(size=n) * noise
    s[:, 7:10] = rng.normal(size=(n, 3)) * 0.1
    s[:, 13 + nj:] = rng.normal(size=(n, nj)) * 0.5
    return s.astype(np.float32).astype(np.float64)


def euler_slots(d0, blocks):
    idx = []
    for b in range(blocks):
        idx += [b * d0 + 1, b * d0 + 2, b * d0 + 3]
    return np.array(idx)


def flag_slots(d0, nj, blocks):
    """Indices of the four foot-contact flags of every D0 block of an observation (solo.py:219-220)."""
    return np.array([b * d0 + 10 + 2 * nj + k for b in range(blocks) for k in range(4)])


def obs_diff(a, b, d0):
    """|a - b| with the three Euler slots of every D0 block compared modulo 1: the reference's
    observation maps an Euler angle e to (e mod 2)/2 (solo.py:206), which jumps by 1
    at e = 0, so two implementations that agree to 1e-9 on e can differ by ~1 in that slot."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    blocks = a.shape[-1] // d0
    sl = euler_slots(d0, blocks)
    w = d[..., sl]
    w = np.minimum(w, np.abs(1.0 - w))
    w = np.minimum(w, np.abs(2.0 - d[..., sl]))
    d[..., sl] = w
    return d


def oracle_policy_episodes(policy, cfg, num_envs, seed=0, torch_seed=0, deterministic=False, nthreads=None):
    """First episode of each of `num_envs` ORACLE envs (fp64 CPU restatement) driven by a torch policy on
    the CPU: returns (episode_return, episode_length, last_reward) arrays.  The CPU counterpart of
    solorl_b200.agents.evaluate.evaluate for the episode-return parity test."""
    import torch
    from oracle.oracle import OracleVecEnv
    from solorl_b200.abi import params_from_config
    from solorl_b200.model import SoloModel
    m = SoloModel.resolve(cfg["model_urdf"])
    p = params_from_config(cfg, m)
    v = OracleVecEnv(m, p, num_envs, seed=seed, nthreads=nthreads)
    obs = torch.from_numpy(v.reset())
    g = torch.Generator().manual_seed(torch_seed)
    ret = np.zeros(num_envs); length = np.zeros(num_envs, dtype=np.int64); last = np.zeros(num_envs)
    acc = np.zeros(num_envs); steps = np.zeros(num_envs, dtype=np.int64)
    finished = np.zeros(num_envs, dtype=bool)
    for t in range(int(cfg["episode_length"]) + 1):
        with torch.no_grad():
            value, feat = policy.base(obs)
            mean, logstd = policy.pi_dist(feat)
            a = mean if deterministic else mean + torch.randn(mean.shape, generator=g) * logstd.exp()
        o, r, d = v.step(a.numpy())
        acc += r; steps += 1
        newly = (d > 0.5) & ~finished
        ret[newly] = acc[newly]; length[newly] = steps[newly]; last[newly] = r[newly]
        finished |= newly
        if finished.all():
            break
        obs = torch.from_numpy(o)
    assert finished.all()
    return ret, length, last


def limit_states(rng, n, nj, airborne=False):
    """States whose joints sit at / beyond the +-10 rad URDF limits (one violated joint in 1..4 legs, never two
    in the same leg, shallow and deep violations, moving in and out), on the ground in a bent stance or in
    free flight: the inputs of the joint-limit parity tests."""
    s = stance_states(rng, n, nj, z=0.24, noise=0.1)
    if airborne:
        s[:, 2] += 0.8
        s[:, 10:13] = rng.normal(size=(n, 3))
    njl = nj // 4
    for i in range(n):
        legs = rng.choice(4, size=rng.integers(1, 5), replace=False)
        for l in legs:
            k = rng.integers(0, njl)
            j = l * njl + k
            sign = rng.choice([-1.0, 1.0])
            depth = rng.choice([0.0, 0.005, 0.03, 0.2]) if rng.random() < 0.8 else 0.05 * rng.random()
            s[i, 13 + j] = sign * (10.0 + depth)
            s[i, 13 + nj + j] = rng.normal() * 6.0
    return s.astype(np.float32).astype(np.float64)


def collapsed_states(rng, n, robot, upside_down=0.15, params=None):
    """States of a robot that has fallen, harvested from ORACLE trajectories with body contacts on: dropped from
    0.2-0.3 m with a random tilt (a fraction on its back) and random leg posture under small random torques, then
    60-300 substeps later.  Knees and base-box corners rest on (or are about to hit) the ground in most samples: the
    inputs of the body-contact (SoloSimParams.body_contacts) parity tests."""
    from oracle.oracle import OracleEnv, default_params
    from solorl_b200.model import SoloModel
    m = SoloModel.builtin(robot)
    p = params
    if p is None:
        p = default_params()
        p.body_contacts = 1
        if robot == "solo8":
            p.base_half_x, p.base_half_y = 0.212, 0.1046
    e = OracleEnv(m, p)
    nj = e.nj
    out = np.zeros((n, 13 + 2 * nj))
    for i in range(n):
        s = np.zeros(13 + 2 * nj)
        s[2] = rng.uniform(0.2, 0.3)
        roll = rng.normal() * 0.5 + (np.pi if rng.random() < upside_down else 0.0)
        pitch, yaw = rng.normal() * 0.4, rng.uniform(-np.pi, np.pi)
        cr, sr, cp, sp, cy, sy = np.cos(roll / 2), np.sin(roll / 2), np.cos(pitch / 2), np.sin(pitch / 2), np.cos(yaw / 2), np.sin(yaw / 2)
        s[3:7] = [sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy]
        s[13:13 + nj] = rng.uniform(-2.0, 2.0, size=nj)
        e.set_state(s)
        for _ in range(int(rng.integers(60, 300))):
            e.substep(rng.normal(size=nj) * 0.3)
        out[i] = e.get_state()
    return out.astype(np.float32).astype(np.float64)
