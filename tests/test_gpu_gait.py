"""GPU tests of the gait-env row: actuator tick against the oracle, feet
positions, and the SoloGaitEnvContact shell arithmetic."""
import numpy as np
import pytest
import torch

from tests.helpers import stance_states

pytestmark = pytest.mark.gpu


def _actuator(n, seed=0):
    from solorl_b200.gait import ActuatorSim
    return ActuatorSim(n, solo12=True, dt=0.002, device=0, seed=seed)


def test_actuator_tick_matches_oracle_1e3():
    """tau = clip(P (q_des - q) + D (v_des - v) + tau_ff, +-3) then one 0.002 s simulator tick:
    GPU vs (numpy torque law -> oracle substep); 1e-3 like the contact substep (north_star)."""
    from oracle.oracle import OracleEnv
    rng = np.random.default_rng(3)
    n, nj = 64, 12
    rob = _actuator(n)
    s0 = stance_states(rng, n, nj, z=0.24, noise=0.1)
    rob.sim.set_state(torch.from_numpy(s0.astype(np.float32)).cuda())
    cmd = np.zeros((n, 5, nj), np.float32)
    cmd[:, 0] = s0[:, 13:13 + nj] + rng.normal(size=(n, nj)) * 0.3
    cmd[:, 1] = rng.normal(size=(n, nj))
    cmd[:, 2] = rng.uniform(0, 6, size=(n, nj))
    cmd[:, 3] = rng.uniform(0, 0.3, size=(n, nj))
    cmd[:, 4] = rng.uniform(-1, 1, size=(n, nj))
    ticks = 5
    rob.sim.actuator_step(torch.from_numpy(cmd).cuda(), ticks)
    got = rob.sim.get_state().cpu().numpy()
    feet = rob.sim.get_feet().cpu().numpy()
    worst = worst_feet = 0.0
    contacts = 0
    for i in range(0, n, 4):
        o = OracleEnv(rob.model, rob.params)
        o.set_state(s0[i])
        for _ in range(ticks):
            s = o.get_state()
            q, qd = s[13:13 + nj], s[13 + nj:]
            c = cmd[i].astype(np.float64)
            tau = np.clip(c[2] * (c[0] - q) + c[3] * (c[1] - qd) + c[4], -3.0, 3.0)
            o.substep(tau)
        ref = o.get_state()
        contacts += int(o.get_contacts()[:, 1].sum())
        worst = max(worst, float((np.abs(ref - got[i]) / np.maximum(1.0, np.abs(ref))).max()))
        worst_feet = max(worst_feet, float(np.abs(o.foot_positions() - feet[i]).max()))
    assert contacts > 0
    assert worst < 1e-3, worst
    assert worst_feet < 1e-3, worst_feet


def test_actuator_tick_with_body_contacts_matches_oracle_1e3():
    """The actuator kernel with body_contacts on (knees and base corners on the ground, solo_body.cuh): a fallen robot
    under the joint PD + feed-forward law, one 0.002 s tick at a time from re-injected states; 1e-3 on the samples
    where the fp64 oracle is itself stable to a one-ulp perturbation of its input (see test_body_contact_substep_1e3)."""
    from oracle.oracle import OracleVecEnv
    from solorl_b200.gait import ActuatorSim
    from tests.helpers import collapsed_states
    rng = np.random.default_rng(5)
    n, nj = 64, 12
    rob = ActuatorSim(n, solo12=True, dt=0.002, device=0, seed=0, body_contacts=1)
    assert rob.params.body_contacts == 1
    ov, ov2 = OracleVecEnv(rob.model, rob.params, n), OracleVecEnv(rob.model, rob.params, n)
    cur = collapsed_states(rng, n, "solo12", params=rob.params)
    errs, sens = [], []
    for t in range(12):
        cmd = np.zeros((n, 5, nj), np.float32)
        cmd[:, 0] = cur[:, 13:13 + nj] + rng.normal(size=(n, nj)) * 0.3
        cmd[:, 2] = rng.uniform(0, 6, size=(n, nj))
        cmd[:, 3] = rng.uniform(0, 0.3, size=(n, nj))
        cmd[:, 4] = rng.uniform(-1, 1, size=(n, nj))
        rob.sim.set_state(torch.from_numpy(cur.astype(np.float32)).cuda())
        rob.sim.actuator_step(torch.from_numpy(cmd).cuda(), 1)
        got = rob.sim.get_state().cpu().numpy().astype(np.float64)
        c = cmd.astype(np.float64)
        q, qd = cur[:, 13:13 + nj], cur[:, 13 + nj:]
        tau = np.clip(c[:, 2] * (c[:, 0] - q) + c[:, 3] * (c[:, 1] - qd) + c[:, 4], -3.0, 3.0)
        ref, _, _ = ov.substep_from(cur, tau)
        pert, _, _ = ov2.substep_from(cur * (1 + rng.choice([-1, 1], size=cur.shape) * 6e-8), tau)
        scale = np.maximum(1.0, np.abs(ref))
        errs.append((np.abs(ref - got) / scale).max(axis=1))
        sens.append((np.abs(ref - pert) / scale).max(axis=1))
        cur = got
    errs, sens = np.concatenate(errs), np.concatenate(sens)
    stable = sens < 1e-4
    assert stable.mean() > 0.9 and np.median(errs) < 1e-4
    rate = float((errs[stable] > 1e-3).mean())
    assert rate <= 5e-3 and errs[stable].max() < 1e-2, (rate, errs[stable].max())


def test_feet_positions_match_oracle_1e5():
    from oracle.oracle import OracleEnv
    from tests.helpers import random_states
    rng = np.random.default_rng(4)
    n = 32
    rob = _actuator(n)
    s = random_states(rng, n, 12)
    rob.sim.set_state(torch.from_numpy(s.astype(np.float32)).cuda())
    feet = rob.sim.get_feet().cpu().numpy()
    o = OracleEnv(rob.model, rob.params)
    for i in range(n):
        o.set_state(s[i])
        assert np.abs(o.foot_positions() - feet[i]).max() < 1e-5


def test_gait_shell_observation_reward_and_termination():
    from solorl_b200.gait import GAIT_TABLE, SoloGaitVecEnv, Q_INIT
    cfg = {"solo12": True, "episode_length": 3, "vel_switch": 1000, "mode": "headless", "num_history_stack": 1,
           "flat_ground": True, "auto_vel_switch": True}
    n = 16
    env = SoloGaitVecEnv(cfg, n, seed=1)
    assert env.k_rl == 80 and env.observation_space.shape == (64,) and env.action_space.n == 9
    with pytest.raises(AssertionError):
        env.step(torch.zeros(n, dtype=torch.long))                        # baseControlEnv.py:135
    obs = env.reset()
    assert obs.shape == (n, 64)
    o = obs.cpu().numpy()
    assert np.allclose(o[:, 0], env.robot.z_init, atol=1e-6) and np.allclose(o[:, 1:4], 0, atol=1e-6)
    assert np.allclose(o[:, 10:22], np.array(Q_INIT), atol=1e-6) and np.allclose(o[:, 22:34], 0)
    assert np.allclose(o[:, 46:58], 0) and np.allclose(o[:, 58:64], 0)    # past gaits = -1 -> zeros; v_ref masked to 0
    feet_z = o[:, 34:46].reshape(n, 4, 3)[:, :, 2]
    assert np.allclose(feet_z.min(axis=1), env.robot.model.foot_radius, atol=1e-5)
    a = torch.tensor([0, 5, 7, 8] * 4)
    for t in range(3):
        obs, rew, done, infos = env.step(a)
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all() and (rew <= 1.0 + 1e-6).all()
        if t < 2:
            seq = obs[:, 46:58].reshape(n, 3, 4)
            assert torch.allclose(seq[:, 2].cpu(), torch.tensor(GAIT_TABLE)[a])    # newest pattern last
    assert done.sum().item() >= 1
    i = int(torch.nonzero(done)[0])
    info = infos[i]
    assert info["episode_length"] <= 3 and set(["success", "timeout", "dr/Energy_pen", "dr/body_velocity"]) <= set(info)
    # an env that timed out at exactly episode_length is a success (baseControlEnv.py:186)
    to = [k for k in range(n) if done[k] > 0.5 and infos[k]["timeout"]]
    assert all(infos[k]["success"] and infos[k]["episode_length"] == 3 for k in to)
    assert (env.timestep[done > 0.5] == 0).all()                          # auto-reset
    env.close()


def test_posture_controller_keeps_the_robot_up_and_power_model():
    from solorl_b200.gait import COULOMB_TAU, K_MOTOR, VISCOUS_B, SoloGaitVecEnv
    cfg = {"solo12": True, "episode_length": 50, "mode": "headless", "flat_ground": True, "auto_vel_switch": True}
    env = SoloGaitVecEnv(cfg, 8, seed=2)
    env.reset()
    for t in range(10):
        obs, rew, done, infos = env.step(torch.zeros(8, dtype=torch.long))          # "Static": all feet in stance
    assert (obs[:, 0] > 0.15).all() and done.sum().item() == 0
    assert (rew > 0.5).all()                                                        # small energy, zero velocity error
    qd = env.robot.v_mes.cpu().numpy().astype(np.float64)
    tau = env.robot.tau_ff.cpu().numpy().astype(np.float64)
    ref = (COULOMB_TAU * np.sign(qd) + VISCOUS_B * qd) * qd + K_MOTOR * tau ** 2     # baseControlEnv.py:436-445
    assert np.abs(env.get_joints_power().cpu().numpy() - ref).max() < 1e-6
    env.close()


def test_make_vec_envs_accepts_the_contact_env():
    from solorl_b200.envs import make_vec_envs
    from solorl_b200.gait import SoloGaitEnvContact
    import yaml, os
    from tests.helpers import ROOT
    cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "basic_contact.yaml")))
    envs = make_vec_envs(cfg, 4, SoloGaitEnvContact)
    obs = envs.reset()
    obs, rew, done, infos = envs.step(torch.tensor([1, 2, 3, 4]))
    assert obs.shape == (4, 64) and rew.shape == (4, 1) and done.shape == (4,)
    envs.close()


def test_graphed_tick_loop_matches_eager():
    """The 80-tick controller loop of one RL step replayed as one CUDA graph against the eager loop: same
    observations, rewards and done flags over several RL steps including an auto-reset."""
    from solorl_b200.gait import SoloGaitVecEnv
    cfg = {"solo12": True, "episode_length": 3, "mode": "headless", "flat_ground": True, "auto_vel_switch": True}
    envs = [SoloGaitVecEnv(cfg, 64, seed=3, cuda_graph=g) for g in (True, False)]
    o1, o2 = envs[0].reset(), envs[1].reset()
    assert torch.equal(o1, o2)
    gen = torch.Generator(device="cuda").manual_seed(5)
    for t in range(5):
        a = torch.randint(0, 9, (64,), device="cuda", generator=gen)
        r1, r2 = envs[0].step(a), envs[1].step(a)
        assert torch.allclose(r1[0], r2[0], atol=1e-6) and torch.allclose(r1[1], r2[1], atol=1e-6)
        assert torch.equal(r1[2], r2[2])
    assert envs[0]._graph is not None
    for e in envs:
        e.close()


def test_actuator_tick_with_external_force_matches_oracle_1e3():
    """solo_set_external_force: a base-frame push at the base origin through five actuator ticks, GPU vs oracle
    (whose ABA and CRBA derivations agree on the same force to 1e-14 in the CPU suite); in free flight the base
    picks up exactly F dt / m_total per tick along the pushed axis when nothing else acts."""
    from oracle.oracle import OracleEnv
    rng = np.random.default_rng(6)
    n, nj = 64, 12
    rob = _actuator(n)
    s0 = stance_states(rng, n, nj, z=0.24, noise=0.1)
    s0[1::2, 2] += 0.6                                           # every other env in free flight
    F = (rng.normal(size=(n, 3)) * 6.0).astype(np.float32)
    rob.sim.set_state(torch.from_numpy(s0.astype(np.float32)).cuda())
    rob.sim.set_external_force(torch.from_numpy(F).cuda())
    cmd = np.zeros((n, 5, nj), np.float32)
    cmd[:, 0] = s0[:, 13:13 + nj]
    cmd[:, 2], cmd[:, 3] = 3.0, 0.2
    ticks = 5
    rob.sim.actuator_step(torch.from_numpy(cmd).cuda(), ticks)
    got = rob.sim.get_state().cpu().numpy()
    worst = 0.0
    for i in range(0, n, 3):
        o = OracleEnv(rob.model, rob.params)
        o.set_state(s0[i])
        o.set_external_force(F[i].astype(np.float64))
        for _ in range(ticks):
            s = o.get_state()
            c = cmd[i].astype(np.float64)
            tau = np.clip(c[2] * (c[0] - s[13:13 + nj]) + c[3] * (c[1] - s[13 + nj:]), -3.0, 3.0)
            o.substep(tau)
        ref = o.get_state()
        worst = max(worst, float((np.abs(ref - got[i]) / np.maximum(1.0, np.abs(ref))).max()))
    assert worst < 1e-3, worst
    # the force is not left on for solo_substep / a fresh handle: a second handle without the call falls freely
    rob2 = _actuator(n)
    rob2.sim.set_state(torch.from_numpy(s0.astype(np.float32)).cuda())
    rob2.sim.actuator_step(torch.from_numpy(cmd).cuda(), ticks)
    diff = (rob2.sim.get_state() - rob.sim.get_state())[1::2, 7:10].abs().max().item()
    assert diff > 1e-3                                           # the push did move the airborne bases
    rob.Stop(); rob2.Stop()


def test_gait_env_pushes_and_curriculum():
    """add_external_force + use_curriculum (baseControlEnv.py:117-123,276-289,321-328): per-episode pushes drawn in
    [min, max] N on one base axis, applied as a smooth bump over their window, and the curriculum widens both the
    velocity range and the force range."""
    import yaml, os
    from tests.helpers import ROOT
    from solorl_b200.gait import SoloGaitVecEnv, MAXFORCE
    cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "basic_contact.yaml")))
    cfg.update(add_external_force=True, use_curriculum=True, episode_length=20)
    env = SoloGaitVecEnv(cfg, 64, seed=3)
    assert env.min_max_force == [0, 2] and env.max_velocity == 0.0
    env.reset()
    assert (env.push_F.abs().sum(1) <= 2.0 + 1e-6).all() and ((env.push_F != 0).sum(1) <= 1).all()
    assert (env.push_start >= 500).all() and (env.push_start <= int(80 * 20 * 2 / 3)).all()
    for _ in range(9):
        env.increment_curriculum()
    assert env.min_max_force == [min(9, MAXFORCE - 2), MAXFORCE] and abs(env.max_velocity - 0.8) < 1e-6
    env.reset()
    mag = env.push_F.abs().sum(1)
    assert (mag >= 8 - 1e-6).all() and (mag <= 10 + 1e-6).all()
    # put every push window at the very start of the episode and compare with an env that is never pushed
    env.push_start.fill_(0); env.push_dur.fill_(1000)
    calm = SoloGaitVecEnv(dict(cfg, add_external_force=False), 64, seed=3)
    calm.reset()
    moved = 0.0
    for t in range(6):
        o1, r1, d1, _ = env.step(torch.zeros(64, dtype=torch.long))
        o2, r2, d2, _ = calm.step(torch.zeros(64, dtype=torch.long))
        assert torch.isfinite(o1).all()
    moved = (env.robot.sim.get_state()[:, :3] - calm.robot.sim.get_state()[:, :3]).abs().max().item()
    assert moved > 5e-3, moved                                   # a 8..10 N bump on a 2.5 kg robot shows within 0.96 s
    assert (env.k_tick == 6 * 80).all()
    # reset_vel_ref (baseControlEnv.py:236-238; `reset_vel` in testing/test_ppo.py:108): the imposed reference
    # velocity is what the next observation carries in its last six entries
    calm.reset_vel(np.array([[0.3, 0.0, 0.0, 0.0, 0.0, 0.1]]))
    assert torch.allclose(calm.get_observation()[:, -6:], torch.tensor([0.3, 0, 0, 0, 0, 0.1], device="cuda").expand(64, 6))
    env.close(); calm.close()


def test_ppo_on_the_gait_shell_improves_reward():
    """Config 4 is trainable: Categorical head over Discrete(9) (policy.py:22-23), `--env-name contact`, PPO on
    4096 gait-shell envs for 20 updates; the mean step reward of the later updates beats the first ones (with the
    stand-in posture controller and a zero velocity reference the static pattern is the cheapest, and PPO finds it)."""
    import yaml, os
    from tests.helpers import ROOT
    from solorl_b200.agents.train import default_args, train
    from solorl_b200.gait import SoloGaitEnvContact
    cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "basic_contact.yaml")))
    args = default_args(num_agents=4096, num_steps=8, mini_batch_size=8192, ppo_epoch=4, lr=3e-3, use_gae=True,
                        entropy_coef=0.0, num_env_steps=4096 * 8 * 20, log_interval=1, seed=5)
    out = train(args, cfg, SoloGaitEnvContact)
    hist = out["history"]
    assert len(hist) == 20 and out["actor_critic"].discrete
    first = np.mean([h["episode_reward"] for h in hist[:4] if h["episodes"] > 0] or [0.0])
    rew = [h.get("mean_step_reward") for h in hist]
    assert all(r is not None and np.isfinite(r) for r in rew)
    assert np.mean(rew[-5:]) > np.mean(rew[:5]) + 0.01, (rew[:5], rew[-5:])
