"""Kernel math (solo_core.cuh / solo_env.cuh replayed on the CPU in fp32 by tests/emu) against
the fp64 oracle, at the tolerances BASELINE.json:north_star states:
  contact-free joint accelerations 1e-5 relative; PD torque / reward / observation 1e-6 given
  identical state; single-step state with foot contact 1e-3.
The same comparisons run against the real CUDA path in test_gpu_parity.py (-m gpu)."""
import numpy as np
import pytest

from oracle.oracle import OracleEnv, default_params
from solorl_b200.abi import params_from_config
from solorl_b200.model import SoloModel
from tests.emu.emu import EmuEnv
from tests.helpers import make_config, obs_diff, random_states, stance_states

ROBOTS = ("solo8", "solo12")
TOL_QDD = 1e-5      # north_star: contact-free joint-space accelerations, relative
TOL_ENV = 1e-6      # north_star: PD torques, rewards, observations given identical state
TOL_CONTACT = 1e-3  # north_star: single-step state with foot contact


def pair(robot, **kw):
    cfg = make_config(robot, **kw)
    m = SoloModel.resolve(robot)
    p = params_from_config(cfg, m)
    return OracleEnv(m, p, seed=5, env_id=2), EmuEnv(m, p, seed=5, env_id=2), p, m


@pytest.mark.parametrize("robot", ROBOTS)
def test_forward_dynamics_1e5(robot):
    rng = np.random.default_rng(10)
    o, e, _, _ = pair(robot)
    worst = 0.0
    for s in random_states(rng, 300, o.nj):
        tau = rng.uniform(-3, 3, size=o.nj).astype(np.float32).astype(np.float64)
        o.set_state(s)
        e.set_state(s)
        ref, got = o.forward_dynamics(tau), e.forward_dynamics(tau).astype(np.float64)
        worst = max(worst, np.linalg.norm(ref[6:] - got[6:]) / np.linalg.norm(ref[6:]),
                    np.linalg.norm(ref - got) / np.linalg.norm(ref))
    assert worst < TOL_QDD, worst


@pytest.mark.parametrize("control", ["torque", "pd", "vpd"])
def test_action_to_torque_1e6(control):
    rng = np.random.default_rng(11)
    o, e, _, _ = pair("solo12", control=control)
    for s in random_states(rng, 50, o.nj, vel_scale=0.2):
        a = rng.uniform(-1.5, 1.5, size=o.act_dim).astype(np.float32).astype(np.float64)
        if control == "vpd":
            a[-2:] = [rng.uniform(0, 6), rng.uniform(0, 0.3)]
            a = a.astype(np.float32).astype(np.float64)
        o.set_state(s)
        e.set_state(s)
        assert np.abs(o.action_to_torque(a) - e.action_to_torque(a)).max() < TOL_ENV * 3.0   # |tau| <= 3


@pytest.mark.parametrize("robot,task,H", [("solo8", "stand", 0), ("solo12", "walk", 1), ("solo12", "pointgoal", 2)])
def test_observation_given_identical_state_1e6(robot, task, H):
    rng = np.random.default_rng(12)
    o, e, _, _ = pair(robot, task=task, H=H)
    for s in random_states(rng, 50, o.nj, vel_scale=0.3):
        if task == "pointgoal":      # goal first: set_state fills the history with the current goal
            o.set_goal(1.5, -1.25)
            e.set_goal(1.5, -1.25)
        o.set_state(s)
        e.set_state(s)
        a, b = o.get_observation(), e.get_observation()
        scale = np.maximum(1.0, np.abs(a))
        assert (obs_diff(a, b, o.d0) / scale).max() < TOL_ENV


@pytest.mark.parametrize("task,control", [("stand", "torque"), ("walk", "torque"), ("pointgoal", "torque"), ("stand", "pd")])
def test_reward_given_identical_state_1e6(task, control):
    """Physics frozen (no gravity, no damping, at rest in the air) so the post-step state IS the
    injected one: isolates get_reward / is_episode_finished arithmetic."""
    rng = np.random.default_rng(13)
    cfg = make_config("solo12", task=task, control=control, H=1)
    m = SoloModel.resolve("solo12")
    p = params_from_config(cfg, m)
    p.gravity_z = 0.0
    p.lin_damping = p.ang_damping = 0.0
    if control == "pd":
        p.kp = p.kd = 0.0
    o, e = OracleEnv(m, p), EmuEnv(m, p)
    for i in range(40):
        s = np.zeros(13 + 24)
        if i % 4:                                   # well clear of the ground: no contact rows
            s[2] = rng.uniform(0.6, 1.5)
            ang = rng.normal(size=3) * 0.3
            s[3:7] = [np.sin(ang[0] / 2), 0, 0, np.cos(ang[0] / 2)]
            s[13:25] = rng.uniform(-1, 1, size=12)
        else:                                       # below stand_z (0.2): legs folded forward
            s[2], s[6] = rng.uniform(0.1, 0.19), 1.0
            s[13:25] = rng.uniform(-0.1, 0.1, size=12)
            s[14:25:3] += np.pi / 2
        s = s.astype(np.float32).astype(np.float64)
        a = np.zeros(12) if control == "torque" else rng.uniform(-1, 1, size=12)
        if control == "torque":
            # zero torque keeps the state frozen; the raw-action penalty is tested through vx instead
            s[7] = np.float32(rng.normal())
        a = a.astype(np.float32).astype(np.float64)
        o.set_state(s)
        e.set_state(s)
        if task == "pointgoal":
            o.set_goal(3.0, 0.5)
            e.set_goal(3.0, 0.5)
        _, ro, do, io = o.step(a)
        _, re, de, ie = e.step(a)
        assert do == de
        # pointgoal: progress = (potential_old - potential_new) * 60 (baseEnv.py:137): one fp32 ulp of
        # the stored base position (1.2e-7 at |x| ~ 1-2 m) is already 7e-6 of reward, so the moving
        # pointgoal case is bounded by the state representation, not by the reward arithmetic
        tol = 3e-5 if task == "pointgoal" else TOL_ENV
        assert abs(ro - re) < tol * max(1.0, abs(ro)), (ro, re)
        for k in ("dr_stand", "dr_joint_pose", "dr_torque", "dr_balance", "dr_progress"):
            assert abs(io[k] - ie[k]) < (tol if k == "dr_progress" else TOL_ENV) * max(1.0, abs(io[k]))


def test_reward_torque_penalty_uses_raw_action():
    o, e, _, _ = pair("solo8", task="stand", H=0)
    s = np.zeros(13 + 16)
    s[2], s[6] = 2.0, 1.0
    a = np.array([1.7, -2.2, 0.3, 0, 0, 0, 0, 0.9], dtype=np.float32).astype(np.float64)
    o.set_state(s)
    e.set_state(s)
    _, ro, _, io = o.step(a)
    _, re, _, ie = e.step(a)
    assert io["dr_torque"] == pytest.approx(-0.01 * np.sum(a ** 2), abs=1e-12)
    assert abs(io["dr_torque"] - ie["dr_torque"]) < TOL_ENV and abs(ro - re) < 1e-5


def _hold_torque(s, nj, target, kp=3.0, kd=0.05):
    q, qd = s[13:13 + nj], s[13 + nj:]
    return np.clip(kp * (target - q) - kd * qd, -3, 3)


@pytest.mark.parametrize("robot", ROBOTS)
def test_contact_substep_1e3(robot):
    """Standing on bent legs under a joint-space PD hold plus random torque noise: 3-4 feet in
    contact, sticking and sliding friction.  Identical (fp32-representable) state injected in both
    before every substep."""
    rng = np.random.default_rng(14)
    o, e, _, m = pair(robot)
    nj = o.nj
    errs, ncs = [], []
    for s0 in stance_states(rng, 12, nj):
        target = s0[13:13 + nj].copy()
        o.set_state(s0)
        for t in range(60):
            s = o.get_state().astype(np.float32).astype(np.float64)
            tau = (_hold_torque(s, nj, target) + rng.normal(size=nj) * 0.3).astype(np.float32).astype(np.float64)
            o.set_state(s)
            e.set_state(s)
            o.substep(tau)
            e.substep(tau)
            so, se = o.get_state(), e.get_state().astype(np.float64)
            errs.append((np.abs(so - se) / np.maximum(1.0, np.abs(so))).max())
            co, ce = o.get_contacts(), e.get_contacts()
            assert (co[:, 1] == ce[:, 1]).all()
            f_scale = max(1.0, co[:, 2].max())
            assert np.abs(co[:, 2] - ce[:, 2]).max() < 2e-2 * f_scale
            ncs.append(co[:, 1].sum())
    assert np.mean(ncs) > 2.5, "test must exercise multi-foot contact"
    assert max(errs) < TOL_CONTACT, (max(errs), np.median(errs))


@pytest.mark.parametrize("robot", ROBOTS)
def test_contact_substep_singular_reset_pose(robot):
    """The reference resets into the straight-leg (kinematically singular) pose and drops the robot
    onto four feet: a statically indeterminate, ill-conditioned contact problem.  There the fp64
    oracle ITSELF moves by up to ~1e-3 when its input state is perturbed by one fp32 ulp, so
    the fixed 1e-3 tolerance is widened by that measured sensitivity (and only there)."""
    rng = np.random.default_rng(15)
    o, e, p, m = pair(robot)
    o2 = OracleEnv(m, p)
    nj = o.nj
    errs, sens = [], []
    for ep in range(6):
        o.reset()
        for t in range(40):
            s = o.get_state().astype(np.float32).astype(np.float64)
            tau = (rng.uniform(-3, 3, size=nj) * (t % 4 == 0)).astype(np.float32).astype(np.float64)
            s2 = s * (1 + rng.choice([-1, 1], size=s.shape) * 6e-8)
            o.set_state(s)
            e.set_state(s)
            o2.set_state(s2)
            o.substep(tau)
            e.substep(tau)
            o2.substep(tau)
            so, se, sp = o.get_state(), e.get_state().astype(np.float64), o2.get_state()
            scale = np.maximum(1.0, np.abs(so))
            errs.append((np.abs(so - se) / scale).max())
            sens.append((np.abs(so - sp) / scale).max())
    errs, sens = np.array(errs), np.array(sens)
    assert np.median(errs) < 1e-5
    assert np.quantile(errs, 0.9) < TOL_CONTACT
    assert np.all(errs < TOL_CONTACT + 50 * np.maximum(sens, np.quantile(sens, 0.99)))


@pytest.mark.parametrize("robot,task,control,H", [("solo8", "walk", "torque", 1), ("solo12", "pointgoal", "torque", 1),
                                                   ("solo8", "stand", "pd", 0), ("solo12", "walk", "torque", 2)])
def test_env_rollout_with_reset(robot, task, control, H):
    """reset() (settle count and goal from the shared Philox stream) then a short rollout with
    auto-reset: observations (Euler slots compared modulo 1), rewards, done flags and
    episode records.  No state re-injection here, so the solver runs its fixed iteration count
    (solver_residual_threshold 0): with Bullet's early exit fp32 and fp64 can stop one sweep apart,
    which is a discontinuity that multi-step trajectories amplify (checked separately below)."""
    rng = np.random.default_rng(16)
    cfg = make_config(robot, task=task, control=control, H=H, episode_length=12, solver_residual_threshold=0.0)
    m = SoloModel.resolve(robot)
    p = params_from_config(cfg, m)
    for env_id in range(3):
        o, e = OracleEnv(m, p, seed=3, env_id=env_id), EmuEnv(m, p, seed=3, env_id=env_id)
        a0, b0 = o.reset(), e.reset()
        assert o.settle_count_last == e.settle_count_last
        assert obs_diff(a0, b0, o.d0).max() < 5e-4   # 20-44 substeps through the singular straight-leg drop
        for t in range(30):
            a = rng.uniform(-1.2, 1.2, size=o.act_dim).astype(np.float32).astype(np.float64)
            oo, ro, do, io = o.step(a, auto_reset=True)
            eo, re, de, ie = e.step(a, auto_reset=True)
            assert do == de
            assert obs_diff(oo, eo, o.d0).max() < 2e-2, (t, obs_diff(oo, eo, o.d0).max())   # free-running trajectories: glue logic, not numerics
            assert abs(ro - re) < 2e-2 * max(1.0, abs(ro))
            if do:
                assert io["episode_length"] == ie["episode_length"] and io["success"] == ie["success"]
                assert io["timeout"] == ie["timeout"] and io["goals_reached"] == ie["goals_reached"]
                assert abs(io["episode_return"] - ie["episode_return"]) < 2e-2 * max(1.0, abs(io["episode_return"]))


def test_env_rollout_default_threshold_statistics():
    """Same rollout with PyBullet's default residual threshold (1e-7): typical agreement stays at the
    1e-4 level; isolated steps where the two precisions stop one sweep apart may drift."""
    rng = np.random.default_rng(17)
    cfg = make_config("solo12", task="walk", H=1, episode_length=12)
    m = SoloModel.resolve("solo12")
    p = params_from_config(cfg, m)
    assert p.solver_residual_threshold == 1e-7
    errs = []
    for env_id in range(4):
        o, e = OracleEnv(m, p, seed=3, env_id=env_id), EmuEnv(m, p, seed=3, env_id=env_id)
        o.reset()
        e.reset()
        for t in range(24):
            a = rng.uniform(-1.2, 1.2, size=12).astype(np.float32).astype(np.float64)
            oo, ro, do, _ = o.step(a, auto_reset=True)
            eo, re, de, _ = e.step(a, auto_reset=True)
            if do != de:
                break
            errs.append(obs_diff(oo, eo, o.d0).max())
    assert len(errs) > 60 and np.median(errs) < 1e-3


def test_solver_early_exit_matches_fixed_count_when_converged():
    """The residual threshold only ends the sweep loop early; on a well-conditioned stance the
    result differs from the 50-sweep result by less than the threshold's velocity scale."""
    rng = np.random.default_rng(18)
    m = SoloModel.resolve("solo12")
    pa = params_from_config(make_config("solo12"), m)
    pb = params_from_config(make_config("solo12", solver_residual_threshold=0.0), m)
    ea, eb = EmuEnv(m, pa), EmuEnv(m, pb)
    oa = OracleEnv(m, pa)
    for s in stance_states(rng, 10, 12):
        ea.set_state(s)
        eb.set_state(s)
        oa.set_state(s)
        tau = np.zeros(12, np.float32)
        ea.substep(tau)
        eb.substep(tau)
        oa.substep(tau.astype(np.float64))
        assert oa.last_solver_iters <= 50
        assert np.abs(ea.get_state() - eb.get_state()).max() < 5e-3


def test_step_before_reset_is_an_error():
    _, e, _, _ = pair("solo8")
    with pytest.raises(AssertionError):
        e.step(np.zeros(8, np.float32))


@pytest.mark.parametrize("airborne", [False, True])
@pytest.mark.parametrize("robot", ROBOTS)
def test_joint_limit_substep_1e3(robot, airborne):
    """Joint-limit rows ([3P] btMultiBodyJointLimitConstraint): one substep from states with joints at or
    beyond +-10 rad, with and without foot contacts, lane program against the oracle."""
    from tests.helpers import limit_states
    rng = np.random.default_rng(41)
    o, e, _, m = pair(robot)
    nj = o.nj
    errs, rows, ncs = [], [], []
    for s in limit_states(rng, 80, nj, airborne=airborne):
        tau = rng.uniform(-3, 3, size=nj).astype(np.float32).astype(np.float64)
        o.set_state(s); e.set_state(s)
        o.substep(tau); e.substep(tau)
        so, se = o.get_state(), e.get_state().astype(np.float64)
        errs.append((np.abs(so - se) / np.maximum(1.0, np.abs(so))).max())
        rows.append(o.last_limit_rows)
        ncs.append(o.get_contacts()[:, 1].sum())
        assert (o.get_contacts()[:, 1] == e.get_contacts()[:, 1]).all()
    assert np.mean(rows) > 1.5 and (airborne or np.mean(ncs) > 0.8)      # a leg parked at 10 rad points away from the ground
    assert max(errs) < TOL_CONTACT, (max(errs), np.median(errs))


def test_joint_limits_bound_the_joint_range_in_rollouts():
    """Random-action rollout of the lane program: without the rows a free leg spins to +-20 rad and beyond,
    with them |q| stays within a step's travel of the 10 rad limit."""
    rng = np.random.default_rng(42)
    worst = {}
    for jl in (0, 1):
        cfg = make_config("solo12", task="walk", H=1, episode_length=200, joint_limits=jl)
        m = SoloModel.resolve("solo12")
        e = EmuEnv(m, params_from_config(cfg, m), seed=3, env_id=1)
        e.reset()
        mx = 0.0
        r2 = np.random.default_rng(7)
        for t in range(300):
            obs, r, d, info = e.step(r2.uniform(-1, 1, size=12), auto_reset=True)
            mx = max(mx, np.abs(np.asarray(obs)[10:22]).max() * 10.0)
        worst[jl] = mx
    assert worst[0] > 12.0 and worst[1] < 10.6, worst


@pytest.mark.parametrize("robot", ROBOTS)
def test_body_contact_substep_1e3(robot):
    """Knees and base-box corners on the ground (SoloSimParams.body_contacts, solo_body.cuh): a fallen robot resting
    on 5-11 contact points, some with joint-limit rows.  Same single-substep bound as foot contact on every sample
    where the fp64 oracle itself is stable to a one-ulp (fp32) perturbation of its input; a robot lying on the
    ground is a redundant contact problem, and on a few per cent of the samples 50 unconverged Gauss-Seidel sweeps
    amplify that perturbation to 1e-2..1 in the oracle itself (see the GPU twin of this test)."""
    from tests.helpers import collapsed_states
    rng = np.random.default_rng(31)
    o, e, p, m = pair(robot, body_contacts=1)
    o2 = OracleEnv(m, p)
    assert p.body_contacts == 1
    nj = o.nj
    errs, sens, body = [], [], 0
    for s0 in collapsed_states(rng, 24, robot, params=p):
        o.set_state(s0)
        for t in range(25):
            s = o.get_state().astype(np.float32).astype(np.float64)
            tau = (rng.normal(size=nj) * 0.5).astype(np.float32).astype(np.float64)
            o.set_state(s)
            e.set_state(s)
            o2.set_state(s * (1 + rng.choice([-1, 1], size=s.shape) * 6e-8))
            body += int((o.contact_rows(tau)["kind"] == 0).sum() > 4)        # more points than feet
            o.substep(tau)
            e.substep(tau)
            o2.substep(tau)
            so, se, sp = o.get_state(), e.get_state().astype(np.float64), o2.get_state()
            scale = np.maximum(1.0, np.abs(so))
            errs.append((np.abs(so - se) / scale).max())
            sens.append((np.abs(so - sp) / scale).max())
            co, ce = o.get_contacts(), e.get_contacts()
            assert (co[:, 1] == ce[:, 1]).all()
    errs, sens = np.array(errs), np.array(sens)
    stable = sens < 1e-4
    print(f"[{robot}] {len(errs)} substeps, {body} with more than four contact points, {stable.mean():.3f} stable: "
          f"median {np.median(errs):.2e}, stable max {errs[stable].max():.2e}, overall max {errs.max():.2e}")
    assert body > 100 and stable.mean() > 0.9
    assert np.median(errs) < 1e-4 and errs[stable].max() < TOL_CONTACT
    assert np.all(errs < TOL_CONTACT + 50 * sens)
