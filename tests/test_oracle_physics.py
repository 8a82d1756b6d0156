"""Independent physics checks of the CPU oracle (they stand in for the missing
Bullet oracle — PyBullet is not installable here, physics parity is unpinned)."""
import numpy as np
import pytest

from tests.helpers import random_states, stance_states
from oracle.oracle import OracleEnv, default_params
from solorl_b200.model import SoloModel

ROBOTS = ("solo8", "solo12")


def params_for(robot, **kw):
    p = default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def model_and_params(robot, **kw):
    return SoloModel.builtin(robot), params_for(robot, **kw)


@pytest.mark.parametrize("robot", ROBOTS)
def test_aba_equals_crba_rnea(robot):
    """ABA in link-COM frames == mass-matrix solve in world-origin spatial coordinates."""
    rng = np.random.default_rng(0)
    m = SoloModel.builtin(robot)
    e = OracleEnv(m, default_params())
    for s in random_states(rng, 100, e.nj):
        s[:3] *= 0.05   # the CRBA variant takes moments about the WORLD origin: keep the base near it
        e.set_state(s)
        tau = rng.uniform(-3, 3, size=e.nj)
        a, b = e.forward_dynamics(tau), e.forward_dynamics_crba(tau)
        assert np.abs(a - b).max() <= 1e-9 * max(1.0, np.abs(a).max())


@pytest.mark.parametrize("robot", ROBOTS)
def test_free_fall(robot):
    """Damping off, at rest: base accelerates at g, joints do not move (uniform field)."""
    m = SoloModel.builtin(robot)
    p = default_params()
    p.lin_damping = 0.0
    p.ang_damping = 0.0
    e = OracleEnv(m, p)
    s = np.zeros(13 + 2 * e.nj)
    s[2], s[6] = 1.0, 1.0
    s[13:13 + e.nj] = np.linspace(-1, 1, e.nj)
    e.set_state(s)
    qdd = e.forward_dynamics(np.zeros(e.nj))
    assert np.allclose(qdd[:3], 0, atol=1e-12) and np.allclose(qdd[3:6], [0, 0, -9.81], atol=1e-12)
    assert np.abs(qdd[6:]).max() < 1e-10


@pytest.mark.parametrize("robot", ROBOTS)
def test_damping_opposes_motion(robot):
    """Bullet link damping 0.04(1+|v|) decelerates a translating body: a = g - 0.04(1+|v|)v."""
    m = SoloModel.builtin(robot)
    e = OracleEnv(m, default_params())
    s = np.zeros(13 + 2 * e.nj)
    s[2], s[6], s[7] = 1.0, 1.0, 1.0   # 1 m/s along x, no rotation: every link moves at v
    e.set_state(s)
    qdd = e.forward_dynamics(np.zeros(e.nj))
    assert qdd[3] == pytest.approx(-0.08, abs=1e-9)      # 0.08 m/s^2 at 1 m/s
    assert qdd[5] == pytest.approx(-9.81, abs=1e-9)


@pytest.mark.parametrize("robot", ROBOTS)
def test_energy_conserved_without_damping_or_contact(robot):
    rng = np.random.default_rng(1)
    m = SoloModel.builtin(robot)
    p = default_params()
    p.lin_damping = p.ang_damping = 0.0
    p.dt = 1e-4
    e = OracleEnv(m, p)
    s = random_states(rng, 1, e.nj, vel_scale=0.3)[0]
    s[2] = 5.0
    e.set_state(s)
    E0 = e.energy()
    for _ in range(1500):
        e.substep(np.zeros(e.nj))
    assert abs(e.energy() - E0) < 2e-4 * abs(E0)


@pytest.mark.parametrize("robot", ROBOTS)
def test_static_stance_supports_weight(robot):
    """After the reset drop the four normal forces add up to m g."""
    m = SoloModel.builtin(robot)
    e = OracleEnv(m, default_params(), seed=0, env_id=0)
    e.reset()
    c = e.get_contacts()
    assert (c[:, 1] == 1).all()
    assert c[:, 2].sum() == pytest.approx(m.total_mass * 9.81, rel=2e-3)
    feet = e.foot_positions()
    assert np.all(feet[:, 2] - m.foot_radius > -1e-3) and np.all(feet[:, 2] - m.foot_radius < 0.02)


@pytest.mark.parametrize("robot", ROBOTS)
def test_friction_cone_and_no_ground_penetration(robot):
    """Push a standing robot sideways: tangential force stays inside the cone; feet in contact
    do not sink."""
    rng = np.random.default_rng(2)
    m = SoloModel.builtin(robot)
    e = OracleEnv(m, default_params(), seed=0, env_id=1)
    e.reset()
    s = e.get_state()
    s[7] = 1.5   # sudden sideways base velocity
    e.set_state(s)
    for _ in range(10):
        v_before = e.get_state()[7]
        e.substep(np.zeros(e.nj))
        assert abs(e.get_state()[7]) <= abs(v_before) + 1e-9   # friction never speeds the base up
        feet = e.foot_positions()
        assert np.all(feet[:, 2] - m.foot_radius > -2e-3)


def test_contact_flag_semantics():
    """flag = 1 iff a contact point exists with normal force < 0.2 N."""
    m = SoloModel.builtin("solo8")
    e = OracleEnv(m, default_params(), seed=0, env_id=0)
    e.reset()
    c = e.get_contacts()
    assert (c[:, 2] > 0.2).all() and (c[:, 0] == 0).all()      # firmly loaded feet -> flag 0
    s = e.get_state()
    s[2] += 0.01     # lift 1 cm: inside the 0.02 margin, zero force -> point exists, flag 1
    e.set_state(s)
    e.substep(np.zeros(e.nj))
    c = e.get_contacts()
    assert (c[:, 1] == 1).all() and (c[:, 2] < 0.2).all() and (c[:, 0] == 1).all()
    s[2] += 0.5      # far above the ground: no contact point, flag 0
    e.set_state(s)
    e.substep(np.zeros(e.nj))
    assert (e.get_contacts()[:, :2] == 0).all()


def test_max_coordinate_velocity_clamp():
    m = SoloModel.builtin("solo8")
    e = OracleEnv(m, default_params())
    s = np.zeros(13 + 2 * e.nj)
    s[2], s[6] = 5.0, 1.0
    s[13 + e.nj] = 99.9
    e.set_state(s)
    e.substep(np.full(e.nj, 3.0))
    assert np.abs(e.get_state()[13 + e.nj:]).max() <= 100.0 + 1e-12   # [3P] m_maxCoordinateVelocity


# ---- joint-limit rows ([3P] btMultiBodyJointLimitConstraint, created by PyBullet's URDF importer) ---------------
def _free_flight_state(nj, q_over, qd_over, joint=4):
    s = np.zeros(13 + 2 * nj)
    s[2] = 1.0; s[6] = 1.0
    s[13:13 + nj] = np.linspace(-0.5, 0.5, nj)
    s[13 + nj:] = np.linspace(0.3, -0.3, nj)
    s[13 + joint] = q_over
    s[13 + nj + joint] = qd_over
    return s


@pytest.mark.parametrize("robot", ROBOTS)
def test_joint_limit_row_stops_the_joint(robot):
    m, p = model_and_params(robot)
    nj = m.nj
    dt = p.dt
    o = OracleEnv(m, p)
    # upper bound, shallow violation: target velocity = erp * |pen| / dt towards the range
    o.set_state(_free_flight_state(nj, 10.02, 5.0))
    o.substep(np.zeros(nj))
    assert o.last_limit_rows == 1
    qd = o.get_state()[13 + nj + 4]
    assert abs(qd - (-0.2 * 0.02 / dt)) < 1e-9
    # lower bound, deep violation (beyond the split-impulse threshold): velocity target only
    o.set_state(_free_flight_state(nj, -10.3, -5.0))
    o.substep(np.zeros(nj))
    assert o.last_limit_rows == 1 and abs(o.get_state()[13 + nj + 4]) < 1e-9
    # moving back into the range already faster than the target: the row stays slack (unilateral)
    o.set_state(_free_flight_state(nj, 10.02, -5.0))
    o.substep(np.zeros(nj))
    ref = OracleEnv(m, params_for(robot, joint_limits=0))
    ref.set_state(_free_flight_state(nj, 10.02, -5.0))
    ref.substep(np.zeros(nj))
    assert np.abs(o.get_state() - ref.get_state()).max() < 1e-12
    # inside the range: no row
    o.set_state(_free_flight_state(nj, 9.99, 5.0))
    o.substep(np.zeros(nj))
    assert o.last_limit_rows == 0


@pytest.mark.parametrize("robot", ROBOTS)
def test_joint_limit_impulse_is_a_pure_joint_impulse(robot):
    """The velocity change a limit row makes is M^-1 J^T lambda with J a unit vector on the joint: M dv is zero
    on the base and on every other joint (momentum is conserved, the reaction goes where the mass matrix says)."""
    m, p = model_and_params(robot)
    nj = m.nj
    rng = np.random.default_rng(5)
    for joint in (0, nj // 2, nj - 1):
        s = random_states(rng, 1, nj)[0]
        s[13 + joint] = 10.01
        s[13 + nj + joint] = 8.0
        a, b = OracleEnv(m, p), OracleEnv(m, params_for(robot, joint_limits=0))
        a.set_state(s); b.set_state(s)
        _, M = a.forward_dynamics_crba(np.zeros(nj), want_M=True)
        a.substep(np.zeros(nj)); b.substep(np.zeros(nj))
        assert a.last_limit_rows == 1
        sa, sb = a.get_state(), b.get_state()
        dw, dvl = sa[10:13] - sb[10:13], sa[7:10] - sb[7:10]
        # the mass matrix of forward_dynamics_crba is in spatial coordinates about the WORLD origin: v_O = v + p x w
        dv = np.concatenate([dw, dvl + np.cross(s[:3], dw), sa[13 + nj:] - sb[13 + nj:]])
        f = M @ dv
        lam = f[6 + joint]
        assert lam < 0 and abs(lam) > 1e-4                      # pushes the joint back (upper bound: negative)
        f[6 + joint] = 0
        assert np.abs(f).max() < 1e-9 * max(1.0, abs(lam)) + 1e-10


def test_kernel_mode_limit_rows_against_bullet_order():
    """limit_rows_per_leg = 1 (what the kernels solve: one row per leg, the rows of different legs relaxed as one
    simultaneous group) against limit_rows_per_leg = 0 (Bullet: every violated joint, strictly row by row):
    identical with a single limit row, equal to solver tolerance with one row in each of two legs (they couple
    only through the base), and a second violated joint of the same leg is the documented difference."""
    m, _ = model_and_params("solo12")
    rng = np.random.default_rng(6)
    s = stance_states(rng, 1, 12)[0]
    s[13 + 1] = 10.05; s[13 + 12 + 1] = 3.0          # FL_HFE over the upper bound
    a = OracleEnv(m, params_for("solo12", limit_rows_per_leg=1))
    b = OracleEnv(m, params_for("solo12", limit_rows_per_leg=0))
    a.set_state(s); b.set_state(s)
    a.substep(np.zeros(12)); b.substep(np.zeros(12))
    assert a.last_limit_rows == 1 and b.last_limit_rows == 1
    assert np.abs(a.get_state() - b.get_state()).max() == 0.0
    s[13 + 8] = -10.02; s[13 + 12 + 8] = -2.0        # HL_KFE under the lower bound (another leg)
    a.set_state(s); b.set_state(s)
    a.substep(np.zeros(12)); b.substep(np.zeros(12))
    assert a.last_limit_rows == 2 and b.last_limit_rows == 2
    assert np.abs(a.get_state() - b.get_state()).max() < 1e-3
    s[13 + 2] = 10.5; s[13 + 12 + 2] = 1.0           # a second joint of the FL leg, deeper: replaces FL_HFE
    a.set_state(s); b.set_state(s)
    a.substep(np.zeros(12)); b.substep(np.zeros(12))
    assert a.last_limit_rows == 2 and b.last_limit_rows == 3
