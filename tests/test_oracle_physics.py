"""Independent physics checks of the CPU oracle (they stand in for the missing
Bullet oracle — PyBullet is not installable here, physics parity is unpinned)."""
import numpy as np
import pytest

from tests.helpers import random_states
from oracle.oracle import OracleEnv, default_params
from solorl_b200.model import SoloModel

ROBOTS = ("solo8", "solo12")


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
