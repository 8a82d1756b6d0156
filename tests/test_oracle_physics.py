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


# ---- second, independent check of the contact pipeline -------------------------------
def _base_origin_mass_matrix(e, s, tau):
    """Dense joint-space mass matrix in the coordinates the constraint rows use: (w world, v of the base
    origin world, qd).  oracle_forward_dynamics_crba builds M for (w, v of the body-fixed point at the WORLD
    origin, qd); v_base = v_O + w x p, i.e. x_base = T x_O with T = [[1,0,0],[-[p]x,1,0],[0,0,1]]."""
    _, M = e.forward_dynamics_crba(tau, want_M=True)
    nd = M.shape[0]
    p = s[:3]
    px = np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])
    T = np.eye(nd)
    T[3:6, 0:3] = -px
    Ti = np.linalg.inv(T)
    return Ti.T @ M @ Ti


def _reference_pgs(A, target, kind, owner, mu, max_impulse, sweeps):
    """Projected Gauss-Seidel in Bullet's order on the dense Delassus matrix: limit rows, normals, then each
    contact's friction pair projected onto the cone (same projections as oracle_substep), no early exit."""
    n = len(target)
    lam = np.zeros(n)
    d = np.diag(A).copy()
    normal_of = {owner[r]: r for r in range(n) if kind[r] == 0}
    fa = [r for r in range(n) if kind[r] == 1]
    for _ in range(sweeps):
        for r in range(n):
            if kind[r] == 3:
                lam[r] = min(max(lam[r] + (target[r] - A[r] @ lam) / d[r], 0.0), max_impulse)
        for r in range(n):
            if kind[r] == 0:
                lam[r] = max(lam[r] + (target[r] - A[r] @ lam) / d[r], 0.0)
        for a in fa:
            b = a + 1
            assert kind[b] == 2 and owner[b] == owner[a]
            lim = mu * lam[normal_of[owner[a]]]
            sa = lam[a] + (target[a] - A[a] @ lam) / d[a]
            sb = lam[b] + (target[b] - A[b] @ lam) / d[b]
            nrm = np.hypot(sa, sb)
            sc = min(1.0, lim / nrm) if nrm > 0 else 1.0
            lam[a], lam[b] = sa * sc, sb * sc
    return lam


@pytest.mark.parametrize("robot", ROBOTS)
def test_delassus_rows_equal_dense_mass_matrix_solve(robot):
    """J M^-1 J^T from the dense CRBA mass matrix (world-origin spatial coordinates, Cholesky) against the
    rows the substep builds with one ABA impulse-response pass per row (link-COM frames): two independent
    derivations of M^-1 J^T and of the Delassus matrix, to 1e-10; contact rows and joint-limit rows."""
    from tests.helpers import limit_states
    rng = np.random.default_rng(7)
    m = SoloModel.builtin(robot)
    p = default_params()
    p.limit_rows_per_leg = 0
    e = OracleEnv(m, p)
    nj = e.nj
    states = np.concatenate([stance_states(rng, 12, nj), limit_states(rng, 12, nj)])
    states[:, :2] = rng.normal(size=(len(states), 2)) * 0.05
    nrows, nlim = 0, 0
    for s in states:
        e.set_state(s)
        tau = rng.uniform(-3, 3, size=nj)
        rows = e.contact_rows(tau)
        J, U = rows["J"], rows["U"]
        assert len(J) >= 3
        Mb = _base_origin_mass_matrix(e, s, tau)
        U_dense = np.linalg.solve(Mb, J.T).T
        assert np.abs(U - U_dense).max() <= 1e-10 * max(1.0, np.abs(U_dense).max())
        A_rows, A_dense = J @ U.T, J @ np.linalg.solve(Mb, J.T)
        assert np.abs(A_rows - A_dense).max() <= 1e-10 * np.abs(A_dense).max()
        assert np.abs(A_rows - A_rows.T).max() <= 1e-10 * np.abs(A_dense).max()      # symmetric
        assert np.linalg.eigvalsh(0.5 * (A_rows + A_rows.T)).min() > -1e-9           # positive semi-definite
        nrows += len(J)
        nlim += int((rows["kind"] == 3).sum())
    assert nlim >= 12 and nrows >= 24 * 7


@pytest.mark.parametrize("robot", ROBOTS)
def test_pgs_at_50_sweeps_against_a_converged_solve(robot):
    """The substep's solve against the same projected Gauss-Seidel run in numpy on the dense Delassus matrix
    assembled from the CRBA mass matrix (nothing shared with the substep but the Jacobian rows):
    (i) with the sweep count the substep used, the post-solve velocities agree to 1e-10 -- an independent
        derivation of M^-1 J^T, of the sweep order and of the cone projection;
    (ii) against 10^4 sweeps: where Bullet's residual exit fired (< 50 sweeps) the contact-space velocity is
        within the residual it allows (sqrt(1e-7) ~ 3e-4 m/s per row, a few rows deep); substeps that hit the
        50-sweep cap are NOT converged (up to 0.1 rad/s on a joint here) -- that is the reference engine's
        behaviour, which the oracle and the kernels reproduce rather than improve on;
    (iii) the converged impulses satisfy non-penetration / complementarity / the friction cone."""
    rng = np.random.default_rng(8)
    m = SoloModel.builtin(robot)
    p = default_params()
    e = OracleEnv(m, p)
    nj = e.nj
    same, exit_gap, capped = 0.0, 0.0, 0
    for s in stance_states(rng, 16, nj):
        e.set_state(s)
        tau = np.clip(-0.05 * s[13 + nj:] + rng.normal(size=nj) * 0.3, -3, 3)     # joint damping hold + noise
        rows = e.contact_rows(tau)
        J, target, kind, owner = rows["J"], rows["target"], rows["kind"], rows["owner"]
        Mb = _base_origin_mass_matrix(e, s, tau)
        MinvJt = np.linalg.solve(Mb, J.T)
        A = J @ MinvJt
        lam = _reference_pgs(A, target, kind, owner, p.friction, p.joint_limit_max_impulse, 10000)
        v_conv = rows["vstar"] + MinvJt @ lam
        for r in range(len(lam)):
            if kind[r] == 0:
                resid = A[r] @ lam - target[r]        # post-solve normal velocity minus its target
                assert lam[r] >= 0 and resid >= -1e-8 and abs(lam[r] * resid) < 1e-8
        for a in [r for r in range(len(lam)) if kind[r] == 1]:
            n = [r for r in range(len(lam)) if kind[r] == 0 and owner[r] == owner[a]][0]
            assert np.hypot(lam[a], lam[a + 1]) <= p.friction * lam[n] + 1e-10
        e.substep(tau)
        after = e.get_state()
        v_sub = np.concatenate([after[10:13], after[7:10], after[13 + nj:]])
        k = e.last_solver_iters
        lam_k = _reference_pgs(A, target, kind, owner, p.friction, p.joint_limit_max_impulse, k)
        same = max(same, np.abs(v_sub - (rows["vstar"] + MinvJt @ lam_k)).max())
        if k < p.solver_iters:
            exit_gap = max(exit_gap, np.abs(J @ (v_sub - v_conv)).max())
        else:
            capped += 1
    assert same < 1e-10, same
    assert exit_gap < 1.5e-3, exit_gap
    assert capped <= 4


# ---- contacts of knees / base-box corners (SoloSimParams.body_contacts) ------------------
@pytest.mark.parametrize("robot", ROBOTS)
def test_body_contact_rows_equal_dense_mass_matrix_solve(robot):
    """The rows of knee and base-corner contacts (base-link impulse response, a second contact point on the lower
    leg) against J M^-1 J^T from the dense CRBA mass matrix, to 1e-10; then the substep's solve against the numpy
    PGS on that dense matrix with the sweep count the substep used (row order: limit rows, all normals, then the
    friction pair of each point; points ordered feet, knees, lower corners, upper corners)."""
    from tests.helpers import collapsed_states
    rng = np.random.default_rng(17)
    m = SoloModel.builtin(robot)
    p = default_params()
    p.body_contacts = 1
    p.limit_rows_per_leg = 0          # Bullet's sequential order for the limit rows, which _reference_pgs restates
    if robot == "solo8":
        p.base_half_x, p.base_half_y = 0.212, 0.1046
    e = OracleEnv(m, p)
    nj = e.nj
    body_rows, most, same = 0, 0, 0.0
    for s in collapsed_states(rng, 24, robot, params=p):
        e.set_state(s)
        tau = rng.uniform(-1, 1, size=nj)
        rows = e.contact_rows(tau)
        J, U, kind, owner, target = rows["J"], rows["U"], rows["kind"], rows["owner"], rows["target"]
        if len(J) == 0:
            continue
        Mb = _base_origin_mass_matrix(e, s, tau)
        MinvJt = np.linalg.solve(Mb, J.T)
        assert np.abs(U - MinvJt.T).max() <= 1e-10 * max(1.0, np.abs(MinvJt).max())
        A = J @ MinvJt
        assert np.abs(J @ U.T - A).max() <= 1e-10 * np.abs(A).max()
        nn = int((kind == 0).sum())
        most = max(most, nn)
        body_rows += 3 * max(0, nn - 4)
        e.substep(tau)
        after = e.get_state()
        v_sub = np.concatenate([after[10:13], after[7:10], after[13 + nj:]])
        lam_k = _reference_pgs(A, target, kind, owner, p.friction, p.joint_limit_max_impulse, e.last_solver_iters)
        v_ref = np.clip(rows["vstar"] + MinvJt @ lam_k, -p.max_coord_vel, p.max_coord_vel)
        same = max(same, np.abs(v_sub - v_ref).max())
    assert most >= 6 and body_rows >= 60, (most, body_rows)       # the sample must hold knee / corner contacts
    assert same < 1e-9, same


@pytest.mark.parametrize("robot", ROBOTS)
def test_body_contacts_hold_a_collapsed_robot(robot):
    """Dropped flat on its belly with the legs folded up, the robot comes to rest on the lower corners of its base
    box (base origin ~0.025 m above the ground, below the z < 0.05 termination height of baseEnv.py:169); with
    feet-only contacts (round 1) the same robot sinks through the floor.  Switching body contacts on changes nothing
    while only the feet are near the ground."""
    m = SoloModel.builtin(robot)
    nj = 8 if robot == "solo8" else 12
    njl = nj // 4

    def run(body, s0, steps):
        p = default_params()
        p.body_contacts = body
        e = OracleEnv(m, p)
        e.set_state(s0)
        for _ in range(steps):
            e.substep(np.zeros(nj))
        return e.get_state()

    s = np.zeros(13 + 2 * nj)
    s[2], s[6] = 0.06, 1.0
    for l in range(4):
        s[13 + l * njl + njl - 2] = np.pi / 2 * (1.0 if l < 2 else -1.0) + np.pi   # upper legs pointing up
    rest, sunk = run(1, s, 240), run(0, s, 240)
    assert 0.015 < rest[2] < 0.035 and np.abs(rest[7:13]).max() < 0.05, rest[:13]
    assert sunk[2] < -0.1
    rng = np.random.default_rng(3)
    st = stance_states(rng, 1, nj)[0]
    assert np.array_equal(run(1, st, 40), run(0, st, 40))
