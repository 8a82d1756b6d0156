#!/usr/bin/env python
"""Record golden fixtures from the UNMODIFIED reference (michel-aractingi/soloRL) running on PyBullet.

PyBullet, gym and pybullet_envs are not installable in the build image, so the physics oracle of
this repository is "parity unpinned".  This script is what pins it the day a machine with the reference's
dependencies is at hand:

    pip install pybullet gym==0.21 numpy torch pyyaml          # any box, no GPU needed
    git clone https://github.com/michel-aractingi/soloRL
    python tools/record_pybullet_fixtures.py --reference /path/to/soloRL --out tests/golden/pybullet

It drives the reference's own ``SoloBaseEnv`` (baseEnv.py:6-187) and, for the single-step records, the very
``p.setJointMotorControlArray(TORQUE_CONTROL)`` + ``p.stepSimulation()`` pair of solo.py:256-265, and writes one
``<robot>_<task>_<control>.npz`` per configuration with

  engine_json       p.getPhysicsEngineParameters() (every [3P-UNVERIFIED] constant of SURVEY Appendix B:
                    numSolverIterations, erp / contactERP / frictionERP, solverResidualThreshold, ...)
  dynamics_json     p.getDynamicsInfo(robot, link) for the base and every link (mass, local inertia diagonal,
                    lateral friction, restitution, damping, contact stiffness) and p.getJointInfo rows
  pybullet_version  p.getAPIVersion() and pybullet's package version
  step_*            env-level transitions of SoloBaseEnv.step: pre-state, action, post-state, observation (the
                    terminal one is None in the reference: recorded as NaN), reward, done, contact points
  sub_*             single p.stepSimulation() transitions with the applied joint torques: pre-state, tau,
                    post-state, contact points of every robot link against the ground after the step

State rows use this repository's layout (include/solo_b200.h): pos(3) quat xyzw(4) linvel(3) angvel(3) q(nj)
qd(nj), so tests/test_pybullet_fixtures.py can inject them through oracle_set_state / solo_set_state unchanged.
SoloBase keeps a class-level ``loaded`` flag (solo.py:15): one robot per process, hence one worker process per
configuration.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROBOTS = {"solo8": "solo.urdf", "solo12": "solo12.urdf"}
TASKS = ("stand", "walk", "pointgoal")
CONTROLS = ("torque", "pd")
MAX_POINTS = 8          # contact points kept per (link, ground) pair


def import_reference(reference):
    """Make ``import soloRL.baseEnv`` work for a checkout at `reference` whatever its directory is called."""
    reference = os.path.abspath(reference)
    parent, name = os.path.split(reference.rstrip("/"))
    if name != "soloRL":
        link_dir = tempfile.mkdtemp(prefix="solorl_ref_")
        os.symlink(reference, os.path.join(link_dir, "soloRL"))
        parent = link_dir
    sys.path.insert(0, parent)
    from soloRL.baseEnv import SoloBaseEnv      # noqa: E402  (the reference, unmodified)
    return SoloBaseEnv


def snapshot(p, robot):
    """State row in this repository's layout from the same PyBullet queries the reference makes
    (solo.py:201-210)."""
    pos, orn = p.getBasePositionAndOrientation(robot.id)
    lin, ang = p.getBaseVelocity(robot.id)
    js = [j.get_state() for j in robot.ordered_joints]
    q = [s[0] for s in js]
    qd = [s[1] for s in js]
    return np.array(list(pos) + list(orn) + list(lin) + list(ang) + q + qd, dtype=np.float64)


def contact_record(p, robot):
    """Contact points of every robot link (base = -1) against the ground, as p.getContactPoints returns them
    (solo.py:313-317): rows of (link index, contactDistance [8], normalForce [9], position on the robot xyz [6],
    lateral friction forces [10], [12]); NaN-padded to a fixed number of rows."""
    rows = []
    for c in p.getContactPoints(bodyA=robot.ground_id, bodyB=robot.id):
        rows.append([c[4], c[8], c[9], c[6][0], c[6][1], c[6][2], c[10], c[12]])
    out = np.full((MAX_POINTS * 6, 8), np.nan)
    for i, r in enumerate(rows[:len(out)]):
        out[i] = r
    return out, len(rows)


def worker(args):
    SoloBaseEnv = import_reference(args.reference)
    import pybullet as p
    robot_name, task, control = args.worker
    urdf = os.path.join(os.path.abspath(args.reference), "solo_description", "robots", ROBOTS[robot_name])
    config = {"model_urdf": urdf, "mode": "direct", "episode_length": args.episode_length, "frame_skip": 4,
              "control": control, "task": task, "num_history_stack": 1, "flat_ground": True,
              "use_treadmill": False}
    if control == "pd":
        config["gains"] = [5.0, 0.2]                      # configs/basic_pd.yaml:6
    env = SoloBaseEnv(config)
    if control != "torque":
        # get_reward reads an undefined `torque` for control != 'torque' (baseEnv.py:142-146); the
        # physics and observations are still recordable, the reward is recorded as NaN
        pass
    robot = env.robot
    np.random.seed(args.seed)
    rng = np.random.default_rng(args.seed)
    nj = len(robot.ordered_joints)

    engine = {k: (v if isinstance(v, (int, float, str)) else list(v)) for k, v in p.getPhysicsEngineParameters().items()}
    dyn = {}
    for link in range(-1, p.getNumJoints(robot.id)):
        dyn[str(link)] = [x if isinstance(x, (int, float)) else list(x) for x in p.getDynamicsInfo(robot.id, link)]
    joints = {}
    for j in range(p.getNumJoints(robot.id)):
        info = p.getJointInfo(robot.id, j)
        joints[str(j)] = [x.decode() if isinstance(x, bytes) else (list(x) if isinstance(x, tuple) else x) for x in info]
    dyn["ground"] = [x if isinstance(x, (int, float)) else list(x) for x in p.getDynamicsInfo(robot.ground_id, -1)]

    rec = {k: [] for k in ("step_pre", "step_action", "step_post", "step_obs", "step_reward", "step_done",
                           "step_goal_pre", "step_goal_post", "step_contacts", "step_ncontacts", "step_success",
                           "step_timeout", "step_timestep",
                           "sub_pre", "sub_tau", "sub_post", "sub_contacts", "sub_ncontacts")}
    obs = env.reset()
    D = obs.shape[0]
    for t in range(args.steps):
        pre = snapshot(p, robot)
        goal_pre = np.array(robot.goal_xy, dtype=np.float64) if task == "pointgoal" else np.zeros(2)
        a = rng.uniform(-1.2, 1.2, size=env.action_space.shape[0])
        try:
            o, r, d, info = env.step(a)
        except UnboundLocalError:                          # F9a: reward undefined for pd control
            o, r, d, info = env.get_observation(), float("nan"), False, {}
            env.timestep += 0
        post = snapshot(p, robot)
        con, ncon = contact_record(p, robot)
        rec["step_pre"].append(pre); rec["step_action"].append(a); rec["step_post"].append(post)
        rec["step_obs"].append(np.full(D, np.nan) if o is None else np.asarray(o, dtype=np.float64))
        rec["step_reward"].append(r); rec["step_done"].append(float(d))
        rec["step_goal_pre"].append(goal_pre)
        rec["step_goal_post"].append(np.array(robot.goal_xy, dtype=np.float64) if task == "pointgoal" else np.zeros(2))
        rec["step_contacts"].append(con); rec["step_ncontacts"].append(ncon)
        rec["step_success"].append(float(info.get("success", False))); rec["step_timeout"].append(float(info.get("timeout", False)))
        rec["step_timestep"].append(env.timestep)
        if d:
            env.reset()
    # single-substep transitions on the same world: torque applied for exactly one stepSimulation
    env.reset()
    for t in range(args.substeps):
        pre = snapshot(p, robot)
        tau = rng.uniform(-3.0, 3.0, size=nj) * (rng.random() < 0.8)
        p.setJointMotorControlArray(robot.id, robot.joints_idx, controlMode=p.TORQUE_CONTROL, forces=list(tau))
        p.stepSimulation()
        post = snapshot(p, robot)
        con, ncon = contact_record(p, robot)
        rec["sub_pre"].append(pre); rec["sub_tau"].append(tau); rec["sub_post"].append(post)
        rec["sub_contacts"].append(con); rec["sub_ncontacts"].append(ncon)
        if post[2] < 0.1 or (t + 1) % 60 == 0:
            env.reset()
    out = {k: np.asarray(v) for k, v in rec.items()}
    import pybullet
    out["engine_json"] = np.array(json.dumps(engine))
    out["dynamics_json"] = np.array(json.dumps({"links": dyn, "joints": joints}))
    out["pybullet_version"] = np.array(json.dumps({"api": p.getAPIVersion(),
                                                   "package": getattr(pybullet, "__version__", "unknown")}))
    out["config_json"] = np.array(json.dumps({**config, "model_urdf": ROBOTS[robot_name], "robot": robot_name}))
    out["feet_idx"] = np.asarray(robot.feet_idx)
    out["joints_idx"] = np.asarray(robot.joints_idx)
    os.makedirs(args.out, exist_ok=True)
    path = os.path.join(args.out, f"{robot_name}_{task}_{control}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if v.ndim})
    env.close()


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", required=True, help="checkout of michel-aractingi/soloRL (unmodified)")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                  "tests", "golden", "pybullet"))
    ap.add_argument("--steps", type=int, default=60, help="env-level transitions per configuration")
    ap.add_argument("--substeps", type=int, default=240, help="single-stepSimulation transitions per configuration")
    ap.add_argument("--episode-length", type=int, default=25)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--worker", nargs=3, metavar=("ROBOT", "TASK", "CONTROL"), help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.worker:
        return worker(args)
    try:
        import pybullet  # noqa: F401
    except ImportError:
        raise SystemExit("pybullet is not installed here: run this script where the reference's dependencies are "
                         "(pip install pybullet gym); tests/test_pybullet_fixtures.py skips loudly until "
                         "tests/golden/pybullet/*.npz exist")
    failed = []
    for robot in ROBOTS:
        for task in TASKS:
            for control in CONTROLS:
                cmd = [sys.executable, os.path.abspath(__file__), "--reference", args.reference, "--out", args.out,
                       "--steps", str(args.steps), "--substeps", str(args.substeps), "--episode-length",
                       str(args.episode_length), "--seed", str(args.seed), "--worker", robot, task, control]
                if subprocess.call(cmd) != 0:
                    failed.append((robot, task, control))
    if failed:
        raise SystemExit(f"failed configurations: {failed}")


if __name__ == "__main__":
    main()
