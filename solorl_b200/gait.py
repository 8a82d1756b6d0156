"""Gait-selection env shell on the GPU simulator.

The reference's ``SoloGaitEnvContact`` (``soloGaitEnvContact.py:24-67`` on ``baseControlEnv.py:33-478``)
wraps two objects that are NOT part of the reference tree — ``Controller`` (LAAS-Gepetto
quadruped-reactive-walking: OSQP MPC + pinocchio whole-body control) and ``PyBulletSimulator`` — imported
with ``from scripts import Controller, PyBulletSimulator`` (``baseControlEnv.py:7``).  What is buildable
from the reference alone, and is built here, batched over N envs on the device:

* :class:`ActuatorSim` — the simulator surface the env shell drives: ``UpdateMeasurment``, ``q_mes``,
  ``v_mes``, ``baseState``, ``baseOrientation``, ``b_baseVel``, ``baseAngularVelocity``, ``tau_ff``,
  ``SetDesiredJoint{PDgains,Position,Velocity,Torque}``, ``SendCommand`` (``baseControlEnv.py:256-270,
  359-367,447-451``) on top of ``solo_actuator_step`` (one 0.002 s tick under the joint PD + feed-forward
  law);
* :class:`SoloGaitVecEnv` — the env shell: ``k_rl = rl_dt/dt`` controller ticks per RL step, reward
  ``1 - (20 E_pen + vel_pen)/k_rl`` with the joint-power model of ``get_joints_power`` (:425-445),
  termination ``z < 0.11 | timestep >= episode_length`` (:389-408), NaN guard (:171-175), the 64-d
  observation (``soloGaitEnvContact.py:54-67``), ``Discrete(9)`` contact patterns (:11-20), velocity
  reference switching (:302-312), masked auto-reset;
* a PLUGGABLE batched controller (``compute(robot, gait, v_ref) -> P, D, q_des, v_des, tau_ff``).
  :class:`PostureGaitController` is a stand-in (joint-space posture hold that tucks the legs the chosen
  contact pattern lifts); it is NOT the reference's MPC — physics and controller parity for this path
  are unpinned by construction, the shell arithmetic is what the tests pin.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .abi import default_params
from .envs import Box
from .model import SoloModel
from .sim import SoloSim

# baseControlEnv.py:13-15
COULOMB_TAU = 0.0477
VISCOUS_B = 0.000135
K_MOTOR = 4.81
VMAX = 0.8                                   # :20
MAXFORCE = 10                                # :21
DEFAULTFORCE = (7, 10)                       # :25  [min, max] push magnitude in N
PUSH_DURATIONS = (1000, 2000, 3000, 4000, 5000)   # :26  in controller ticks
Q_INIT = (0.0, 0.7, -1.4, -0.0, 0.7, -1.4, 0.0, -0.7, +1.4, -0.0, -0.7, +1.4)   # :41
# soloGaitEnvContact.py:11-20 (index 9 = the "-1" entry: no gait yet)
GAIT_TABLE = ((1., 1., 1., 1.), (1., 1., 1., 0.), (1., 1., 0., 1.), (1., 0., 1., 1.), (0., 1., 1., 1.),
              (1., 0., 1., 0.), (0., 1., 0., 1.), (1., 0., 0., 1.), (0., 1., 1., 0.), (0., 0., 0., 0.))
GAIT_NAMES = ("Static", "Walk1", "Walk2", "Walk3", "Walk4", "Pace1", "Pace2", "Trot1", "Trot2")


class Discrete:
    """Stand-in for gym.spaces.Discrete (the trainers read ``.n`` and the class name)."""

    def __init__(self, n):
        self.n = int(n)
        self.shape = ()

    def sample(self):
        return int(np.random.randint(self.n))


def quat_rotate_inverse(q, v):
    """R(q)^T v for quaternions (x,y,z,w) [N,4] and vectors [N,3]."""
    qv, w = q[:, :3], q[:, 3:4]
    t = 2.0 * torch.cross(qv, v, dim=1)
    return v - w * t + torch.cross(qv, t, dim=1)


def quat_to_rpy(q):
    """p.getEulerFromQuaternion [3P]: roll, pitch, yaw from (x,y,z,w)."""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    roll = torch.atan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z)
    pitch = torch.asin((-2 * (x * z - w * y)).clamp(-1, 1))
    yaw = torch.atan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z)
    return torch.stack([roll, pitch, yaw], dim=1)


class ActuatorSim:
    """Batched counterpart of the external ``PyBulletSimulator`` object the gait envs hold as
    ``self.robot`` (attribute and method names kept, including the ``UpdateMeasurment`` spelling)."""

    def __init__(self, num_envs, solo12=True, dt=0.002, device=0, seed=0, q_init=Q_INIT, z_init=None, **params):
        self.model = SoloModel.builtin("solo12" if solo12 else "solo8")
        p = default_params()
        p.dt = float(dt)
        p.frame_skip = 1
        p.episode_length = 1 << 30
        for k, v in params.items():
            setattr(p, k, v)
        self.params, self.dt = p, float(dt)
        self.sim = SoloSim(self.model, p, num_envs, device=device, seed=seed)
        self.n, self.nj, self.device = num_envs, self.model.nj, self.sim.device
        qi = torch.tensor(q_init, dtype=torch.float32, device=self.device)
        if self.nj == 8:
            qi = qi.reshape(4, 3)[:, 1:].reshape(-1)
        self.q_init = qi
        self._state0 = torch.zeros(self.n, 13 + 2 * self.nj, device=self.device)
        self._state0[:, 6] = 1.0
        self._state0[:, 13:13 + self.nj] = qi
        self.sim.set_state(self._state0)
        # rest height: lowest foot sphere touches the ground at this posture
        if z_init is None:
            z_init = float(-self.sim.get_feet()[0, :, 2].min().item() + self.model.foot_radius)
        self.z_init = z_init
        self._state0[:, 2] = z_init
        f = dict(dtype=torch.float32, device=self.device)
        self.P = torch.zeros(self.n, self.nj, **f)
        self.D = torch.zeros(self.n, self.nj, **f)
        self.q_des = self.q_init.repeat(self.n, 1)
        self.v_des = torch.zeros(self.n, self.nj, **f)
        self.tau_ff = torch.zeros(self.n, self.nj, **f)
        self.reset()

    # ---- PyBulletSimulator surface -------------------------------------------------------
    def reset(self, mask=None):
        if mask is None:
            self.sim.set_state(self._state0)
        else:
            cur = self.sim.get_state()
            self.sim.set_state(torch.where(mask.reshape(-1, 1) > 0.5, self._state0, cur))
        self.UpdateMeasurment()

    def UpdateMeasurment(self):
        s = self.sim.get_state()
        self.baseState = (s[:, 0:3], s[:, 3:7])            # ((x,y,z), quaternion) like pyb.getBasePositionAndOrientation
        self.baseOrientation = s[:, 3:7]
        self.baseLinearVelocity = s[:, 7:10]               # world frame
        self.baseAngularVelocity = quat_rotate_inverse(s[:, 3:7], s[:, 10:13])   # body frame (get_base_vel)
        self.b_baseVel = quat_rotate_inverse(s[:, 3:7], s[:, 7:10])
        self.q_mes = s[:, 13:13 + self.nj]
        self.v_mes = s[:, 13 + self.nj:]
        self._state = s

    def SetDesiredJointPDgains(self, P, D):
        self.P = torch.as_tensor(P, dtype=torch.float32, device=self.device).expand(self.n, self.nj)
        self.D = torch.as_tensor(D, dtype=torch.float32, device=self.device).expand(self.n, self.nj)

    def SetDesiredJointPosition(self, q_des):
        self.q_des = torch.as_tensor(q_des, dtype=torch.float32, device=self.device).expand(self.n, self.nj)

    def SetDesiredJointVelocity(self, v_des):
        self.v_des = torch.as_tensor(v_des, dtype=torch.float32, device=self.device).expand(self.n, self.nj)

    def SetDesiredJointTorque(self, tau_ff):
        self.tau_ff = torch.as_tensor(tau_ff, dtype=torch.float32, device=self.device).expand(self.n, self.nj)

    def SendCommand(self, WaitEndOfCycle=False, n_ticks=1):
        cmd = torch.stack([self.q_des, self.v_des, self.P, self.D, self.tau_ff], dim=1).contiguous()
        self.sim.actuator_step(cmd, n_ticks)

    def get_feet_positions(self):
        return self.sim.get_feet()

    def Stop(self):
        self.sim.close()


class PostureGaitController:
    """Stand-in for the external MPC/WBC ``Controller``: holds the nominal posture with joint PD and
    tucks (flexes) the legs that the selected contact pattern lifts, with a small hip pitch offset
    proportional to the commanded forward velocity.  Pluggable: anything with the same ``compute``
    signature can replace it."""

    def __init__(self, robot: ActuatorSim, kp=3.0, kd=0.2, tuck=0.35):
        self.kp, self.kd, self.tuck = kp, kd, tuck
        self.gait = torch.tensor(GAIT_TABLE, dtype=torch.float32, device=robot.device)
        self.njl = robot.nj // 4
        self.error = torch.zeros(robot.n, dtype=torch.bool, device=robot.device)

    def reset(self, mask=None):
        if mask is None:
            self.error.zero_()
        else:
            self.error &= ~(mask > 0.5)

    def compute(self, robot, gait, v_ref):
        contact = self.gait[gait]                                   # [N,4], 1 = stance
        swing = (1.0 - contact).unsqueeze(-1)                       # [N,4,1]
        q = robot.q_init.reshape(4, self.njl).unsqueeze(0).repeat(robot.n, 1, 1)
        sign = torch.sign(q[:, :, -2:-1])                           # front legs flex with +, hind legs with -
        off = torch.zeros_like(q)
        off[:, :, -2] = self.tuck
        off[:, :, -1] = -2.0 * self.tuck
        q = q + swing * sign * off
        q[:, :, -2] = q[:, :, -2] - 0.2 * v_ref[:, 0:1]
        q_des = q.reshape(robot.n, -1)
        zeros = torch.zeros_like(q_des)
        return (torch.full_like(q_des, self.kp), torch.full_like(q_des, self.kd), q_des, zeros, zeros)


class SoloGaitVecEnv:
    """N ``SoloGaitEnvContact`` envs (config 4, ``configs/basic_contact.yaml``)."""

    OBS_DIM = 64          # soloGaitEnvContact.py:36-38

    def __init__(self, config, num_envs, device=None, seed=0, controller_factory=None, cuda_graph=True,
                 env_id_offset=0):
        self.config = dict(config)
        self.env_id_offset = int(env_id_offset)
        self.dt = float(config.get("dt", 0.002))                       # baseControlEnv.py:37
        self.T_gait = float(config.get("T_gait", 0.32))
        self.rl_dt = self.T_gait / 2                                   # soloGaitEnvContact.py:27-28
        self.k_rl = int(self.rl_dt / self.dt)                          # baseControlEnv.py:58 (= 80)
        self.episode_length = int(config.get("episode_length", 100))
        if not config.get("flat_ground", True):
            raise NotImplementedError("flat_ground: False")
        self.add_external_force = bool(config.get("add_external_force", False))       # baseControlEnv.py:54
        self.auto_vel_switch = bool(config.get("auto_vel_switch", True))
        self.vel_switch = int(config.get("vel_switch", 30))
        self.use_curriculum = bool(config.get("use_curriculum", False))
        self.max_velocity = 0.0 if self.use_curriculum else VMAX       # :99
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.nenvs = int(num_envs)
        self.robot = ActuatorSim(num_envs, solo12=bool(config.get("solo12", True)), dt=self.dt, device=dev.index or 0,
                                 seed=seed)
        if self.robot.nj != 12:
            raise NotImplementedError("the 64-d observation of SoloGaitEnvContact assumes Solo12")
        self.controller = (controller_factory or PostureGaitController)(self.robot)
        self.num_actions = 9
        self.action_space = Discrete(self.num_actions)
        self.observation_space = Box(-np.inf * np.ones(self.OBS_DIM), np.inf * np.ones(self.OBS_DIM))
        f = dict(dtype=torch.float32, device=dev)
        self.gait_table = torch.tensor(GAIT_TABLE, **f)
        # one stream per shard: ranks of a multi-GPU run (env_id_offset = rank * N) must not draw identical
        # velocity references / pushes
        self.gen = torch.Generator(device=dev).manual_seed(int(seed) * 1000003 + self.env_id_offset)
        self.k_rl_ticks_total = 0
        self.timestep = torch.zeros(self.nenvs, dtype=torch.int32, device=dev)
        self.past_gaits = torch.full((self.nenvs, 3), 9, dtype=torch.long, device=dev)     # deque([-1,-1,-1])
        self.vel_ref = torch.zeros(self.nenvs, 6, **f)
        self.vel_mask = torch.zeros(6, **f)                            # `mask` of baseControlEnv.py:28 (all zero!)
        self.ep_reward = torch.zeros(self.nenvs, **f)
        self.ep_length = torch.zeros(self.nenvs, dtype=torch.int32, device=dev)
        self.dr = torch.zeros(self.nenvs, 3, **f)                      # Torque_pen, body_velocity, Energy_pen
        # random pushes (baseControlEnv.py:117-123,276-289): one per episode, a force along one base axis with an
        # integer magnitude in [min, max] N, from a random controller tick on, for 1000..5000 ticks
        if self.add_external_force:
            mm = (0, 2) if self.use_curriculum else DEFAULTFORCE
        else:
            mm = (0, 0)
        self.min_max_force = [int(mm[0]), int(mm[1])]
        self.k_tick = torch.zeros(self.nenvs, dtype=torch.int32, device=dev)             # controller.k
        self.push_F = torch.zeros(self.nenvs, 3, **f)
        self.push_start = torch.zeros(self.nenvs, dtype=torch.int32, device=dev)
        self.push_dur = torch.ones(self.nenvs, dtype=torch.int32, device=dev)
        self._durations = torch.tensor(PUSH_DURATIONS, dtype=torch.int32, device=dev)
        self.last_info = None
        self._was_reset = False
        self.cuda_graph = bool(cuda_graph)
        self._graph = None

    # ---- helpers ---------------------------------------------------------------------------
    def _new_random_vel(self, n):
        """new_random_vel (baseControlEnv.py:29-32): uniform in +-max_velocity times the module mask."""
        v = (torch.rand(n, 6, device=self.device, generator=self.gen) - 0.5) * 2 * self.max_velocity
        return v * self.vel_mask

    def _create_force_function(self, sel):
        """create_force_function (baseControlEnv.py:276-289) for the envs in `sel` (bool [N]); tensors are updated
        in place because the captured tick loop reads them."""
        n, dev, g = self.nenvs, self.device, self.gen
        lo, hi = self.min_max_force
        axis = torch.randint(0, 3, (n,), device=dev, generator=g)
        mag = torch.randint(lo, hi + 1, (n,), device=dev, generator=g).float()        # np.random.randint(min, max + 1)
        sign = torch.randint(0, 2, (n,), device=dev, generator=g).float() * 2 - 1
        F = torch.zeros(n, 3, device=dev)
        F.scatter_(1, axis.unsqueeze(1), mag.unsqueeze(1))
        F = F * torch.stack([sign, sign, torch.ones_like(sign)], dim=1)                # F *= [sign, sign, 1]
        hi_start = max(500, int(self.k_rl * self.episode_length * (2.0 / 3.0)))
        start = torch.randint(500, hi_start + 1, (n,), device=dev, generator=g, dtype=torch.int32)
        dur = self._durations[torch.randint(0, len(PUSH_DURATIONS), (n,), device=dev, generator=g)]
        m = sel.unsqueeze(1)
        self.push_F.copy_(torch.where(m, F, self.push_F))
        self.push_start.copy_(torch.where(sel, start, self.push_start))
        self.push_dur.copy_(torch.where(sel, dur, self.push_dur))

    def _apply_force(self):
        """_apply_force(controller.k): [3P] PyBulletSimulator.apply_external_force(k, start, duration, F, M) of the
        external simulator -- while start <= k <= start + duration the base is pushed with alpha(k) F in its own
        frame, alpha = 16 s^2 (1 - s)^2, s = (k - start) / duration (a smooth bump that peaks at 1 mid-way;
        restated from the LAAS quadruped-reactive-walking sources, not in the reference tree)."""
        ev = (self.k_tick - self.push_start).float()
        dur = self.push_dur.float()
        inside = (ev >= 0) & (ev <= dur)
        sfrac = (ev / dur).clamp(0.0, 1.0)
        alpha = 16.0 * sfrac * sfrac * (1.0 - sfrac) * (1.0 - sfrac)
        self.robot.sim.set_external_force(torch.where(inside.unsqueeze(1), alpha.unsqueeze(1) * self.push_F,
                                                      torch.zeros_like(self.push_F)))

    def get_base_vel(self):
        return torch.cat([self.robot.b_baseVel, self.robot.baseAngularVelocity], dim=1)   # :447-451

    def get_joints_power(self):
        """baseControlEnv.py:425-445: P = tau_f qd + K tau_cmd^2, tau_f = tau_c sign(qd) + b qd."""
        qd, tau = self.robot.v_mes, self.robot.tau_ff
        tau_f = COULOMB_TAU * torch.sign(qd) + VISCOUS_B * qd
        return tau_f * qd + K_MOTOR * tau * tau

    def get_observation(self):
        """soloGaitEnvContact.py:54-67: [z, rpy(3), v_body(6), q(12), qd(12), feet xyz(12), 3 past contact
        patterns (12), v_ref(6)] = 64."""
        r = self.robot
        r.UpdateMeasurment()
        qu = torch.cat([r.baseState[0][:, 2:3], quat_to_rpy(r.baseOrientation)], dim=1)
        pfeet = r.get_feet_positions().reshape(self.nenvs, 12)
        seq = self.gait_table[self.past_gaits].reshape(self.nenvs, 12)
        return torch.cat([qu, self.get_base_vel(), r.q_mes, r.v_mes, pfeet, seq, self.vel_ref], dim=1)

    def _terminated(self):
        """get_termination (baseControlEnv.py:389-408) without the external controller's error flags."""
        fallen = (self.robot.baseState[0][:, 2] < 0.11) | self.controller.error
        timeout = (self.timestep >= self.episode_length) & ~fallen
        return fallen | timeout, timeout

    # ---- gym-style vec API ------------------------------------------------------------------
    def reset(self, mask=None):
        m = None if mask is None else (mask > 0.5)
        self.robot.reset(None if m is None else m.float())
        self.controller.reset(None if m is None else m.float())
        sel = torch.ones(self.nenvs, dtype=torch.bool, device=self.device) if m is None else m
        if self.auto_vel_switch:
            self.vel_ref = torch.where(sel.unsqueeze(1), self._new_random_vel(self.nenvs), self.vel_ref)
        self.past_gaits = torch.where(sel.unsqueeze(1), torch.full_like(self.past_gaits, 9), self.past_gaits)
        self.k_tick.masked_fill_(sel, 0)                                           # a fresh controller: k = 0
        if self.add_external_force:
            self._create_force_function(sel)                                        # baseControlEnv.py:213
        self.timestep.masked_fill_(sel, 0)
        self.ep_reward = torch.where(sel, torch.zeros_like(self.ep_reward), self.ep_reward)
        self.ep_length = torch.where(sel, torch.zeros_like(self.ep_length), self.ep_length)
        self.dr = torch.where(sel.unsqueeze(1), torch.zeros_like(self.dr), self.dr)
        self._was_reset = True
        return self.get_observation()

    def step(self, action, snapshot=False):
        assert self._was_reset, "env.reset() must be called before step"          # baseControlEnv.py:135
        a = torch.as_tensor(action, device=self.device).long().reshape(self.nenvs)
        self.past_gaits = torch.cat([self.past_gaits[:, 1:], a.unsqueeze(1)], dim=1)   # soloGaitEnvContact.py:42
        self.timestep += 1                       # in place: the captured tick loop reads this tensor
        torque_pen, vel_pen, joints_power = self._run_ticks(a)
        if self.auto_vel_switch:                                                   # switch_velocities :302-312
            sw = (self.timestep % self.vel_switch) == 0
            self.vel_ref = torch.where(sw.unsqueeze(1), self._new_random_vel(self.nenvs), self.vel_ref)
        obs = self.get_observation()
        done, timeout = self._terminated()
        energy_pen = joints_power.sum(1) * self.dt
        reward = 1.0 - (1.0 / self.k_rl) * (20.0 * energy_pen + vel_pen)           # :169-170
        nan = torch.isnan(obs).any(dim=1) | torch.isnan(reward)                    # :171-175
        obs = torch.where(nan.unsqueeze(1), torch.zeros_like(obs), obs)
        reward = torch.where(nan, torch.zeros_like(reward), reward)
        done = done | nan
        self.ep_length = self.ep_length + 1
        self.ep_reward = self.ep_reward + reward
        self.dr = self.dr + torch.stack([torque_pen, vel_pen, energy_pen], dim=1) / self.k_rl
        self.last_info = {"episode_length": self.ep_length.clone(), "episode_reward": self.ep_reward.clone(),
                          "success": timeout & done, "timeout": timeout, "nan": nan,
                          "max_velocity": self.max_velocity, "dr/Torque_pen": self.dr[:, 0].clone(),
                          "dr/body_velocity": self.dr[:, 1].clone(), "dr/Energy_pen": self.dr[:, 2].clone()}
        donef = done.float()
        if bool(done.any()):                                                       # worker auto-reset (envs.py:38-40)
            obs = torch.where(done.unsqueeze(1), self.reset(donef), obs)
        return obs, reward, donef, GaitInfos(self.last_info, donef)

    # ---- the k_rl controller ticks of one RL step (baseControlEnv.py:147-162) -----------------------------
    def _ticks(self, a):
        r = self.robot
        r.UpdateMeasurment()
        alive = ~self._terminated()[0]
        torque_pen = torch.zeros(self.nenvs, device=self.device)
        vel_pen = torch.zeros_like(torque_pen)
        joints_power = torch.zeros(self.nenvs, 12, device=self.device)
        for _ in range(self.k_rl):
            if self.add_external_force:
                self._apply_force()                                                  # baseControlEnv.py:148
            self.k_tick += 1
            r.UpdateMeasurment()
            P, D, q_des, v_des, tau_ff = self.controller.compute(r, a, self.vel_ref)
            r.SetDesiredJointPDgains(P, D)
            r.SetDesiredJointPosition(q_des)
            r.SetDesiredJointVelocity(v_des)
            r.SetDesiredJointTorque(tau_ff)
            r.SendCommand(WaitEndOfCycle=False)
            live = alive.float()
            torque_pen = torque_pen + live * (r.tau_ff ** 2).sum(1)
            vel_pen = vel_pen + live * ((self.vel_ref - self.get_base_vel()) ** 2).sum(1)
            joints_power = joints_power + live.unsqueeze(1) * self.get_joints_power()
            # `if done: break`: an env that terminates mid-step stops accumulating
            z = r.sim.get_state()[:, 2]
            alive = alive & ~(z < 0.11)
        return torque_pen, vel_pen, joints_power

    def _run_ticks(self, a):
        """Eagerly, or — the ~60 small launches per tick x 80 ticks cost far more CPU time than GPU time — as ONE
        CUDA graph captured on first use (the controller must then be capturable: tensor ops only, no host
        reads; pass cuda_graph=False otherwise)."""
        if not self.cuda_graph:
            return self._ticks(a)
        if self._graph is None:
            try:
                self._g_a = torch.zeros(self.nenvs, dtype=torch.long, device=self.device)
                self._g_vel = torch.zeros(self.nenvs, 6, device=self.device)
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                state, k0 = self.robot.sim.get_state().clone(), self.k_tick.clone()
                with torch.cuda.stream(s):                    # warm-up on a side stream, then restore the state
                    self._ticks(a)
                torch.cuda.current_stream().wait_stream(s)
                self.robot.sim.set_state(state)
                self.k_tick.copy_(k0)                         # the warm-up must not advance the controller clock
                real_vel = self.vel_ref
                self.vel_ref = self._g_vel                    # the captured program reads the static copies
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._g_out = self._ticks(self._g_a)
                self.vel_ref = real_vel
                self._graph = g
            except Exception as e:                            # pragma: no cover - capture is best effort
                print(f"[gait] CUDA graph capture of the tick loop failed ({e}); running it eagerly", flush=True)
                self.cuda_graph = False
                return self._ticks(a)
        self._g_a.copy_(a)
        self._g_vel.copy_(self.vel_ref)
        self._graph.replay()
        return tuple(t.clone() for t in self._g_out)

    def reset_vel_ref(self, vel):
        """baseControlEnv.py:236-238: impose the reference base velocity ([6] or [N, 6]); the controller reads it on
        its next tick.  `reset_vel` is the name testing/test_ppo.py:108 calls on the vec-env."""
        v = torch.as_tensor(np.asarray(vel, dtype=np.float32), device=self.device).reshape(-1, 6)
        self.vel_ref = v.expand(self.nenvs, 6).clone()

    reset_vel = reset_vel_ref

    def increment_curriculum(self, val=0.1):                                       # :321-328
        if not self.use_curriculum:
            return
        self.max_velocity = float(np.clip(self.max_velocity + val, 0.0, VMAX))
        lo, hi = self.min_max_force                                                # increment min and max force
        self.min_max_force = [int(np.clip(lo + 1, 0, MAXFORCE - 2)), int(np.clip(hi + 1, 0, MAXFORCE))]

    def get_torques(self):
        return self.robot.tau_ff

    def close(self):
        self.robot.Stop()

    def __len__(self):
        return self.nenvs


class GaitInfos:
    """Lazy per-env info dicts (keys of baseControlEnv.py:180-189) for envs that finished."""

    def __init__(self, info, done):
        self._info, self._done = info, done
        self._host = None

    def _fetch(self):
        if self._host is None:
            self._host = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in self._info.items()}
            self._host["_done"] = self._done.detach().cpu().numpy() > 0.5
        return self._host

    def __len__(self):
        return int(self._done.shape[0])

    def __getitem__(self, i):
        h = self._fetch()
        if not h["_done"][i]:
            return {}
        out = {}
        for k, v in h.items():
            if k == "_done":
                continue
            x = v[i] if isinstance(v, np.ndarray) else v
            out[k] = x.item() if isinstance(x, np.generic) else x
        out["min_force"], out["max_force"] = 0.0, 0.0
        return out

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


class SoloGaitEnvContact:
    """``env_constructor`` token for ``make_vec_envs(config, n, SoloGaitEnvContact)`` (``--env-name contact``,
    training/train_ppo.py:80-81); the batched implementation is :class:`SoloGaitVecEnv`."""

    def __new__(cls, config, device=None, seed=0):
        return SoloGaitVecEnv(config, 1, device=device, seed=seed)
