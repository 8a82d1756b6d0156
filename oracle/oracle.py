"""ctypes wrapper of the CPU ORACLE (``oracle/solo_oracle.c``).

TEST INFRASTRUCTURE ONLY.  May be imported by ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs; never by the
``solorl_b200`` package (tests/test_boundary.py greps for that).
PARITY UNPINNED for the physics (see solo_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from solorl_b200.abi import SoloModelTable, SoloSimParams, model_table
from solorl_b200.model import SoloModel

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsolo_oracle.so")
_lib = None


class OracleInfo(C.Structure):
    _fields_ = [("episode_reward", C.c_double), ("episode_return", C.c_double),
                ("episode_length", C.c_int), ("success", C.c_int), ("timeout", C.c_int),
                ("goals_reached", C.c_int), ("dr_stand", C.c_double),
                ("dr_joint_pose", C.c_double), ("dr_torque", C.c_double),
                ("dr_balance", C.c_double), ("dr_progress", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "solo_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(_LIB_PATH)
        for p in (src, os.path.join(_HERE, "solo_oracle.h"),
                  os.path.join(_HERE, "..", "include", "solo_b200.h")))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        fp = C.POINTER(C.c_float)
        vp = C.c_void_p
        L.oracle_default_params.argtypes = [C.POINTER(SoloSimParams)]
        L.oracle_env_create.restype = vp
        L.oracle_env_create.argtypes = [C.POINTER(SoloModelTable), C.POINTER(SoloSimParams),
                                        C.c_uint64, C.c_int64]
        L.oracle_env_destroy.argtypes = [vp]
        for f in ("oracle_nj", "oracle_act_dim", "oracle_obs_dim0", "oracle_obs_dim",
                  "oracle_settle_count_last", "oracle_last_solver_iters", "oracle_last_limit_rows"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = C.c_int
        L.oracle_env_reset.argtypes = [vp, dp]
        L.oracle_env_step.argtypes = [vp, dp, C.c_int, dp, dp, C.POINTER(C.c_int),
                                      C.POINTER(OracleInfo)]
        L.oracle_env_step.restype = C.c_int
        L.oracle_get_observation.argtypes = [vp, dp]
        L.oracle_get_current_state.argtypes = [vp, dp]
        L.oracle_get_state.argtypes = [vp, dp]
        L.oracle_set_state.argtypes = [vp, dp]
        L.oracle_set_goal.argtypes = [vp, C.c_double, C.c_double]
        L.oracle_get_goal.argtypes = [vp, dp]
        L.oracle_set_goal_radius.argtypes = [vp, C.c_double]
        L.oracle_action_to_torque.argtypes = [vp, dp, dp]
        L.oracle_forward_dynamics.argtypes = [vp, dp, dp]
        L.oracle_forward_dynamics_crba.argtypes = [vp, dp, dp, dp]
        L.oracle_substep.argtypes = [vp, dp]
        L.oracle_get_contacts.argtypes = [vp, dp]
        L.oracle_set_contacts.argtypes = [vp, dp]
        L.oracle_set_external_force.argtypes = [vp, dp]
        ip = C.POINTER(C.c_int)
        L.oracle_contact_rows.argtypes = [vp, dp, dp, dp, dp, ip, ip, dp]
        L.oracle_contact_rows.restype = C.c_int
        L.oracle_energy.argtypes = [vp]
        L.oracle_energy.restype = C.c_double
        L.oracle_foot_positions.argtypes = [vp, dp]
        L.oracle_gae.argtypes = [fp, fp, fp, fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
        L.oracle_batch_reset.argtypes = [C.POINTER(vp), C.c_int, fp, C.c_int]
        L.oracle_batch_step.argtypes = [C.POINTER(vp), C.c_int, fp, fp, fp, fp, C.c_int]
        L.oracle_batch_substep.argtypes = [C.POINTER(vp), C.c_int, dp, dp, dp, dp, C.POINTER(C.c_int), C.c_int]
        L.oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def default_params() -> SoloSimParams:
    p = SoloSimParams()
    lib().oracle_default_params(C.byref(p))
    return p


class OracleEnv:
    """One double-precision env; mirrors SoloBaseEnv (baseEnv.py:6-187)."""

    def __init__(self, model: SoloModel, params: SoloSimParams, seed: int = 0, env_id: int = 0):
        self.L = lib()
        self.model = model
        self.table = model_table(model)
        self.params = params
        self.h = self.L.oracle_env_create(C.byref(self.table), C.byref(params), seed, env_id)
        self.nj = self.L.oracle_nj(self.h)
        self.act_dim = self.L.oracle_act_dim(self.h)
        self.d0 = self.L.oracle_obs_dim0(self.h)
        self.d = self.L.oracle_obs_dim(self.h)

    def __del__(self):
        try:
            if self.h:
                self.L.oracle_env_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def reset(self):
        obs = np.zeros(self.d)
        self.L.oracle_env_reset(self.h, _dp(obs))
        return obs

    def step(self, action, auto_reset=False):
        a = np.ascontiguousarray(action, dtype=np.float64)
        assert a.shape == (self.act_dim,)
        obs = np.zeros(self.d)
        r = C.c_double()
        d = C.c_int()
        info = OracleInfo()
        rc = self.L.oracle_env_step(self.h, _dp(a), int(auto_reset), _dp(obs), C.byref(r),
                                    C.byref(d), C.byref(info))
        if rc != 0:
            raise AssertionError("env.reset() must be called before step")  # baseEnv.py:43
        return obs, r.value, bool(d.value), info.as_dict()

    def get_observation(self):
        obs = np.zeros(self.d)
        self.L.oracle_get_observation(self.h, _dp(obs))
        return obs

    def get_current_state(self):
        s = np.zeros(self.d0)
        self.L.oracle_get_current_state(self.h, _dp(s))
        return s

    def get_state(self):
        s = np.zeros(13 + 2 * self.nj)
        self.L.oracle_get_state(self.h, _dp(s))
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64)
        assert s.shape == (13 + 2 * self.nj,)
        self.L.oracle_set_state(self.h, _dp(s))

    def set_goal(self, gx, gy):
        self.L.oracle_set_goal(self.h, gx, gy)

    def get_goal(self):
        g = np.zeros(2)
        self.L.oracle_get_goal(self.h, _dp(g))
        return g

    def set_goal_radius(self, r):
        self.L.oracle_set_goal_radius(self.h, r)

    @property
    def settle_count_last(self):
        return self.L.oracle_settle_count_last(self.h)

    @property
    def last_solver_iters(self):
        return self.L.oracle_last_solver_iters(self.h)

    @property
    def last_limit_rows(self):
        return self.L.oracle_last_limit_rows(self.h)

    def action_to_torque(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        tau = np.zeros(self.nj)
        self.L.oracle_action_to_torque(self.h, _dp(a), _dp(tau))
        return tau

    def forward_dynamics(self, tau):
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        out = np.zeros(6 + self.nj)
        self.L.oracle_forward_dynamics(self.h, _dp(tau), _dp(out))
        return out

    def forward_dynamics_crba(self, tau, want_M=False):
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        out = np.zeros(6 + self.nj)
        M = np.zeros((6 + self.nj, 6 + self.nj))
        self.L.oracle_forward_dynamics_crba(self.h, _dp(tau), _dp(out), _dp(M))
        return (out, M) if want_M else out

    def substep(self, tau):
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        self.L.oracle_substep(self.h, _dp(tau))

    def get_contacts(self):
        out = np.zeros((4, 3))
        self.L.oracle_get_contacts(self.h, _dp(out))
        return out

    def set_contacts(self, force):
        f = np.ascontiguousarray(force, dtype=np.float64)
        assert f.shape == (4,)
        self.L.oracle_set_contacts(self.h, _dp(f))

    def set_external_force(self, f):
        f = np.ascontiguousarray(f, dtype=np.float64)
        assert f.shape == (3,)
        self.L.oracle_set_external_force(self.h, _dp(f))

    def contact_rows(self, tau):
        """Constraint rows the next substep(tau) would build (env not advanced): dict with J, U [rows, 6+nj],
        target, kind (0 normal, 1/2 friction, 3 joint limit), owner, vstar."""
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        nd, cap = 6 + self.nj, 48 + 2 * self.nj      # 16 contact points x 3 rows + two limit rows per joint
        J, U = np.zeros((cap, nd)), np.zeros((cap, nd))
        target, vstar = np.zeros(cap), np.zeros(nd)
        kind, owner = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        ip = C.POINTER(C.c_int)
        n = self.L.oracle_contact_rows(self.h, _dp(tau), _dp(J), _dp(U), _dp(target), kind.ctypes.data_as(ip),
                                       owner.ctypes.data_as(ip), _dp(vstar))
        return dict(J=J[:n], U=U[:n], target=target[:n], kind=kind[:n], owner=owner[:n], vstar=vstar)

    def energy(self):
        return self.L.oracle_energy(self.h)

    def foot_positions(self):
        out = np.zeros((4, 3))
        self.L.oracle_foot_positions(self.h, _dp(out))
        return out


def gae(rewards, values, masks, gamma, lam, use_gae=True, bootstrap_returns=None):
    """float32 GAE in the op order of agents/ppo/storage.py:35-55.
    rewards [T,N], values [T+1,N], masks [T+1,N] -> returns [T+1,N]."""
    T, N = rewards.shape
    r = np.ascontiguousarray(rewards, np.float32)
    v = np.ascontiguousarray(values, np.float32)
    m = np.ascontiguousarray(masks, np.float32)
    ret = np.zeros((T + 1, N), np.float32)
    if bootstrap_returns is not None:
        ret[T] = bootstrap_returns
    lib().oracle_gae(_fp(r), _fp(v), _fp(m), _fp(ret), T, N, gamma, lam, int(use_gae))
    return ret


class OracleVecEnv:
    """N oracle envs stepped with OpenMP: the CPU baseline leg of bench.py
    (``kind: "port"`` — a CPU restatement, NOT PyBullet)."""

    def __init__(self, model, params, num_envs, seed=0, env_id_offset=0, nthreads=None):
        self.L = lib()
        self.envs = [OracleEnv(model, params, seed, env_id_offset + i) for i in range(num_envs)]
        self.n = num_envs
        self.ptrs = (C.c_void_p * num_envs)(*[e.h for e in self.envs])
        self.nthreads = nthreads or self.L.oracle_max_threads()
        self.d = self.envs[0].d
        self.act_dim = self.envs[0].act_dim

    def reset(self):
        obs = np.zeros((self.n, self.d), np.float32)
        self.L.oracle_batch_reset(self.ptrs, self.n, _fp(obs), self.nthreads)
        return obs

    def substep_from(self, states, taus):
        """env i <- states[i], one substep with taus[i]: (next states, contacts [n,4,3], PGS sweeps [n])."""
        s = np.ascontiguousarray(states, np.float64)
        t = np.ascontiguousarray(taus, np.float64)
        assert s.shape[0] == self.n and t.shape[0] == self.n
        out, con = np.zeros_like(s), np.zeros((self.n, 4, 3))
        it = np.zeros(self.n, np.int32)
        self.L.oracle_batch_substep(self.ptrs, self.n, _dp(s), _dp(t), _dp(out), _dp(con),
                                    it.ctypes.data_as(C.POINTER(C.c_int)), self.nthreads)
        return out, con, it

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.float32)
        obs = np.zeros((self.n, self.d), np.float32)
        rew = np.zeros(self.n, np.float32)
        done = np.zeros(self.n, np.float32)
        self.L.oracle_batch_step(self.ptrs, self.n, _fp(a), _fp(obs), _fp(rew), _fp(done),
                                 self.nthreads)
        return obs, rew, done
