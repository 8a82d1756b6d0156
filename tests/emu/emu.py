"""ctypes wrapper of tests/emu/solo_emu.cpp: the CUDA kernel's lane program replayed on
the CPU in fp32 (TEST HARNESS ONLY; see the header of solo_emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

from solorl_b200.abi import SoloEpisodeStats, SoloModelTable, SoloSimParams, model_table

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_LIB = os.path.join(_ROOT, "tests", "_build", "libsolo_emu.so")
_lib = None


def build():
    srcs = [os.path.join(_HERE, "solo_emu.cpp")] + [
        os.path.join(_ROOT, "solorl_b200", "csrc", f)
        for f in ("solo_core.cuh", "solo_env.cuh", "solo_host_model.h", "solo_body.cuh")] + [
        os.path.join(_ROOT, "include", "solo_b200.h")]
    if (not os.path.exists(_LIB)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared",
                               "-ffp-contract=off", "-Wno-unknown-pragmas", "-o", _LIB, srcs[0]])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        fp = C.POINTER(C.c_float)
        vp = C.c_void_p
        L.emu_create.restype = vp
        L.emu_create.argtypes = [C.POINTER(SoloModelTable), C.POINTER(SoloSimParams), C.c_uint64, C.c_int64]
        L.emu_destroy.argtypes = [vp]
        L.emu_obs_dim.argtypes = [vp]
        L.emu_set_state.argtypes = [vp, fp]
        L.emu_get_state.argtypes = [vp, fp]
        L.emu_set_goal.argtypes = [vp, C.c_float, C.c_float]
        L.emu_forward_dynamics.argtypes = [vp, fp, fp]
        L.emu_substep.argtypes = [vp, fp]
        L.emu_get_contacts.argtypes = [vp, fp]
        L.emu_action_to_torque.argtypes = [vp, fp, fp]
        L.emu_get_observation.argtypes = [vp, fp]
        L.emu_reset.argtypes = [vp, fp]
        L.emu_settle_count_last.argtypes = [vp]
        L.emu_step.argtypes = [vp, fp, C.c_int, fp, fp, C.POINTER(C.c_int), C.POINTER(SoloEpisodeStats)]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class EmuEnv:
    def __init__(self, model, params, seed=0, env_id=0):
        self.L = lib()
        self.table = model_table(model)
        self.h = self.L.emu_create(C.byref(self.table), C.byref(params), seed, env_id)
        assert self.h, "emu_create failed"
        self.nj = model.nj
        self.d = self.L.emu_obs_dim(self.h)
        self.act_dim = self.nj + (2 if params.control == 2 else 0)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.emu_destroy(self.h)
            self.h = None

    def set_state(self, s):
        s = np.ascontiguousarray(s, np.float32)
        self.L.emu_set_state(self.h, _fp(s))

    def get_state(self):
        s = np.zeros(13 + 2 * self.nj, np.float32)
        self.L.emu_get_state(self.h, _fp(s))
        return s

    def set_goal(self, gx, gy):
        self.L.emu_set_goal(self.h, gx, gy)

    def forward_dynamics(self, tau):
        tau = np.ascontiguousarray(tau, np.float32)
        out = np.zeros(6 + self.nj, np.float32)
        self.L.emu_forward_dynamics(self.h, _fp(tau), _fp(out))
        return out

    def substep(self, tau):
        tau = np.ascontiguousarray(tau, np.float32)
        self.L.emu_substep(self.h, _fp(tau))

    def get_contacts(self):
        out = np.zeros((4, 3), np.float32)
        self.L.emu_get_contacts(self.h, _fp(out))
        return out

    def action_to_torque(self, a):
        a = np.ascontiguousarray(a, np.float32)
        tau = np.zeros(self.nj, np.float32)
        self.L.emu_action_to_torque(self.h, _fp(a), _fp(tau))
        return tau

    def get_observation(self):
        o = np.zeros(self.d, np.float32)
        self.L.emu_get_observation(self.h, _fp(o))
        return o

    def reset(self):
        o = np.zeros(self.d, np.float32)
        self.L.emu_reset(self.h, _fp(o))
        return o

    @property
    def settle_count_last(self):
        return self.L.emu_settle_count_last(self.h)

    def step(self, a, auto_reset=False):
        a = np.ascontiguousarray(a, np.float32)
        o = np.zeros(self.d, np.float32)
        r = C.c_float()
        d = C.c_int()
        st = SoloEpisodeStats()
        rc = self.L.emu_step(self.h, _fp(a), int(auto_reset), _fp(o), C.byref(r), C.byref(d), C.byref(st))
        if rc != 0:
            raise AssertionError("env.reset() must be called before step")
        return o, r.value, bool(d.value), {k: getattr(st, k) for k, _ in st._fields_}
