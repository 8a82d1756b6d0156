"""The drop-in boundary: the C-ABI library loads, exports every symbol include/solo_b200.h
declares, fails loudly without a GPU, and the product never touches the oracle."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

from solorl_b200 import _lib, abi, build
from solorl_b200.model import SoloModel
from tests.helpers import ROOT, make_config

HEADER = os.path.join(ROOT, "include", "solo_b200.h")


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(solo_[a-z_0-9]+)\s*\(", src)))


def test_header_functions_all_exported(lib_path):
    names = header_functions()
    assert len(names) >= 20
    assert sorted(_lib.SYMBOLS) == names
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path]).decode()
    exported = set(l.split()[-1] for l in out.splitlines() if " T " in l)
    for n in names:
        assert n in exported, f"{n} declared in the header but not exported"
    L = C.CDLL(lib_path)
    for n in names:
        getattr(L, n)


def test_library_is_sm100a_and_has_the_step_kernel(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sym = subprocess.check_output(["cuobjdump", "-elf", lib_path]).decode()
    assert "step_kernel" in sym and "gae_kernel" in sym


def test_default_params_three_way(lib_path):
    L = _lib.lib()
    p = abi.SoloSimParams()
    assert L.solo_default_params(C.byref(p)) == 0
    q = abi.default_params()
    for name, _ in p._fields_:
        assert getattr(p, name) == getattr(q, name), name


@pytest.mark.parametrize("robot,task,control,H,exp", [("solo8", "stand", "torque", 0, (8, 8, 30, 30)),
                                                       ("solo12", "walk", "torque", 1, (12, 12, 38, 76)),
                                                       ("solo12", "pointgoal", "vpd", 1, (12, 14, 42, 84))])
def test_solo_dims(robot, task, control, H, exp):
    L = _lib.lib()
    m = SoloModel.resolve(robot)
    p = abi.params_from_config(make_config(robot, task=task, control=control, H=H), m)
    t = abi.model_table(m)
    out = [C.c_int32() for _ in range(4)]
    assert L.solo_dims(C.byref(t), C.byref(p), *[C.byref(o) for o in out]) == 0
    assert tuple(o.value for o in out) == exp == abi.dims(m, p)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    L = _lib.lib()
    m = SoloModel.builtin("solo8")
    t, p = abi.model_table(m), abi.default_params()
    h = C.c_void_p()
    rc = L.solo_create(C.byref(t), C.byref(p), 4, 0, 0, 0, C.byref(h))
    assert rc == -3 and not h.value                       # SOLO_E_CUDA
    assert b"no CPU fallback" in L.solo_last_error(None)
    from solorl_b200.sim import SoloSim
    with pytest.raises(RuntimeError):
        SoloSim(m, p, 4)
    from solorl_b200.envs import SoloVecEnv
    with pytest.raises(RuntimeError):
        SoloVecEnv(make_config("solo8"), 4)


def test_create_rejects_bad_arguments():
    L = _lib.lib()
    m = SoloModel.builtin("solo8")
    t, p = abi.model_table(m), abi.default_params()
    h = C.c_void_p()
    assert L.solo_create(C.byref(t), C.byref(p), 0, 0, 0, 0, C.byref(h)) == -1      # SOLO_E_ARG
    assert L.solo_step(None, None, None, None, None, None) == -1
    assert L.solo_gae(None, None, None, None, 1, 1, 0.99, 0.95, 1, None) == -1


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/: the package, the CLIs and
    the developer tools outside tests/ must not."""
    bad = []
    walks = [os.walk(os.path.join(ROOT, d)) for d in ("solorl_b200", "tools", "training", "testing")]
    for dp, _, files in (x for w in walks for x in w):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                if f.endswith(".py"):
                    hit = re.search(r"^\s*(from|import)\s+(oracle|tests)\b", src, flags=re.M) or \
                        "libsolo_oracle" in src or "libsolo_emu" in src
                else:   # C/CUDA: no include of, or dlopen into, the oracle / the emulation harness
                    hit = re.search(r"#\s*include\s*[<\"][^>\"]*(oracle|emu)", src) or "dlopen" in src
                if hit:
                    bad.append(f)
    assert not bad, bad


def test_config_mapping_matches_reference_defaults():
    m = SoloModel.builtin("solo8")
    p = abi.params_from_config({"model_urdf": "solo8", "mode": "headless", "episode_length": 400}, m)
    assert (p.frame_skip, p.control, p.task, p.num_history_stack) == (4, 0, 0, 0)   # baseEnv.py:9-15
    p = abi.params_from_config(make_config("solo8", control="pd"), m)
    assert (p.kp, p.kd) == (5.0, 0.2)                                               # configs/basic_pd.yaml:6
    with pytest.raises(NotImplementedError):
        abi.params_from_config(make_config("solo8", control="bogus"), m)            # solo.py:253-254
    with pytest.raises(KeyError):
        abi.params_from_config({"model_urdf": "solo8", "mode": "headless"}, m)     # episode_length is required
    with pytest.raises(NotImplementedError):
        abi.params_from_config(make_config("solo8", flat_ground=False), m)


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """sizeof / offsetof of every struct of include/solo_b200.h as gcc lays them out, against the ctypes
    mirrors in solorl_b200/abi.py (a reordered or retyped field would silently corrupt the call)."""
    structs = {"SoloModelTable": abi.SoloModelTable, "SoloSimParams": abi.SoloSimParams,
               "SoloEpisodeStats": abi.SoloEpisodeStats}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for sname, cls in structs.items():
        lines.append(f'  printf("{sname} sizeof %zu\\n", sizeof({sname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{sname} {fname} %zu\\n", offsetof({sname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)]).decode().split("\n")
    seen = 0
    for l in out:
        if not l.strip():
            continue
        sname, fname, val = l.split()
        cls = structs[sname]
        if fname == "sizeof":
            assert C.sizeof(cls) == int(val), sname
        else:
            assert getattr(cls, fname).offset == int(val), (sname, fname)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in structs.values())
    assert abi.EPISODE_STATS_DTYPE.itemsize == C.sizeof(abi.SoloEpisodeStats)
    for fname, _ in abi.SoloEpisodeStats._fields_:
        assert abi.EPISODE_STATS_DTYPE.fields[fname][1] == getattr(abi.SoloEpisodeStats, fname).offset
