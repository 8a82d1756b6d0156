"""In-tree build of the CUDA C-ABI library for sm_100a (B200).

``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` cross-compiles without a GPU;
the resulting ``solorl_b200/libsolo_b200.so`` is git-ignored but travels with the tree.
"""
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_PKG, "csrc", "solo_kernels.cu")
_DEPS = [_SRC] + [os.path.join(_PKG, "csrc", f) for f in ("solo_core.cuh", "solo_env.cuh", "solo_host_model.h", "solo_wide.cuh")] + [
    os.path.join(os.path.dirname(_PKG), "include", "solo_b200.h")]
LIB = os.path.join(_PKG, "libsolo_b200.so")
_BENCH_SRC = os.path.join(_PKG, "csrc", "bench_util.cu")
BENCH_LIB = os.path.join(_PKG, "libsolo_benchutil.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def build(force: bool = False, verbose: bool = False) -> str:
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in _DEPS)
    if force or stale:
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, _SRC]
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        subprocess.check_call(cmd, env=env)
    if force or (not os.path.exists(BENCH_LIB)) or os.path.getmtime(_BENCH_SRC) > os.path.getmtime(BENCH_LIB):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        subprocess.check_call([nvcc] + NVCC_FLAGS + ["-o", BENCH_LIB, _BENCH_SRC], env=env)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
