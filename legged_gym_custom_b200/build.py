"""Builds legged_gym_custom_b200/libb200gym.so in-tree with nvcc for sm_100a.

Per-TU flags matter: the env / GAE translation units are compiled with -fmad=false because
their parity contract is "round every op like torch does" (csrc/env_core.cuh); the GEMM /
learner units keep FMA contraction.  `python -m legged_gym_custom_b200.build` or
__graft_entry__.build() runs this; it cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb200gym.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xptxas", "-v"]

# (source, extra flags)
UNITS = [
    ("env_kernels.cu", ["-fmad=false"]),
    ("gae_kernels.cu", ["-fmad=false"]),
    ("learner_kernels.cu", []),
    ("linear_kernels.cu", []),
    ("mlp_tcgen05.cu", []),
    ("dist_adam.cu", []),
    ("terrain_kernels.cu", ["-fmad=false"]),
]


def _newer(target, deps):
    return os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(d) for d in deps)


def build_lib(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "b200gym.h"))
    objs, rebuilt = [], False
    for src, extra in UNITS:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(obj)
        if not force and _newer(obj, [path] + headers):
            continue
        cmd = ["nvcc"] + ARCH + COMMON + extra + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
        if verbose:
            print(r.stderr)
        rebuilt = True
    if rebuilt or not os.path.exists(LIB):
        cmd = ["nvcc"] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-lcuda"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
