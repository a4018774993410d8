"""ctypes driver of tests/host_emul/_env_emul.so (the CUDA kernel source compiled for the host)."""
import ctypes as C
import os
import subprocess

from legged_gym_custom_b200.params import EnvBuffers, EnvParams, InitParams

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_env_emul.so")
SRC = os.path.join(HERE, "env_emul.cpp")
CORE = os.path.join(os.path.dirname(os.path.dirname(HERE)), "legged_gym_custom_b200", "csrc")


def build(force=False):
    deps = [SRC] + [os.path.join(CORE, f) for f in ("env_core.cuh", "philox.cuh")] + \
           [os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "b200gym.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", "-x", "c++", SRC, "-o", SO])
    return SO


def load():
    lib = C.CDLL(build())
    assert lib.emul_params_size() == C.sizeof(EnvParams), (lib.emul_params_size(), C.sizeof(EnvParams))
    assert lib.emul_buffers_size() == C.sizeof(EnvBuffers), (lib.emul_buffers_size(), C.sizeof(EnvBuffers))
    lib.emul_pd_torques.argtypes = [C.POINTER(EnvParams), C.POINTER(EnvBuffers), C.c_void_p, C.c_int]
    lib.emul_post_physics_step.argtypes = [C.POINTER(EnvParams), C.POINTER(EnvBuffers), C.c_int64]
    lib.emul_post_physics_step_variant.argtypes = [C.POINTER(EnvParams), C.POINTER(EnvBuffers), C.c_int64, C.c_int]
    lib.emul_command_curriculum_rule.argtypes = [C.POINTER(EnvParams), C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.emul_env_init.argtypes = [C.POINTER(EnvParams), C.POINTER(EnvBuffers), C.POINTER(InitParams)]
    lib.emul_reset_all.argtypes = [C.POINTER(EnvParams), C.POINTER(EnvBuffers), C.c_int64, C.c_int]
    return lib
