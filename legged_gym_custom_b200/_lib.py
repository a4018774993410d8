"""ctypes binding of libb200gym.so (include/b200gym.h).  There is NO fallback: if the CUDA
library is missing or an entry point is absent, importing a product class raises."""
import ctypes as C
import os

from .params import EnvBuffers, EnvParams

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libb200gym.so")

# every symbol include/b200gym.h declares
SYMBOLS = {
    "b200_last_error": (C.c_char_p, []),
    "b200_abi_version": (C.c_int, []),
    "b200_env_params_size": (C.c_int, []),
    "b200_env_buffers_size": (C.c_int, []),
    "b200_env_create": (C.c_int, [C.POINTER(EnvParams), C.c_int, C.POINTER(C.c_void_p)]),
    "b200_env_destroy": (C.c_int, [C.c_void_p]),
    "b200_pd_torques": (C.c_int, [C.c_void_p, C.POINTER(EnvBuffers), C.c_void_p, C.c_int, C.c_void_p]),
    "b200_post_physics_step": (C.c_int, [C.c_void_p, C.POINTER(EnvBuffers), C.c_int64, C.c_void_p]),
    "b200_reset_all": (C.c_int, [C.c_void_p, C.POINTER(EnvBuffers), C.c_int64, C.c_int, C.c_void_p]),
    "b200_get_heights": (C.c_int, [C.c_void_p, C.POINTER(EnvBuffers), C.c_void_p]),
    "b200_gae_scratch_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "b200_compute_returns": (C.c_int, [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "b200_store_step_scalars": (C.c_int, [C.c_void_p] * 4 + [C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m legged_gym_custom_b200.build` "
                               "(there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)          # AttributeError = header and library disagree
            fn.restype, fn.argtypes = res, args
        if handle.b200_env_params_size() != C.sizeof(EnvParams) or handle.b200_env_buffers_size() != C.sizeof(EnvBuffers):
            raise RuntimeError("ctypes mirror of B200EnvParams/B200EnvBuffers is out of sync with the library")
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(f"libb200gym: {lib().b200_last_error().decode()} (rc={rc})")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
