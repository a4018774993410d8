"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol that
include/b200gym.h declares, the ctypes mirrors have the library's struct sizes, and argument
errors come back as return codes + text (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

from legged_gym_custom_b200 import _lib
from legged_gym_custom_b200.params import EnvBuffers, EnvParams, REWARD_TERMS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b200gym.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200gym.h but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes signature in _lib.SYMBOLS"


def test_struct_mirrors_match_library():
    lib = _lib.lib()
    assert lib.b200_abi_version() == 3
    assert lib.b200_env_params_size() == C.sizeof(EnvParams)
    assert lib.b200_env_buffers_size() == C.sizeof(EnvBuffers)


def test_reward_enum_matches_header():
    text = open(os.path.join(ROOT, "include", "b200gym.h")).read()
    enum = re.search(r"enum B200RewardTerm \{(.*?)\};", text, flags=re.S).group(1)
    names = re.findall(r"B200_REW_(\w+)", enum)
    assert names == REWARD_TERMS


def test_argument_errors_are_return_codes():
    lib = _lib.lib()
    h = C.c_void_p()
    p = EnvParams()
    assert lib.b200_env_create(C.byref(p), 0, C.byref(h)) == -1        # abi_version 0
    assert b"abi_version" in lib.b200_last_error()
    assert lib.b200_compute_returns(None, None, None, None, None, None, 24, 16, 0.99, 0.95, None, None) == -1
    assert b"null" in lib.b200_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(-1)
