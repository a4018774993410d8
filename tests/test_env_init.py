"""Env-creation-time domain randomisation: the kernel source (host emulation) and the CUDA kernel
against the restatement in oracle/init_oracle.py -- bit-exact, same keyed draws -- plus the distributional facts of
legged_robot.py:306-380 / :696-701 / :897-930 (ranges, 64 friction buckets, every terrain column populated)."""
import ctypes as C

import numpy as np
import pytest
import torch

import golden_util as gu
from legged_gym_custom_b200 import configs
from legged_gym_custom_b200.buffers import BufferSet
from legged_gym_custom_b200.params import env_params_from_cfg, init_params_from_cfg
from oracle import init_oracle

KEYS = ("priv_friction", "priv_mass_params", "kp_kd_multipliers", "terrain_levels", "terrain_types", "env_origins")


def _setup(task, N, device):
    cfg = configs.TASKS[task][0]
    hs, origins = gu.terrain_for(task)
    p = env_params_from_cfg(cfg, num_envs=N, seed=77, hs_shape=None if hs is None else hs.shape)
    bufs = BufferSet(p, device)
    if hs is not None:
        bufs["terrain_origins"].copy_(torch.from_numpy(np.asarray(origins, np.float32)))
    ip = init_params_from_cfg(cfg, N, hs is not None)
    return cfg, p, bufs, ip, origins


def _check(bufs, ref, cfg, ip, N):
    for k in KEYS:
        got = bufs[k].cpu().numpy()
        assert np.array_equal(got, ref[k].reshape(got.shape)), k
    kp = ref["kp_kd_multipliers"]
    assert kp.min() >= ip.kp_kd_lo and kp.max() <= ip.kp_kd_hi and kp.shape == (2, N, 12)
    if ip.randomize_friction:
        fr = ref["priv_friction"]
        assert len(np.unique(fr)) <= 64 and fr.min() >= ip.friction_lo and fr.max() <= ip.friction_hi
    if ip.num_init_levels > 0:
        assert ref["terrain_levels"].max() < ip.num_init_levels and set(ref["terrain_types"]) == set(range(ip.terrain_cols))


@pytest.mark.parametrize("task,N", [("go2_parkour", 4096), ("go2", 333)])
def test_env_init_kernel_source_matches_oracle(task, N):
    from host_emul import emul
    lib = emul.load()
    cfg, p, bufs, ip, origins = _setup(task, N, "cpu")
    lib.emul_env_init(C.byref(p), C.byref(bufs.struct), C.byref(ip))
    _check(bufs, init_oracle.init_randomisation(77, N, ip, origins), cfg, ip, N)


@pytest.mark.gpu
@pytest.mark.parametrize("task,N", [("go2_parkour", 4096), ("go2_parkour_finetune", 1000), ("go2", 333)])
def test_env_init_cuda_matches_oracle(task, N):
    from legged_gym_custom_b200 import _lib
    lib = _lib.lib()
    cfg, p, bufs, ip, origins = _setup(task, N, "cuda:0")
    h = C.c_void_p()
    _lib.check(lib.b200_env_create(C.byref(p), 0, C.byref(h)))
    _lib.check(lib.b200_env_init_randomisation(h, C.byref(bufs.struct), C.byref(ip), _lib.stream_ptr()))
    torch.cuda.synchronize()
    _check(bufs, init_oracle.init_randomisation(77, N, ip, origins), cfg, ip, N)
    lib.b200_env_destroy(h)


@pytest.mark.parametrize("task", ["go2", "go2_parkour", "go2_parkour_finetune"])
def test_init_oracle_deterministic_parts_match_reference_golden(task):
    """Pins oracle/init_oracle.py to what the UNMODIFIED reference produced (tests/golden/env_*.npz, oracle/make_golden.py):
    everything _get_env_origins (legged_robot.py:897-930) computes without a random draw -- the terrain type of every env,
    the origin looked up from (level, type), the env_spacing grid of the plane task -- must be identical; the random parts
    (drawn from torch's global stream in the reference, from keyed Philox here) must share the reference's support: ranges,
    at most 64 distinct friction values, every level below max_init_terrain_level + 1."""
    g = gu.load(task)
    N = int(g["init/env_origins"].shape[0])
    cfg = configs.TASKS[task][0]
    has_hf = cfg.terrain.mesh_type in ("heightfield", "trimesh")
    ip = init_params_from_cfg(cfg, N, has_hf)
    origins = g["static/terrain_origins"] if has_hf else None
    ref = init_oracle.init_randomisation(123, N, ip, origins)
    if has_hf:
        assert np.array_equal(ref["terrain_types"], g["init/terrain_types"])
        # the lookup, with the reference's own levels: env_origins = terrain_origins[level, type]
        lv, ty = g["init/terrain_levels"], g["init/terrain_types"]
        assert np.array_equal(np.asarray(origins, np.float32)[lv, ty], g["init/env_origins"])
        assert np.array_equal(ref["env_origins"], np.asarray(origins, np.float32)[ref["terrain_levels"], ref["terrain_types"]])
        assert ref["terrain_levels"].max() < ip.num_init_levels and ref["terrain_levels"].min() >= 0
    else:
        assert np.array_equal(ref["env_origins"], g["init/env_origins"])              # the plane grid is fully deterministic
        assert not ref["terrain_levels"].any()
    # support of the random parts, on the reference's draws and on ours
    for kp in (g["static/kp_kd_multipliers"], ref["kp_kd_multipliers"]):
        assert kp.shape == (2, N, 12) and kp.min() >= ip.kp_kd_lo and kp.max() <= ip.kp_kd_hi
    for fr in (g["static/privileged_friction_coeffs"], ref["priv_friction"]):
        assert len(np.unique(fr)) <= 64
        if ip.randomize_friction:
            assert fr.min() >= ip.friction_lo and fr.max() <= ip.friction_hi
    for m in (g["static/privileged_mass_params"], ref["priv_mass_params"]):
        m = np.asarray(m).reshape(N, 4)
        if ip.randomize_base_mass:
            assert m[:, 0].min() >= ip.mass_lo and m[:, 0].max() <= ip.mass_hi
        if ip.randomize_com:
            assert m[:, 1:].min() >= ip.com_lo and m[:, 1:].max() <= ip.com_hi
