"""The default terrain curriculum (legged_gym/utils/terrain.py:86-100, :134-192 -- the rough-terrain layout of BASELINE
config 2): the deterministic tiles (smooth slopes, stairs up / down, gap) must equal the reference's own Terrain output --
live where /root/reference exists, by sha everywhere -- and the noise tiles must have the reference's support."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

from legged_gym_custom_b200 import terrain as tm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "terrain_curriculum_sha.json")


class TCfg:
    mesh_type, horizontal_scale, vertical_scale, border_size = "heightfield", 0.1, 0.005, 25
    curriculum, parkour, selected, add_roughness_to_selected_terrain = True, False, False, False
    terrain_length = terrain_width = 8.
    num_rows, num_cols = 10, 20
    slope_treshold = 0.75
    terrain_proportions = [0.3, 0.0, 0.3, 0.3, 0.0, 0.0, 0.0]      # slopes | stairs down | stairs up | (rest) gap: no random tile


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_deterministic_curriculum_matches_reference_sha():
    field, origins = tm.make_curriculum_terrain(TCfg)
    want = json.load(open(GOLDEN))
    assert list(field.shape) == want["shape"] and _sha(field) == want["sha256"] and _sha(origins) == want["origins_sha256"]


def test_deterministic_curriculum_matches_reference_live():
    if not os.path.isdir("/root/reference/legged_gym"):
        pytest.skip("authoring container only")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in ("/root/reference/rsl_rl", "/root/reference", os.path.join(root, "oracle", "refshim")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import isaacgym  # noqa: F401  (oracle/refshim)
    import legged_gym.envs  # noqa: F401  (resolves the reference's circular import)
    from legged_gym.utils.terrain import Terrain
    ref = Terrain(TCfg, 64)
    field, origins = tm.make_curriculum_terrain(TCfg)
    assert np.array_equal(field, ref.height_field_raw)
    assert np.array_equal(origins, ref.env_origins.astype(np.float32))
    assert _sha(field) == json.load(open(GOLDEN))["sha256"]


def test_default_proportions_noise_tiles_have_the_reference_support():
    class Cfg(TCfg):
        terrain_proportions = [0.1, 0.1, 0.35, 0.25, 0.2, 0.0, 0.0]      # legged_robot_config.py:46
    field, origins = tm.make_curriculum_terrain(Cfg, seed=3)
    again, _ = tm.make_curriculum_terrain(Cfg, seed=3)
    other, _ = tm.make_curriculum_terrain(Cfg, seed=4)
    assert np.array_equal(field, again) and not np.array_equal(field, other)
    b, px = 250, 80
    tile = lambda i, j: field[b + i * px:b + (i + 1) * px, b + j * px:b + (j + 1) * px]
    det, _ = tm.make_curriculum_terrain(TCfg)
    assert np.array_equal(tile(5, 0), det[b + 5 * px:b + 6 * px, b:b + px])                # column 0 is a smooth slope either way
    # rough slope (columns 2-3) = slope + noise within +-0.06 m (12 units of 5 mm), in steps of 5 mm
    smooth = tm._pyramid_sloped(np.zeros((px, px), np.int16), 0.5 * 0.5, 3., 0.1, 0.005)
    noise = tile(5, 2).astype(int) - smooth
    assert noise.min() >= -12 and noise.max() <= 12 and len(np.unique(noise)) > 5
    # discrete obstacles (columns 16-19): only the four heights + 0, central platform clear
    h = int((0.05 + 0.5 * 0.15) / 0.005)
    t = tile(5, 17)
    assert set(np.unique(t)) <= {-h, -h // 2, 0, h // 2, h} and not t[25:55, 25:55].any()
    assert origins.shape == (10, 20, 3) and np.isfinite(origins).all()
