"""Terrain construction on the device (csrc/terrain_kernels.cu) against the host generators, which
are pinned to the reference's own Terrain / convert_heightfield_to_trimesh (tests/golden/terrain_sha.json,
tests/test_terrain_trimesh.py): bit-identical height fields, vertices and triangles."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import golden_util as gu
from legged_gym_custom_b200 import configs, terrain as tm

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("task", ["go2_parkour", "go2_parkour_finetune"])
def test_parkour_field_on_the_device_matches_reference_sha(task):
    tcfg = configs.TASKS[task][0].terrain
    field, origins = tm.make_parkour_terrain_gpu(tcfg, DEV)
    host, host_origins = tm.make_parkour_terrain(tcfg)
    f = field.cpu().numpy()
    assert f.shape == host.shape and np.array_equal(f, host) and np.array_equal(origins, host_origins)
    sha = json.load(open(os.path.join(gu.GOLDEN_DIR, "terrain_sha.json")))[task]
    assert hashlib.sha256(f.tobytes()).hexdigest() == sha["sha256"]              # = the reference's Terrain.height_field_raw


@pytest.mark.parametrize("slope_threshold", [None, 0.75])
def test_trimesh_on_the_device_matches_host(slope_threshold):
    tcfg = configs.TASKS["go2_parkour"][0].terrain
    host, _ = tm.make_parkour_terrain(tcfg)
    rng = np.random.default_rng(0)
    patch = host[200:700, 200:620].copy()
    patch[100:300, 50:250] += rng.integers(-30, 31, (200, 200)).astype(np.int16)       # slopes on both sides of the threshold
    v_ref, t_ref = tm.heightfield_to_trimesh(patch, tcfg.horizontal_scale, tcfg.vertical_scale, slope_threshold)
    v, t = tm.heightfield_to_trimesh_gpu(torch.from_numpy(patch).to(DEV), tcfg.horizontal_scale, tcfg.vertical_scale, slope_threshold)
    torch.cuda.synchronize()
    assert np.array_equal(v.cpu().numpy(), v_ref)
    assert np.array_equal(t.cpu().numpy().view(np.uint32), t_ref)
    if slope_threshold is not None:
        plain, _ = tm.heightfield_to_trimesh(patch, tcfg.horizontal_scale, tcfg.vertical_scale, None)
        assert not np.array_equal(plain, v_ref)                                      # the correction did move vertices


def test_full_parkour_trimesh_on_the_device():
    """the whole 3860 x 2500 field: 9.65 M vertices, 19.3 M triangles"""
    tcfg = configs.TASKS["go2_parkour"][0].terrain
    field, _ = tm.make_parkour_terrain_gpu(tcfg, DEV)
    v, t = tm.heightfield_to_trimesh_gpu(field, tcfg.horizontal_scale, tcfg.vertical_scale, 0.75)
    torch.cuda.synchronize()
    rows, cols = field.shape
    assert v.shape == (rows * cols, 3) and t.shape == (2 * (rows - 1) * (cols - 1), 3)
    assert int(t.max()) == rows * cols - 1 and int(t.min()) == 0
    z = v[:, 2].view(rows, cols)
    assert torch.equal(z, field.float() * np.float32(1.0) * 0 + (field.double() * tcfg.vertical_scale).float())
