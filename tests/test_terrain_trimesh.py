"""heightfield -> trimesh (simulator-side half) against the reference's convert_heightfield_to_trimesh
(terrain_utils.py:401-465), run live where /root/reference exists, and against closed-form properties everywhere."""
import numpy as np
import pytest

from legged_gym_custom_b200 import terrain
from oracle import ref_runner


def _field(rng, rows, cols):
    hf = (rng.integers(-3, 4, (rows, cols)) * 20).astype(np.int16)
    hf[rows // 2:, : cols // 2] += 400                       # a 2 m step: far beyond any slope threshold
    return hf


def test_trimesh_structure():
    rng = np.random.default_rng(0)
    hf = _field(rng, 7, 5)
    v, t = terrain.heightfield_to_trimesh(hf, 0.1, 0.005, None)
    assert v.dtype == np.float32 and t.dtype == np.uint32 and v.shape == (35, 3) and t.shape == (2 * 6 * 4, 3)
    np.testing.assert_allclose(v[:, 2].reshape(7, 5), hf * 0.005, rtol=1e-6)
    np.testing.assert_allclose(v[:, 0].reshape(7, 5)[:, 0], np.arange(7) * 0.1, rtol=1e-6)
    # every cell is covered by two triangles over its four corners, consistently wound (positive z normal on a flat grid)
    flat, tf = terrain.heightfield_to_trimesh(np.zeros((7, 5), np.int16), 0.1, 0.005, None)
    a, b, c = flat[tf[:, 0]], flat[tf[:, 1]], flat[tf[:, 2]]
    nz = np.cross(b - a, c - a)[:, 2]
    assert (np.abs(nz) > 0).all() and (np.sign(nz) == np.sign(nz[0])).all()
    assert np.isclose(0.5 * np.abs(nz).sum(), (6 * 0.1) * (4 * 0.1))
    # a slope threshold only moves vertices sideways, by whole cells
    vs, ts = terrain.heightfield_to_trimesh(hf, 0.1, 0.005, 0.75)
    assert np.array_equal(ts, t) and np.array_equal(vs[:, 2], v[:, 2])
    shift = np.round((vs[:, :2] - v[:, :2]) / 0.1)
    assert set(np.unique(shift)) <= {-1.0, 0.0, 1.0} and (shift != 0).any()


@pytest.mark.skipif(not ref_runner.reference_available(), reason="reference not present (GPU box)")
@pytest.mark.parametrize("slope_threshold", [None, 0.75])
def test_trimesh_matches_reference(slope_threshold):
    ref_runner._setup_path()
    import isaacgym  # noqa: F401
    from legged_gym.envs import task_registry  # noqa: F401  (the reference's import order)
    from legged_gym.utils import terrain_utils
    rng = np.random.default_rng(1)
    for rows, cols in ((2, 2), (9, 4), (63, 130)):
        hf = _field(rng, rows, cols)
        v0, t0 = terrain_utils.convert_heightfield_to_trimesh(hf.copy(), 0.1, 0.005, slope_threshold)
        v1, t1 = terrain.heightfield_to_trimesh(hf, 0.1, 0.005, slope_threshold)
        assert v0.dtype == v1.dtype and t0.dtype == t1.dtype
        assert np.array_equal(v0, v1) and np.array_equal(t0, t1), (rows, cols)
