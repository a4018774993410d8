"""Parkour height-field generation (init-time, host numpy).

The height-scan kernel's input format is the reference's `height_samples`: int16
[tot_rows, tot_cols] in units of vertical_scale, rows = world x / horizontal_scale,
with a border of border_size metres (legged_robot.py:788-802, terrain.py:9-57).  This
module produces that array and `terrain_origins` [num_rows, num_cols, 3] for the two
parkour layouts the go2 tasks use -- `parkour_curriculum` (terrain.py:103-115,
:195-246) and `parkour_selected_terrain` (terrain.py:118-131) built on `parkour_terrain`
(terrain_utils.py:318-399) -- so tests and benchmarks on the GPU box need neither the
reference nor a 19 MB fixture.  tests/test_terrain.py pins it against the reference's
own `Terrain` output (sha256 in tests/golden/terrain_sha.json).

`heightfield_to_trimesh` is the simulator-side half of SURVEY.md §8(f1): the vertex / triangle arrays PhysX is fed
(terrain_utils.py:401-465, called from terrain.py:52-57), so that an Isaac Gym adapter needs nothing from the reference's
terrain modules.  Init-time host numpy (~1 s for the 3860 x 2500 parkour field, the same as the reference's row loop): the
arrays go to `gym.add_triangle_mesh`, never to our kernels.
"""
import numpy as np


def _course(length_px, width_px, h_scale, v_scale, start_platform_length, start_platform_height, x_positions,
            y_positions, obstacle_lengths, obstacle_heights, half_valid_width, border_width, border_height):
    """One obstacle course tile: rows = along-track (x), cols = across (y)."""
    hf = np.zeros((length_px, width_px), dtype=np.int16)
    hf[:round(start_platform_length / h_scale), :] = round(start_platform_height / v_scale)
    mid = width_px // 2
    half_gap = round(half_valid_width / h_scale)
    for x, y, length, height in zip(x_positions, y_positions, obstacle_lengths, obstacle_heights):
        cx, cy = round(x / h_scale), mid + round(y / h_scale)
        half = round(length / h_scale) // 2
        rows = slice(cx - half, cx + half)
        hf[rows, :] = round(height / v_scale)
        hf[rows, :cy - half_gap] = 0          # outside the valid corridor the obstacle is removed
        hf[rows, cy + half_gap:] = 0
    pad = int(border_width / h_scale)
    hf[:, :pad] = int(border_height / v_scale)
    hf[:, -pad:] = int(border_height / v_scale)
    return hf


def _curriculum_tile_kwargs(choice, difficulty, proportion0):
    """terrain.py:195-246 make_parkour_terrain: gaps widen / hurdles rise with difficulty."""
    if choice < proportion0:
        n, x0, dx = 7, 5.0, 3.5
        lengths, heights = [difficulty] * n, [-2.0] * n
    else:
        n, x0, dx = 14, 4.0, 1.99
        lengths, heights = [0.35] * n, [0.05 + 0.44 * difficulty] * n
    return dict(start_platform_length=3., start_platform_height=0., x_positions=list(np.arange(x0, x0 + n * dx, dx)),
                y_positions=[0.0] * n, obstacle_lengths=lengths, obstacle_heights=heights, half_valid_width=5.0,
                border_width=0.50, border_height=-2.0)


def make_parkour_terrain(tcfg):
    """-> (height_samples int16 [rows, cols], terrain_origins float32 [num_rows, num_cols, 3])."""
    hs, vs = tcfg.horizontal_scale, tcfg.vertical_scale
    width_px, length_px = int(tcfg.terrain_width / hs), int(tcfg.terrain_length / hs)
    border = int(tcfg.border_size / hs)
    rows = int(tcfg.num_rows * length_px) + 2 * border
    cols = int(tcfg.num_cols * width_px) + 2 * border
    field = np.zeros((rows, cols), dtype=np.int16)
    origins = np.zeros((tcfg.num_rows, tcfg.num_cols, 3))
    proportion0 = float(np.sum(tcfg.terrain_proportions[:1]))
    for j in range(tcfg.num_cols):
        for i in range(tcfg.num_rows):
            if tcfg.curriculum:
                kw = _curriculum_tile_kwargs(j / tcfg.num_cols + 0.001, (i + 1) / 10, proportion0)
            else:
                kw = tcfg.parkour_kwargs
            tile = _course(length_px, width_px, hs, vs, **kw)
            r0, c0 = border + i * length_px, border + j * width_px
            field[r0:r0 + length_px, c0:c0 + width_px] = tile
            origins[i, j] = [i * tcfg.terrain_length, (j + 0.5) * tcfg.terrain_width, 0.0]   # start line, centred in y
    return field, origins.astype(np.float32)


def heightfield_to_trimesh(height_field_raw, horizontal_scale, vertical_scale, slope_threshold=None):
    """-> (vertices float32 [rows*cols, 3], triangles uint32 [2*(rows-1)*(cols-1), 3]), identical to the reference's
    convert_heightfield_to_trimesh (terrain_utils.py:401-465).

    Vertex (i, j) sits at (i, j) * horizontal_scale, height * vertical_scale.  With a slope threshold, a vertex at the foot of
    a step steeper than the threshold is pulled under the step's edge (+1 cell) and a vertex at its top is pushed out over the
    foot (-1 cell), along x, along y, and -- where neither applies -- along the diagonal, which turns steep ramps into
    vertical walls.  Cell (i, j) contributes the triangles (v00, v11, v01) and (v00, v10, v11)."""
    hf = np.asarray(height_field_raw)
    rows, cols = hf.shape
    # float64 grid like np.linspace / np.meshgrid in the reference (the fp32 cast happens when the vertices are stored)
    yy, xx = np.meshgrid(np.linspace(0, (cols - 1) * horizontal_scale, cols), np.linspace(0, (rows - 1) * horizontal_scale, rows))
    if slope_threshold is not None:
        thr = slope_threshold * horizontal_scale / vertical_scale
        rise_x = np.diff(hf, axis=0)                      # hf[i+1, j] - hf[i, j]   (int16 arithmetic, as in the reference)
        rise_y = np.diff(hf, axis=1)
        rise_d = hf[1:, 1:] - hf[:-1, :-1]
        move_x, move_y, move_d = (np.zeros((rows, cols)) for _ in range(3))
        move_x[:-1, :] += rise_x > thr
        move_x[1:, :] -= -rise_x > thr
        move_y[:, :-1] += rise_y > thr
        move_y[:, 1:] -= -rise_y > thr
        move_d[:-1, :-1] += rise_d > thr
        move_d[1:, 1:] -= -rise_d > thr
        xx = xx + (move_x + move_d * (move_x == 0)) * horizontal_scale
        yy = yy + (move_y + move_d * (move_y == 0)) * horizontal_scale
    vertices = np.empty((rows * cols, 3), dtype=np.float32)
    vertices[:, 0], vertices[:, 1], vertices[:, 2] = xx.ravel(), yy.ravel(), hf.ravel() * vertical_scale
    v00 = (np.arange(rows - 1, dtype=np.uint32)[:, None] * np.uint32(cols) + np.arange(cols - 1, dtype=np.uint32)[None, :]).ravel()
    v01, v10, v11 = v00 + np.uint32(1), v00 + np.uint32(cols), v00 + np.uint32(cols + 1)
    triangles = np.empty((2 * v00.size, 3), dtype=np.uint32)
    triangles[0::2] = np.stack((v00, v11, v01), axis=1)
    triangles[1::2] = np.stack((v00, v10, v11), axis=1)
    return vertices, triangles
