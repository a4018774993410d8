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


# ---- the default terrain curriculum (Terrain.curriculum / make_terrain, terrain.py:86-100, :134-192): the layout of the rough
# terrain task (BASELINE config 2).  Rows = difficulty, columns = terrain choice by `terrain_proportions`.  The slope, stair and
# gap tiles are deterministic and identical to the reference's (pinned live and by sha, tests/test_terrain_curriculum.py).
# The noise tiles (rough slope, random uniform, discrete obstacles, stepping stones) draw from numpy's GLOBAL stream in the
# reference, so only their distribution can be matched: here they draw from a generator keyed by (seed, row, col), and
# random_uniform_terrain's scipy.interpolate.interp2d(kind='linear') -- removed from SciPy >= 1.14, where the reference's own
# call raises -- is restated as the bilinear interpolation it was.
def _pyramid_sloped(hf, slope, platform_size, hs, vs):
    """terrain_utils.py:72-93"""
    length, width = hf.shape
    ctr_x, ctr_y = width // 2, length // 2
    x = (ctr_x - np.abs(np.arange(width) - ctr_x)) / ctr_x
    y = (ctr_y - np.abs(np.arange(length) - ctr_y)) / ctr_y
    yy, xx = np.meshgrid(y, x, indexing="ij")
    max_h = int(slope * (hs / vs) * (width / 2))
    hf = hf + (max_h * xx * yy).astype(hf.dtype)
    half = int(platform_size / hs / 2)
    x1, y1 = ctr_x - half, ctr_y - half
    return np.clip(hf, min(hf[y1, x1], 0), max(hf[y1, x1], 0))


def _pyramid_stairs(hf, step_width, step_height, platform_size, hs, vs):
    """terrain_utils.py:151-166"""
    step_w, step_h, plat = int(step_width / hs), int(step_height / vs), int(platform_size / hs)
    length, width = hf.shape
    top, r0, r1, c0, c1 = 0, 0, length, 0, width
    while (r1 - r0) > plat and (c1 - c0) > plat:
        r0, r1, c0, c1 = r0 + step_w, r1 - step_w, c0 + step_w, c1 - step_w
        top += step_h
        hf[r0:r1, c0:c1] = top
    return hf


def _random_uniform(hf, rng, min_height, max_height, step, downsampled_scale, hs, vs):
    """terrain_utils.py:9-52 with interp2d(kind='linear') written out (bilinear on the regular down-sampled grid)"""
    length, width = hf.shape
    lo, hi, st = int(min_height / vs), int(max_height / vs), int(step / vs)
    choices = np.arange(lo, hi + st, st)
    down_rows, down_cols = int(length * hs / downsampled_scale), int(width * hs / downsampled_scale)
    z = rng.choice(choices, (down_rows, down_cols)).astype(np.float64)
    fy = np.linspace(0, length * hs, length) / (length * hs) * (down_rows - 1)
    fx = np.linspace(0, width * hs, width) / (width * hs) * (down_cols - 1)
    y0, x0 = np.minimum(fy.astype(int), down_rows - 2), np.minimum(fx.astype(int), down_cols - 2)
    ty, tx = (fy - y0)[:, None], (fx - x0)[None, :]
    z00, z01, z10, z11 = z[y0][:, x0], z[y0][:, x0 + 1], z[y0 + 1][:, x0], z[y0 + 1][:, x0 + 1]
    up = (z00 * (1 - tx) + z01 * tx) * (1 - ty) + (z10 * (1 - tx) + z11 * tx) * ty
    return hf + np.rint(up).astype(np.int16)


def _discrete_obstacles(hf, rng, max_height, min_size, max_size, num_rects, platform_size, hs, vs):
    """terrain_utils.py:95-119"""
    h_max, min_s, max_s, plat = int(max_height / vs), int(min_size / hs), int(max_size / hs), int(platform_size / hs)
    length, width = hf.shape
    heights = [-h_max, -h_max // 2, h_max // 2, h_max]
    for _ in range(num_rects):
        w, l = rng.choice(range(min_s, max_s, 4)), rng.choice(range(min_s, max_s, 4))
        row0, col0 = rng.choice(range(0, length - l, 4)), rng.choice(range(0, width - w, 4))
        hf[row0:row0 + l, col0:col0 + w] = rng.choice(heights)
    hf[(length - plat) // 2:(length + plat) // 2, (width - plat) // 2:(width + plat) // 2] = 0
    return hf


def _stepping_stones(hf, rng, stone_size, stone_distance, max_height, platform_size, hs, vs, depth=-10):
    """terrain_utils.py:168-210"""
    sz, gap, h_max, plat, pit = int(stone_size / hs), int(stone_distance / hs), int(max_height / vs), int(platform_size / hs), int(depth / vs)
    length, width = hf.shape
    hf[:] = pit
    heights = np.arange(-h_max - 1, h_max, 1)
    row = 0
    while row < length:
        row_end = min(length, row + sz)
        col = int(rng.integers(0, sz))
        hf[row:row_end, 0:max(0, col - gap)] = rng.choice(heights)
        while col < width:
            hf[row:row_end, col:min(width, col + sz)] = rng.choice(heights)
            col += sz + gap
        row += sz + gap
    hf[(length - plat) // 2:(length + plat) // 2, (width - plat) // 2:(width + plat) // 2] = 0
    return hf


def _gap(hf, gap_size, platform_size, hs):
    """terrain.py:322-335"""
    length, width = hf.shape
    g, p = int(gap_size / hs), int(platform_size / hs)
    cx, cy = length // 2, width // 2
    x1, y1 = (length - p) // 2, (width - p) // 2
    x2, y2 = x1 + g, y1 + g
    hf[cx - x2:cx + x2, cy - y2:cy + y2] = -1000
    hf[cx - x1:cx + x1, cy - y1:cy + y1] = 0
    return hf


def make_curriculum_tile(choice, difficulty, proportions, length_px, width_px, hs, vs, rng):
    """Terrain.make_terrain (terrain.py:134-192): one tile, int16 [length_px, width_px]"""
    hf = np.zeros((length_px, width_px), dtype=np.int16)
    slope, step_height = difficulty * 0.5, 0.05 + 0.115 * difficulty
    if choice < proportions[0]:
        return _pyramid_sloped(hf, -slope if choice < proportions[0] / 2 else slope, 3., hs, vs)
    if choice < proportions[1]:
        return _random_uniform(_pyramid_sloped(hf, slope, 3., hs, vs), rng, -0.06, 0.06, 0.005, 0.2, hs, vs)
    if choice < proportions[3]:
        return _pyramid_stairs(hf, 0.25, -step_height if choice < proportions[2] else step_height, 2., hs, vs)
    if choice < proportions[4]:
        return _discrete_obstacles(hf, rng, 0.05 + difficulty * 0.15, 1., 2., 20, 3., hs, vs)
    if choice < proportions[5]:
        return _stepping_stones(hf, rng, 1.5 * (1.05 - difficulty), 0.05 if difficulty == 0 else 0.1, 0., 4., hs, vs)
    if choice < proportions[6]:
        return _random_uniform(hf, rng, -0.06, 0.06, 0.005, 0.2, hs, vs)
    return _gap(hf, 1. * difficulty, 3., hs)


def make_curriculum_terrain(tcfg, seed=0):
    """Terrain.curriculum + add_terrain_to_map (terrain.py:86-100, :246-273)
    -> (height_samples int16 [rows, cols], terrain_origins float32 [num_rows, num_cols, 3])."""
    hs, vs = tcfg.horizontal_scale, tcfg.vertical_scale
    width_px, length_px = int(tcfg.terrain_width / hs), int(tcfg.terrain_length / hs)
    border = int(tcfg.border_size / hs)
    field = np.zeros((int(tcfg.num_rows * length_px) + 2 * border, int(tcfg.num_cols * width_px) + 2 * border), dtype=np.int16)
    origins = np.zeros((tcfg.num_rows, tcfg.num_cols, 3))
    proportions = [np.sum(tcfg.terrain_proportions[:i + 1]) for i in range(len(tcfg.terrain_proportions))]
    x1, x2 = int((tcfg.terrain_length / 2. - 1) / hs), int((tcfg.terrain_length / 2. + 1) / hs)
    y1, y2 = int((tcfg.terrain_width / 2. - 1) / hs), int((tcfg.terrain_width / 2. + 1) / hs)
    for j in range(tcfg.num_cols):
        for i in range(tcfg.num_rows):
            rng = np.random.default_rng([int(seed), i, j])
            tile = make_curriculum_tile(j / tcfg.num_cols + 0.001, i / tcfg.num_rows, proportions, length_px, width_px, hs, vs, rng)
            r0, c0 = border + i * length_px, border + j * width_px
            field[r0:r0 + length_px, c0:c0 + width_px] = tile
            origins[i, j] = [(i + 0.5) * tcfg.terrain_length, (j + 0.5) * tcfg.terrain_width, np.max(tile[x1:x2, y1:y2]) * vs]
    return field, origins.astype(np.float32)


def make_terrain(tcfg, seed=0):
    """the layout the reference's Terrain.__init__ picks for this cfg (terrain.py:33-47)"""
    if getattr(tcfg, "parkour", False):
        return make_parkour_terrain(tcfg)
    if getattr(tcfg, "curriculum", False):
        return make_curriculum_terrain(tcfg, seed)
    raise ValueError("terrain layout not supported: `selected` / randomized terrains (terrain.py:44-47) -- pass height_samples / terrain_origins")


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


# ---- the same two constructions on the device (csrc/terrain_kernels.cu): bit-identical to the host generators above ----
def _tile_table(length_px, width_px, h_scale, v_scale, start_platform_length, start_platform_height, x_positions, y_positions,
                obstacle_lengths, obstacle_heights, half_valid_width, border_width, border_height):
    """`_course` as a table: the reference's rounding and Python slice semantics resolved to explicit index ranges"""
    from . import _lib
    t = _lib.ParkourTile()
    t.platform_rows = min(length_px, max(0, round(start_platform_length / h_scale)))
    t.platform_height = round(start_platform_height / v_scale)
    mid, half_gap = width_px // 2, round(half_valid_width / h_scale)
    if len(x_positions) > 32:
        raise ValueError("more than 32 obstacles per tile")
    t.num_obstacles = len(x_positions)
    for k, (x, y, length, height) in enumerate(zip(x_positions, y_positions, obstacle_lengths, obstacle_heights)):
        cx, cy = round(x / h_scale), mid + round(y / h_scale)
        half = round(length / h_scale) // 2
        r0, r1, _ = slice(cx - half, cx + half).indices(length_px)
        _, z0, _ = slice(None, cy - half_gap).indices(width_px)          # hf[rows, :cy - half_gap] = 0
        z1, _, _ = slice(cy + half_gap, None).indices(width_px)          # hf[rows, cy + half_gap:] = 0
        t.row_lo[k], t.row_hi[k], t.zero_below[k], t.zero_from[k], t.height[k] = r0, max(r0, r1), z0, z1, round(height / v_scale)
    t.pad = int(border_width / h_scale)
    t.border_height = int(border_height / v_scale)
    if not 0 < t.pad <= width_px // 2:
        raise ValueError("side walls wider than half a tile")
    return t


def make_parkour_terrain_gpu(tcfg, device):
    """make_parkour_terrain on the device -> (height_samples int16 CUDA tensor [rows, cols], terrain_origins float32 numpy)"""
    import ctypes as C

    import torch

    from . import _lib
    hs, vs = tcfg.horizontal_scale, tcfg.vertical_scale
    width_px, length_px = int(tcfg.terrain_width / hs), int(tcfg.terrain_length / hs)
    border = int(tcfg.border_size / hs)
    rows, cols = int(tcfg.num_rows * length_px) + 2 * border, int(tcfg.num_cols * width_px) + 2 * border
    proportion0 = float(np.sum(tcfg.terrain_proportions[:1]))
    tiles = (_lib.ParkourTile * (tcfg.num_rows * tcfg.num_cols))()
    origins = np.zeros((tcfg.num_rows, tcfg.num_cols, 3))
    for j in range(tcfg.num_cols):
        for i in range(tcfg.num_rows):
            kw = _curriculum_tile_kwargs(j / tcfg.num_cols + 0.001, (i + 1) / 10, proportion0) if tcfg.curriculum else tcfg.parkour_kwargs
            tiles[i * tcfg.num_cols + j] = _tile_table(length_px, width_px, hs, vs, **kw)
            origins[i, j] = [i * tcfg.terrain_length, (j + 0.5) * tcfg.terrain_width, 0.0]
    raw = torch.frombuffer(bytearray(bytes(tiles)), dtype=torch.uint8).to(device)
    field = torch.empty(rows, cols, dtype=torch.int16, device=device)
    lib = _lib.lib()
    _lib.check(lib.b200_parkour_field(C.c_void_p(field.data_ptr()), rows, cols, border, length_px, width_px, tcfg.num_rows, tcfg.num_cols,
                                      C.c_void_p(raw.data_ptr()), _lib.stream_ptr()))
    torch.cuda.current_stream().synchronize()          # `raw` may go away
    return field, origins.astype(np.float32)


def heightfield_to_trimesh_gpu(height_field, horizontal_scale, vertical_scale, slope_threshold=None):
    """heightfield_to_trimesh on the device: int16 CUDA tensor [rows, cols] -> (vertices float32 [rows*cols, 3], triangles
    int32-typed tensor holding uint32 indices [2*(rows-1)*(cols-1), 3]), both CUDA tensors"""
    import ctypes as C

    import torch

    from . import _lib
    hf = height_field.contiguous()
    assert hf.is_cuda and hf.dtype == torch.int16 and hf.dim() == 2
    rows, cols = hf.shape
    vertices = torch.empty(rows * cols, 3, dtype=torch.float32, device=hf.device)
    triangles = torch.empty(2 * (rows - 1) * (cols - 1), 3, dtype=torch.int32, device=hf.device)
    _lib.check(_lib.lib().b200_heightfield_to_trimesh(C.c_void_p(hf.data_ptr()), rows, cols, float(horizontal_scale), float(vertical_scale),
                                                     int(slope_threshold is not None), float(slope_threshold or 0.0),
                                                     C.c_void_p(vertices.data_ptr()), C.c_void_p(triangles.data_ptr()), _lib.stream_ptr()))
    return vertices, triangles
