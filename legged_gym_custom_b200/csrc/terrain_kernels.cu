// Terrain construction on the device (init-time, not on the step path):
//   parkour_field_kernel   the parkour height field (terrain.py:103-131 parkour_curriculum / parkour_selected_terrain over
//                          terrain_utils.py:318-399 parkour_terrain), one thread per cell, from per-tile obstacle tables
//   trimesh_kernel         height field -> vertices / triangles with the slope-threshold correction
//                          (terrain_utils.py:401-465 convert_heightfield_to_trimesh), one thread per vertex / cell
// Both are bit-identical to the host generators in legged_gym_custom_b200/terrain.py (which are pinned to the reference).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
parkour_field_kernel(int16_t* __restrict__ field, int rows, int cols, int border, int length_px, int width_px, int tile_rows, int tile_cols,
                     const B200ParkourTile* __restrict__ tiles) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (int64_t)rows * cols) return;
  const int R = (int)(idx / cols), Cc = (int)(idx % cols);
  int16_t v = 0;
  const int r_in = R - border, c_in = Cc - border;
  if (r_in >= 0 && c_in >= 0 && r_in < tile_rows * length_px && c_in < tile_cols * width_px) {
    const int ti = r_in / length_px, tj = c_in / width_px;
    const int r = r_in - ti * length_px, c = c_in - tj * width_px;
    const B200ParkourTile& t = tiles[ti * tile_cols + tj];
    if (r < t.platform_rows) v = t.platform_height;
    for (int k = 0; k < t.num_obstacles; ++k) {          // later obstacles overwrite earlier ones, as the reference's loop does
      if (r >= t.row_lo[k] && r < t.row_hi[k]) v = (c < t.zero_below[k] || c >= t.zero_from[k]) ? (int16_t)0 : t.height[k];
    }
    if (c < t.pad || c >= width_px - t.pad) v = t.border_height;
  }
  field[idx] = v;
}

// np.linspace(0, (n - 1) * h, n)[i] in float64: i * step, the last element forced to the stop value
__device__ __forceinline__ double linspace_at(int i, int n, double h) {
  const double stop = (double)(n - 1) * h;
  if (n == 1) return 0.0;
  return i == n - 1 ? stop : (double)i * (stop / (double)(n - 1));
}

__global__ void __launch_bounds__(256)
trimesh_kernel(const int16_t* __restrict__ hf, int rows, int cols, double hscale, double vscale, int use_threshold, double thr,
               float* __restrict__ vertices, uint32_t* __restrict__ triangles) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (int64_t)rows * cols) return;
  const int i = (int)(idx / cols), j = (int)(idx % cols);
  auto H = [&](int a, int b) { return hf[(int64_t)a * cols + b]; };
  double x = linspace_at(i, rows, hscale), y = linspace_at(j, cols, hscale);
  if (use_threshold) {
    // int16 differences like np.diff on an int16 array (wrap-around included), compared with the float64 threshold
    auto d16 = [](int16_t a, int16_t b) { return (double)(int16_t)(a - b); };
    double mx = 0.0, my = 0.0, md = 0.0;
    if (i < rows - 1) mx += d16(H(i + 1, j), H(i, j)) > thr;
    if (i > 0) mx -= -d16(H(i, j), H(i - 1, j)) > thr;         // (the negation of an int16 diff: numpy negates in int16)
    if (j < cols - 1) my += d16(H(i, j + 1), H(i, j)) > thr;
    if (j > 0) my -= -d16(H(i, j), H(i, j - 1)) > thr;
    if (i < rows - 1 && j < cols - 1) md += d16(H(i + 1, j + 1), H(i, j)) > thr;
    if (i > 0 && j > 0) md -= -d16(H(i, j), H(i - 1, j - 1)) > thr;
    x = x + (mx + md * (mx == 0.0 ? 1.0 : 0.0)) * hscale;
    y = y + (my + md * (my == 0.0 ? 1.0 : 0.0)) * hscale;
  }
  vertices[idx * 3 + 0] = (float)x;
  vertices[idx * 3 + 1] = (float)y;
  vertices[idx * 3 + 2] = (float)((double)H(i, j) * vscale);
  if (i < rows - 1 && j < cols - 1) {
    const uint32_t v00 = (uint32_t)i * (uint32_t)cols + (uint32_t)j, v01 = v00 + 1u, v10 = v00 + (uint32_t)cols, v11 = v10 + 1u;
    uint32_t* t = triangles + ((int64_t)i * (cols - 1) + j) * 6;
    t[0] = v00; t[1] = v11; t[2] = v01;
    t[3] = v00; t[4] = v10; t[5] = v11;
  }
}

}  // namespace

extern "C" {

int b200_parkour_field(int16_t* field, int rows, int cols, int border, int length_px, int width_px, int tile_rows, int tile_cols,
                       const B200ParkourTile* tiles_dev, void* stream) {
  B200_CHECK_ARG(field && tiles_dev && rows > 0 && cols > 0 && length_px > 0 && width_px > 0 && tile_rows > 0 && tile_cols > 0 && border >= 0,
                 "b200_parkour_field: bad argument");
  B200_CHECK_ARG(rows == tile_rows * length_px + 2 * border && cols == tile_cols * width_px + 2 * border, "b200_parkour_field: shape mismatch");
  const int64_t n = (int64_t)rows * cols;
  parkour_field_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(field, rows, cols, border, length_px, width_px, tile_rows,
                                                                                       tile_cols, tiles_dev);
  B200_CHECK_LAUNCH("parkour_field_kernel");
  return 0;
}

int b200_heightfield_to_trimesh(const int16_t* height_field, int rows, int cols, double horizontal_scale, double vertical_scale,
                                int use_slope_threshold, double slope_threshold, float* vertices, uint32_t* triangles, void* stream) {
  B200_CHECK_ARG(height_field && vertices && triangles && rows >= 2 && cols >= 2, "b200_heightfield_to_trimesh: bad argument");
  B200_CHECK_ARG((int64_t)rows * cols < (1ll << 32), "b200_heightfield_to_trimesh: more than 2^32 vertices");
  const int64_t n = (int64_t)rows * cols;
  const double thr = slope_threshold * horizontal_scale / vertical_scale;
  trimesh_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(height_field, rows, cols, horizontal_scale, vertical_scale,
                                                                                 use_slope_threshold, thr, vertices, triangles);
  B200_CHECK_LAUNCH("trimesh_kernel");
  return 0;
}

}  // extern "C"
