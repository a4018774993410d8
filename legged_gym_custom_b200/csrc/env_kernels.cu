// Env-side kernels of libb200gym.so (compiled with -fmad=false, see env_core.cuh).
//
//   K1 pd_torques_kernel        one thread per (env, dof)            legged_robot.py:74-75, :440-478
//   K2 post_physics_kernel      CTA = 8 envs x 8 warps, five phases (see the kernel)   go2.py:345-387 and callees
//   K3 extras_kernel            one CTA per reward term (+1)         go2.py:246-263 (episode means, time_outs)
//      reset_all_kernel         one thread per env                   base_task.py:131-133
//      heights_kernel           one thread per scan point            legged_robot.py:997-1032
//
// All are HBM-bound streaming kernels: rows are read/written with consecutive lanes on
// consecutive addresses (16-byte vectors for the bulk rows), the per-env working set lives in
// shared memory, and the params/pointer structs travel as __grid_constant__ kernel arguments
// (constant bank, no extra copy per launch).
#include <cuda.h>
#include <stdarg.h>

#include "bulk_copy.cuh"
#include "common.cuh"
#include "env_core.cuh"

static thread_local char g_err[512] = "";
void b200_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

constexpr int kEnvsPerCta = 8;      // one warp per env for the row-parallel stages
constexpr int kHistThreads = (kEnvsPerCta - 1) * 32;   // warps 1..7 move the history rows
constexpr int kHistPerThread = 5;   // 8 envs x 130 vectors = 1040 <= 5 x 224
static_assert(B200_TERM_PARTS == kEnvsPerCta, "one reward-term part per warp");
static_assert(kEnvsPerCta * (B200_NUM_BODIES + B200_NUM_DOF + 1) <= kEnvsPerCta * 32, "item pass 1 must fit the CTA");

// `step_dev` != nullptr: the step counter lives in device memory (CUDA-graph replay) and `step` is an offset to it (1: the
// counter is advanced AFTER the step's kernels, by extras_kernel); else `step` is the counter.
//
// Phases of a CTA (8 envs, 8 warps; every phase ends with __syncthreads):
//   A   warp w: load env w's small rows, height scan (stage 0 / 1 of env_core.cuh)
//   E   the items of all 8 envs packed type by type: pass 1 = 152 bodies | 96 dofs | 8 flags over the 256 threads;
//       pass 2 = warp 0: 32 legs, warp 1: 32 angles (+ command update), warp 2: velocities, warp 3: feet + push (8 lanes)
//   B1  warp p, lane s < 8: part p of the reward terms of env slot s; warp w, 7 lanes: the Philox blocks of env w's reset
//       (if it resets); warps 1-7 then LOAD the 8 history rows
//   B2  warp 0, lane s < 8: reward sum, termination reward, reset of env slot s;
//       warps 1-7 meanwhile STORE the history rows (history -> clip -> obs / critic, history shifted in place)
//   C   warp w: current observation + critic tail of env w, write-back
template <bool FIXED>
__global__ void __launch_bounds__(kEnvsPerCta * 32, 4)
post_physics_kernel(const __grid_constant__ B200EnvParams P, const __grid_constant__ B200EnvBuffers B, int64_t step,
                    const int64_t* __restrict__ step_dev, unsigned long long* __restrict__ trace, int prefetch, int dry) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  EnvScratch* scratch = reinterpret_cast<EnvScratch*>(smem_raw);
  __shared__ float pt_x[B200_MAX_SCAN], pt_y[B200_MAX_SCAN];
  __shared__ EnvTables T;
  if (threadIdx.x < P.num_scan) scan_point(P, threadIdx.x, &pt_x[threadIdx.x], &pt_y[threadIdx.x]);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + B200_MAX_PROPRIO) env_tables_fill(P, T, threadIdx.x - 64);
  __syncthreads();
  if (step_dev) step += *step_dev;
  // `dry`: the probe pass of a command curriculum (launched only when P.command_curriculum): an ordinary step ends here
  if (dry && step % P.max_episode_length != 0) return;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  // optional phase trace (b200_env_set_phase_trace): %globaltimer of warp 0 at the phase boundaries, [cta][8]
#define B200_TRACE(slot)                                                                          \
  if (trace && t == 0) {                                                                          \
    unsigned long long t_;                                                                        \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                        \
    trace[(size_t)blockIdx.x * 8 + (slot)] = t_;                                                  \
  }
  B200_TRACE(0)
  const int e0 = blockIdx.x * kEnvsPerCta;
  const int e = e0 + warp;
  const bool live = e < P.num_envs;
  const int n_live = min(kEnvsPerCta, P.num_envs - e0);
  if (prefetch) {
    // b200_env_set_prefetch: the CTA's history rows (one contiguous block, first read in B1 behind two dependent round
    // trips) start towards L2 now; pays only when the rows are not L2-resident (env counts beyond ~32 k, cold benchmarks)
    const int hn = (FIXED ? B200_GO2_HISTORY : P.history_len) * B200_PROPRIO;
    const char* hp = reinterpret_cast<const char*>(B.obs_history_buf + (int64_t)e0 * hn);
    const int bytes = n_live * hn * 4;
    for (int off = t * 128; off < bytes; off += kEnvsPerCta * 32 * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(__cvta_generic_to_global(hp + off)));
  }

  // ---- A
  if (live) env_warp_pre<FIXED>(P, B, scratch[warp], pt_x, pt_y, e, lane, lane + 1);
  B200_TRACE(1)
  __syncthreads();

  // ---- E
  {
    constexpr int kBodies = kEnvsPerCta * B200_NUM_BODIES, kDofs = kEnvsPerCta * B200_NUM_DOF;
    if (t < kBodies) {
      const int slot = t / B200_NUM_BODIES;
      if (slot < n_live) env_item_body(P, scratch[slot], t - slot * B200_NUM_BODIES);
    } else if (t < kBodies + kDofs) {
      const int slot = (t - kBodies) / B200_NUM_DOF;
      if (slot < n_live) env_item_dof(P, T, scratch[slot], (t - kBodies) - slot * B200_NUM_DOF);
    } else if (t - (kBodies + kDofs) < n_live) {
      env_item_flags(P, scratch[t - (kBodies + kDofs)]);
    }
    const int slot4 = lane >> 2;
    if (warp == 0) {
      if (slot4 < n_live) env_item_leg(P, scratch[slot4], lane & 3);
    } else if (warp == 1) {
      if (slot4 < n_live) env_item_angle(P, scratch[slot4], lane & 3, (uint32_t)(e0 + slot4), (uint32_t)step);
    } else if (warp == 2) {
      if (lane < n_live) env_item_velocities(scratch[lane]);
    } else if (warp == 3) {
      if (lane < n_live) env_item_feet_push(P, scratch[lane], (uint32_t)(e0 + lane), step);
    }
  }
  B200_TRACE(2)
  __syncthreads();
  if (dry) {      // go2.py:222-223: (episode sum of tracking_lin_vel, reset flag) of every env, nothing else is written
    if (warp == 0 && lane < n_live) env_cc_probe<FIXED>(P, B, scratch[lane], e0 + lane);
    return;
  }

  // ---- B1 / B2
  const int hn4 = (FIXED ? B200_GO2_HISTORY : P.history_len) * (B200_PROPRIO / 4);
  const int total = kEnvsPerCta * hn4;
  const int ht = t - 32;
  if (lane < n_live) env_terms_part<FIXED>(P, scratch[lane], warp);
  if (live && scratch[warp].early_reset && lane < B200_RESET_BLOCKS)       // rare: the 7 Philox blocks of a reset, in parallel
    env_reset_draw(P, scratch[warp].reset_draws, (uint32_t)e, (uint32_t)step, lane);
  for (int base = 0; base < total; base += kHistThreads * kHistPerThread) {      // one trip for the go2 layout
    f4_ v[kHistPerThread];
    if (warp > 0) {
#pragma unroll
      for (int j = 0; j < kHistPerThread; ++j) {
        const int k = base + ht + j * kHistThreads;
        const int slot = k / hn4, i = k - slot * hn4;
        if (k < total && slot < n_live) v[j] = env_hist_load<FIXED>(P, B, e0 + slot, i);
      }
    }
    if (base == 0) { B200_TRACE(3) }
    __syncthreads();                                   // terms complete; every load of the rows before any store
    if (warp == 0) {
      if (base == 0 && lane < n_live) env_finalize(P, B, scratch[lane]);
    } else {
#pragma unroll
      for (int j = 0; j < kHistPerThread; ++j) {
        const int k = base + ht + j * kHistThreads;
        const int slot = k / hn4, i = k - slot * hn4;
        if (k < total && slot < n_live)
          env_hist_store<FIXED>(P, B, e0 + slot, i, v[j], scratch[slot].early_reset, scratch[slot].early_refill);
      }
    }
  }
  B200_TRACE(4)
  __syncthreads();

  // ---- C
  if (live) env_warp_post<FIXED>(P, B, T, scratch[warp], e, step, lane, lane + 1);
  B200_TRACE(5)
#undef B200_TRACE
}

// ---- post_physics_tile_kernel: the go2 layout with `alias_outputs` (the observation outputs exist once, as rows of
// critic_obs_buf = [history 520 | cur 52 | priv 29 | est 3 | scan 132]).  Same pieces, same phases and barriers as
// post_physics_kernel; what changes is how the rows move:
//   * every env's output row is assembled in a shared-memory TILE [8][736] and leaves with ONE bulk copy (cp.async.bulk,
//     2944 B) issued by its warp -- no per-lane 16-byte stores, no second copy of the 572 observation values;
//   * the history rows never touch registers: a bulk copy at kernel entry drops env's [10 x 52] history at the head of its
//     tile row (it arrives under phases A / E), the shifted history goes back to global memory as a bulk copy of
//     tile[52:520] as soon as the termination flags are known (B1), and the new proprioceptive row is written behind it
//     (tile[520:572]) -- so tile[0:572] IS the unclipped observation; one in-place clip pass turns it into the output;
//   * the reward sum / reset of the 8 envs (one warp, B2) overlaps the cur / tail rows of the envs that do not reset.
typedef EnvScratchT<4, 4, B200_GO2_SCAN_NX * B200_GO2_SCAN_NY> TileScratch;      // cur / tail live in the tile
static_assert((sizeof(TileScratch) / 16) % 2 == 1, "TileScratch stride would bank-conflict the per-env lanes");
constexpr int kTileHist = B200_GO2_HISTORY * B200_PROPRIO;                       // 520
constexpr int kTileObs = kTileHist + B200_PROPRIO;                               // 572
constexpr int kTileRow = kTileObs + 32 + B200_GO2_SCAN_NX * B200_GO2_SCAN_NY;     // 736
constexpr size_t kTileSmem = (size_t)kEnvsPerCta * (kTileRow * sizeof(float) + sizeof(TileScratch)) + 128;
static_assert((kTileRow * sizeof(float)) % 16 == 0 && (kTileHist * sizeof(float)) % 16 == 0, "rows move as 16-byte multiples");

// cur (unclipped, go2.py:506-519) and the critic tail's est / scan parts of env `e` into its tile row; one warp
__device__ __forceinline__ void tile_cur_tail(const B200EnvParams& P, const EnvTables& T, const TileScratch& S, float* row, int e,
                                              int64_t step64, int lane) {
  constexpr int NP = B200_PROPRIO, NS = B200_GO2_SCAN_NX * B200_GO2_SCAN_NY, NPRIV = 29;
  const float c = P.clip_obs;
  Philox4 r;
  if (P.add_noise) r = keyed_block(P.seed, SITE_OBS_NOISE, (uint32_t)step64, (uint32_t)e, (uint32_t)(lane & 15));
#pragma unroll
  for (int i = lane; i < NP; i += 32) {
    float u = 0.0f;
    if (P.add_noise) {
      const uint32_t w = (uint32_t)(i >> 4);
      u = u32_to_uniform(w == 0 ? r.v[0] : (w == 1 ? r.v[1] : (w == 2 ? r.v[2] : r.v[3])));
    }
    row[kTileHist + i] = cur_obs_element(P, T, S, u, i);
  }
  if (lane >= NPRIV) row[kTileObs + lane] = clampf(S.blv[lane - NPRIV] * P.obs_lin_vel, -c, c);      // priv part: stage 0
#pragma unroll
  for (int j = lane; j < NS; j += 32) row[kTileObs + 32 + j] = clampf((S.root_out[2] - 0.3f) - S.heights[j], -1.0f, 1.0f);
}

__global__ void __launch_bounds__(kEnvsPerCta * 32, 4)
post_physics_tile_kernel(const __grid_constant__ B200EnvParams P, const __grid_constant__ B200EnvBuffers B, int64_t step,
                         const int64_t* __restrict__ step_dev, unsigned long long* __restrict__ trace,
                         const __grid_constant__ CUtensorMap hs_map, int terrain_tiles) {
  extern __shared__ __align__(128) uint8_t tile_smem_raw[];
  // tile rows start on 128-byte boundaries of the shared window (TMA destinations): align by hand, the launch adds 128 bytes
  uint8_t* tile_base = tile_smem_raw + ((128u - (bulk::smem_addr(tile_smem_raw) & 127u)) & 127u);
  float* tile = reinterpret_cast<float*>(tile_base);
  TileScratch* scratch = reinterpret_cast<TileScratch*>(tile_base + (size_t)kEnvsPerCta * kTileRow * sizeof(float));
  __shared__ float pt_x[B200_MAX_SCAN], pt_y[B200_MAX_SCAN];
  __shared__ EnvTables T;
  __shared__ __align__(8) uint64_t hist_bar;
  __shared__ __align__(8) uint64_t terrain_bar[kEnvsPerCta];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int e0 = blockIdx.x * kEnvsPerCta;
  const int e = e0 + warp;
  const bool live = e < P.num_envs;
  const int n_live = min(kEnvsPerCta, P.num_envs - e0);
  constexpr int NP = B200_PROPRIO, NP4 = NP / 4, HN4 = kTileHist / 4, OBS4 = kTileObs / 4, NS4 = B200_GO2_SCAN_NX * B200_GO2_SCAN_NY / 4;
  if (t == 0) {      // the history rows start towards shared memory before anything else happens ...
    // (... unless the head of every tile row first hosts the env's TERRAIN tile: then warp w fetches its history itself,
    // right after its height scan -- n_live arrivals instead of one)
    bulk::mbar_init(&hist_bar, terrain_tiles ? (uint32_t)n_live : 1u);
    if (terrain_tiles) {
      for (int w = 0; w < kEnvsPerCta; ++w) bulk::mbar_init(terrain_bar + w, 1);
    } else {
      bulk::mbar_expect_tx(&hist_bar, (uint32_t)(n_live * kTileHist * sizeof(float)));
      for (int slot = 0; slot < n_live; ++slot)
        bulk::load(tile + slot * kTileRow, B.obs_history_buf + (int64_t)(e0 + slot) * kTileHist, kTileHist * sizeof(float), &hist_bar);
    }
  }
  if (t < P.num_scan) scan_point(P, t, &pt_x[t], &pt_y[t]);
  if (t >= 64 && t < 64 + B200_MAX_PROPRIO) env_tables_fill<TileScratch>(P, T, t - 64);
  if (step_dev) step += *step_dev;
#define B200_TRACE(slot)                                                                          \
  if (trace && t == 0) {                                                                          \
    unsigned long long t_;                                                                        \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                        \
    trace[(size_t)blockIdx.x * 8 + (slot)] = t_;                                                  \
  }
  B200_TRACE(0)
  float* row = tile + warp * kTileRow;

  // ---- A0: small rows -> scratch (the privileged statics straight into the tile's tail).  Needs neither the tables nor
  //      the scan points, so the barrier behind it is also the one that publishes them (no barrier of its own for those).
  if (live) env_warp_pre<true>(P, B, scratch[warp], pt_x, pt_y, e, lane, lane + 1, row + kTileObs, /*do_scan=*/false);
  B200_TRACE(1)
  __syncthreads();

  // ---- A1 + E: the items of all 8 envs packed type by type (as post_physics_kernel; the termination flags, which need the
  //      scan's outlier counts, moved to B1) and then every warp's height scan: both read only what A0 staged, so the
  //      gathers of one warp fly under the item arithmetic of the others
  {
    constexpr int kBodies = kEnvsPerCta * B200_NUM_BODIES, kDofs = kEnvsPerCta * B200_NUM_DOF;
    if (t < kBodies) {
      const int slot = t / B200_NUM_BODIES;
      if (slot < n_live) env_item_body(P, scratch[slot], t - slot * B200_NUM_BODIES);
    } else if (t < kBodies + kDofs) {
      const int slot = (t - kBodies) / B200_NUM_DOF;
      if (slot < n_live) env_item_dof(P, T, scratch[slot], (t - kBodies) - slot * B200_NUM_DOF);
    }
    const int slot4 = lane >> 2;
    if (warp == 0) {
      if (slot4 < n_live) env_item_leg(P, scratch[slot4], lane & 3);
    } else if (warp == 1) {
      if (slot4 < n_live) env_item_angle(P, scratch[slot4], lane & 3, (uint32_t)(e0 + slot4), (uint32_t)step);
    } else if (warp == 2) {
      if (lane < n_live) env_item_velocities(scratch[lane]);
    } else if (warp == 3) {
      if (lane < n_live) env_item_feet_push(P, scratch[lane], (uint32_t)(e0 + lane), step);
    }
  }
  if (live && !terrain_tiles) env_warp_scan<true>(P, B, scratch[warp], pt_x, pt_y, e, lane, lane + 1);
  if (live && terrain_tiles) {
    // Height scan from a shared-memory terrain tile (north_star design choice 2; legged_robot.py:997-1032).  The 132 scan
    // points of an env cover at most ~23 x 23 cells of the field (1.65 m x 1.5 m rotated by any yaw, 0.1 m cells): the lanes
    // compute their cells, the warp reduces the bounding box, lane 0 fetches the 32 x 32 int16 box at its corner with ONE
    // 2-D TMA copy into the (still unused) head of the env's tile row, and the 3 cells of every point are read from there.
    // A box wider than 31 cells (a layout this kernel does not run) falls back to the gathers.
    TileScratch& S = scratch[warp];
    constexpr int NS = B200_GO2_SCAN_NX * B200_GO2_SCAN_NY, PER = (NS + 31) / 32;
    const YawQuat yq = yaw_quat(S.root + 3);
    const float inv_h = 1.0f / P.horizontal_scale;
    int px[PER], py[PER];
    int lo_x = 0x7fffffff, lo_y = 0x7fffffff, hi_x = 0, hi_y = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int j = lane + 32 * k, jc = j < NS ? j : NS - 1;
      height_cell_pt(P, pt_x[jc], pt_y[jc], yq, S.root, inv_h, &px[k], &py[k]);
      lo_x = min(lo_x, px[k]); hi_x = max(hi_x, px[k]);
      lo_y = min(lo_y, py[k]); hi_y = max(hi_y, py[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo_x = min(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o)); hi_x = max(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
      lo_y = min(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o)); hi_y = max(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
    }
    // a TMA box must start on a 16-byte boundary of its innermost dimension (8 int16 cells): measured -- an odd column start
    // raises "illegal instruction" (tools/scratch/tma_u16_test.cu) -- so the box starts at the column rounded down to 8
    lo_y &= ~7;
    const bool fits = hi_x - lo_x <= 30 && hi_y - lo_y <= 30;       // + the (px + 1, py) / (px, py + 1) neighbours
    const int16_t* tl = reinterpret_cast<const int16_t*>(row);       // [32 x][32 y] int16
    if (fits) {
      if (lane == 0) {
        bulk::mbar_expect_tx(terrain_bar + warp, 32 * 32 * 2);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                         bulk::smem_addr(row)),
                     "l"(&hs_map), "r"(bulk::smem_addr(terrain_bar + warp)), "r"(lo_y), "r"(lo_x)
                     : "memory");
      }
      bulk::mbar_wait(terrain_bar + warp, 0);
    }
    int n_out = 0;
    const int pitch = hs_pitch_of(P);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int j = lane + 32 * k;
      if (j < NS) {
        int16_t a, b, c;
        if (fits) {
          const int o = (px[k] - lo_x) * 32 + (py[k] - lo_y);
          a = tl[o]; b = tl[o + 32]; c = tl[o + 1];
        } else {
          const int16_t* p = B.height_samples + (px[k] * pitch + py[k]);
          a = __ldg(p); b = __ldg(p + pitch); c = __ldg(p + 1);
        }
        int16_t m = a < b ? a : b;
        m = m < c ? m : c;
        const float h = (float)m * P.vertical_scale;
        S.heights[j] = h;
        n_out += fabsf(h) > 0.1f;
      }
    }
    S.outliers[lane] = n_out;
    __syncwarp();                                                    // every lane is done with the terrain tile: the history may land
    if (lane == 0) {
      bulk::mbar_expect_tx(&hist_bar, (uint32_t)(kTileHist * sizeof(float)));
      bulk::load(row, B.obs_history_buf + (int64_t)e * kTileHist, kTileHist * sizeof(float), &hist_bar);
    }
  }
  B200_TRACE(2)
  __syncthreads();

  // ---- B1: lane 16 of warp w: termination flags of env w (check_termination + next step's jump flags); then reward terms
  //      (warp = part: a compile-time constant per case, so every warp runs straight-line code for ITS terms only; lane = env
  //      slot), the Philox blocks of a reset, and lane 16 sends env w's shifted history home (in place in global memory: the
  //      source is the shared-memory copy, which has fully arrived)
  if (live && lane == 16) env_item_flags(P, scratch[warp]);
  __syncwarp();
  if (lane < n_live) {
    switch (warp) {
      case 0: env_terms_part<true>(P, scratch[lane], 0); break;
      case 1: env_terms_part<true>(P, scratch[lane], 1); break;
      case 2: env_terms_part<true>(P, scratch[lane], 2); break;
      case 3: env_terms_part<true>(P, scratch[lane], 3); break;
      case 4: env_terms_part<true>(P, scratch[lane], 4); break;
      case 5: env_terms_part<true>(P, scratch[lane], 5); break;
      case 6: env_terms_part<true>(P, scratch[lane], 6); break;
      default: env_terms_part<true>(P, scratch[lane], 7); break;
    }
  }
  if (live && scratch[warp].early_reset && lane >= 8 && lane < 8 + B200_RESET_BLOCKS)
    env_reset_draw(P, scratch[warp].reset_draws, (uint32_t)e, (uint32_t)step, lane - 8);
  const bool shift = live && !scratch[warp].early_refill;
  if (lane == 16 && shift) {
    bulk::mbar_wait(&hist_bar, 0);
    bulk::store(B.obs_history_buf + (int64_t)e * kTileHist, row + NP, (kTileHist - NP) * sizeof(float));
    bulk::commit();
  }
  B200_TRACE(3)
  __syncthreads();

  // ---- B2: warp 0 = reward sum, termination reward, reset of the 8 envs; the other warps meanwhile write the cur / tail
  //      rows of their envs unless the env resets (finalize rewrites what those rows read only for a resetting env)
  if (warp == 0) {
    if (lane < n_live) env_finalize(P, B, scratch[lane]);
  } else if (live && !scratch[warp].early_reset) {
    tile_cur_tail(P, T, scratch[warp], row, e, step, lane);
  }
  B200_TRACE(4)
  __syncthreads();

  // ---- C: remaining cur / tail rows, clip in place, write-back
  if (!live) return;
  TileScratch& S = scratch[warp];
  if (warp == 0 || S.early_reset) tile_cur_tail(P, T, S, row, e, step, lane);
  bulk::mbar_wait(&hist_bar, 0);                        // every reader of the tile observes the arrival itself
  if (lane == 16 && shift) bulk::wait_read();           // the shift store has read tile[52:520]: the row may now change
  __syncwarp();
  {
    const float c = P.clip_obs;
    const bool reset = S.reset != 0, refill = S.ep_len_out <= 1;      // refill == S.early_refill (go2.py:570-574)
    f4_* row4 = reinterpret_cast<f4_*>(row);
    f4_* hist4 = reinterpret_cast<f4_*>(B.obs_history_buf + (int64_t)e * kTileHist);
#pragma unroll
    for (int i = lane; i < OBS4; i += 32) {
      const f4_ v = row4[i];
      if (i >= HN4) {                                   // the new row: history's newest slot, or all of them after a reset
        if (!refill) hist4[i - NP4] = v;
        else
#pragma unroll
          for (int k = 0; k < B200_GO2_HISTORY; ++k) hist4[k * NP4 + (i - HN4)] = v;
      }
      f4_ o = clamp4(v, c);
      if (reset && i < HN4) o.x = o.y = o.z = o.w = 0.0f;      // obs_history_buf[env_ids] = 0 before the observation (go2.py:238)
      row4[i] = o;
    }
    bulk::fence_smem_writes();
    __syncwarp();
    if (lane == 0) {
      bulk::store(B.critic_obs_buf + (int64_t)e * kTileRow, row, kTileRow * sizeof(float));
      bulk::commit();
    }
    const f4_* h4 = reinterpret_cast<const f4_*>(S.heights);
    for (int i = lane; i < NS4; i += 32) reinterpret_cast<f4_*>(B.measured_heights + (int64_t)e * (NS4 * 4))[i] = h4[i];
    // persistent state (go2.py:380-384 and the in-place updates of reset_idx) -- as env_warp_post
    if (lane < 12) {
      B.last_actions[(int64_t)e * 12 + lane] = S.act[lane];
      B.last_dof_vel[(int64_t)e * 12 + lane] = S.dof_out[2 * lane + 1];
      B.last_torques[(int64_t)e * 12 + lane] = S.tq[lane];
    }
    if (lane < 6) B.last_root_vel[(int64_t)e * 6 + lane] = S.root_out[7 + lane];
    if (lane < 3) {
      B.last_base_lin_vel[(int64_t)e * 3 + lane] = S.blv[lane];
      B.base_lin_vel[(int64_t)e * 3 + lane] = S.blv[lane];
      B.base_ang_vel[(int64_t)e * 3 + lane] = S.bav[lane];
      B.projected_gravity[(int64_t)e * 3 + lane] = S.pg[lane];
      B.rpy[(int64_t)e * 3 + lane] = S.rpy[lane];
      B.env_origins[(int64_t)e * 3 + lane] = S.origin_out[lane];
    }
    if (lane < 5) B.phases[(int64_t)e * 5 + lane] = S.phases[lane];
    if (lane < 4) {
      B.commands[(int64_t)e * 4 + lane] = S.cmd_out[lane];
      B.last_contact_heights[(int64_t)e * 4 + lane] = S.lch_out[lane];
      B.feet_air_time[(int64_t)e * 4 + lane] = S.fat_out[lane];
      B.last_contacts[(int64_t)e * 4 + lane] = (uint8_t)S.contact_cur[lane];
      B.foot_contacts[(int64_t)e * 4 + lane] = (uint8_t)S.contact_filt[lane];
    }
    if (S.root_dirty && lane < 13) B.root_states[(int64_t)e * 13 + lane] = S.root_out[lane];
    if (S.dof_dirty && lane < 24) B.dof_state[(int64_t)e * 24 + lane] = S.dof_out[lane];
    for (int k = lane; k < B200_NUM_REWARD_TERMS; k += 32) {
      const float total = S.sums[k] + S.term[k];
      B.episode_sums[(int64_t)e * B200_NUM_REWARD_TERMS + k] = S.reset ? 0.0f : total;
      if (S.reset) B.reset_episode_sums[(int64_t)e * B200_NUM_REWARD_TERMS + k] = total;
    }
    if (lane == 0) {
      B.episode_length_buf[e] = S.ep_len_out;
      B.terrain_levels[e] = S.level_out;
      B.jump_flags[e] = S.jump_flag_out;
      B.rew_buf[e] = S.rew;
      B.reset_buf[e] = (uint8_t)S.reset;
      B.time_out_buf[e] = (uint8_t)S.time_out;
      bulk::wait_read();                                // the row store has read shared memory: the CTA may retire
    }
  }
  B200_TRACE(5)
#undef B200_TRACE
}

__global__ void __launch_bounds__(256)
pd_torques_kernel(const __grid_constant__ B200EnvParams P, const __grid_constant__ B200EnvBuffers B,
                  const float* __restrict__ actions_in, int clip_and_store) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)P.num_envs * B200_NUM_DOF) return;
  pd_torque_element(P, B, actions_in, clip_and_store, idx);
}

__global__ void __launch_bounds__(128)
reset_all_kernel(const __grid_constant__ B200EnvParams P, const __grid_constant__ B200EnvBuffers B, int64_t step, int init_done) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P.num_envs) return;
  env_reset_only(P, B, e, step, init_done);
}

__global__ void __launch_bounds__(128)
env_init_kernel(const __grid_constant__ B200EnvParams P, const __grid_constant__ B200EnvBuffers B, const __grid_constant__ B200InitParams I) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < P.num_envs) env_init_one(P, B, I, e);
}

__global__ void __launch_bounds__(256)
heights_kernel(const __grid_constant__ B200EnvParams P, const __grid_constant__ B200EnvBuffers B) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)P.num_envs * P.num_scan) return;
  const int e = (int)(idx / P.num_scan), j = (int)(idx % P.num_scan);
  float h = 0.0f;
  int px = 0, py = 0;
  if (P.has_height_samples) {
    float root[7];
    for (int i = 0; i < 7; ++i) root[i] = B.root_states[(int64_t)e * 13 + i];
    float vx, vy;
    scan_point(P, j, &vx, &vy);
    height_cell_pt(P, vx, vy, yaw_quat(root + 3), root, 1.0f / P.horizontal_scale, &px, &py);
    h = height_at(P, B.height_samples, px, py);
  }
  B.measured_heights[idx] = h;
  if (B.height_index) {
    B.height_index[idx * 2] = px;
    B.height_index[idx * 2 + 1] = py;
  }
}

// Deterministic CTA-wide sums (fixed tree), used for the episode means.
template <typename T>
__device__ T cta_sum_256(T v, T* smem) {
  const int t = threadIdx.x;
  smem[t] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (t < s) smem[t] += smem[t + s];
    __syncthreads();
  }
  const T r = smem[0];
  __syncthreads();
  return r;
}

// extras["episode"]["rew_<term>"] = mean(episode_sums[term][reset ids]) / max_episode_length_s,
// extras["episode"]["terrain_level"] = mean(terrain_levels), extras["time_outs"] = time_out_buf --
// all only when at least one env reset this step (go2.py:214-215, :246-263).
__global__ void __launch_bounds__(256)
extras_kernel(const __grid_constant__ B200EnvParams P, const __grid_constant__ B200EnvBuffers B, int64_t* __restrict__ step_dev) {
  __shared__ float fsm[256];
  __shared__ int ism[256];
  const int k = blockIdx.x, T = B200_NUM_REWARD_TERMS, N = P.num_envs;
  // device-resident step counter (go2.py:355): the step's kernels ran with counter + 1, the last launch of the step commits it
  if (step_dev && k == 0 && threadIdx.x == 0) *step_dev += 1;
  int cnt = 0;
  float acc = 0.0f;
  // thread t owns envs [16 t, 16 t + 16) of every 4096-env chunk: the 16 reset flags arrive as one 16-byte load and only
  // the (rare) flagged envs touch reset_episode_sums; fixed order, so the means are deterministic
  for (int e0 = threadIdx.x * 16; e0 < N; e0 += 256 * 16) {
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (e0 + 16 <= N && (((uintptr_t)B.reset_buf) & 15) == 0) {
      const uint4 v = *reinterpret_cast<const uint4*>(B.reset_buf + e0);
      w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
      for (int i = 0; i < 16 && e0 + i < N; ++i) w[i >> 2] |= (uint32_t)(B.reset_buf[e0 + i] != 0) << (8 * (i & 3));
    }
    if (k < T) {
      if (w[0] | w[1] | w[2] | w[3]) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if ((w[i >> 2] >> (8 * (i & 3))) & 0xffu) {
            ++cnt;
            acc += B.reset_episode_sums[(int64_t)(e0 + i) * T + k];
          }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (e0 + i < N) {
          cnt += ((w[i >> 2] >> (8 * (i & 3))) & 0xffu) != 0;
          acc += (float)B.terrain_levels[e0 + i];
        }
      }
    }
  }
  const int count = cta_sum_256<int>(cnt, ism);
  const float sum = cta_sum_256<float>(acc, fsm);
  if (count == 0) {
    if (k == 0 && threadIdx.x == 0) B.reset_count[0] = 0;
    return;
  }
  if (threadIdx.x == 0) {
    if (k < T) {
      if (P.reward_scales[k] != 0.0f) B.extras_episode[k] = (sum / (float)count) / P.max_episode_length_s;
    } else if (P.curriculum) {
      B.extras_episode[T] = sum / (float)N;
    }
    if (k == 0) B.reset_count[0] = count;
  }
  const int per = (N + gridDim.x - 1) / gridDim.x;
  const int lo = k * per, hi = min(N, lo + per);
  for (int e = lo + threadIdx.x; e < hi; e += 256) B.extras_time_outs[e] = B.time_out_buf[e];
}

// ---- C ABI -----------------------------------------------------------------------------------
extern "C" {

const char* b200_last_error(void) { return g_err; }
int b200_abi_version(void) { return B200_ABI_VERSION; }
int b200_env_params_size(void) { return (int)sizeof(B200EnvParams); }
int b200_env_buffers_size(void) { return (int)sizeof(B200EnvBuffers); }

int b200_env_create(const B200EnvParams* p, int device, B200Env** out) {
  B200_CHECK_ARG(p && out, "b200_env_create: null argument");
  B200_CHECK_ARG(p->abi_version == B200_ABI_VERSION, "b200_env_create: abi_version %d != %d", p->abi_version, B200_ABI_VERSION);
  B200_CHECK_ARG(p->num_envs > 0, "b200_env_create: num_envs must be > 0");
  B200_CHECK_ARG(p->num_proprio == B200_PROPRIO, "b200_env_create: num_proprio %d unsupported (go2 layout is %d)", p->num_proprio, B200_PROPRIO);
  B200_CHECK_ARG(p->history_len > 0 && p->history_len * p->num_proprio <= B200_MAX_HIST, "b200_env_create: history too long");
  B200_CHECK_ARG(p->num_scan == p->scan_nx * p->scan_ny && p->num_scan <= B200_MAX_SCAN, "b200_env_create: bad scan grid");
  B200_CHECK_ARG(p->num_scan % 4 == 0, "b200_env_create: num_scan must be a multiple of 4 (rows move as 16-byte vectors)");
  B200_CHECK_ARG(p->num_priv == 29 && p->num_est == 3, "b200_env_create: privileged/estimated layout must be 29/3");
  B200_CHECK_ARG(p->n_penalised <= B200_NUM_BODIES && p->n_termination <= B200_NUM_BODIES, "b200_env_create: body tables");
  B200_CHECK_ARG(!p->has_height_samples || (p->hs_rows >= 2 && p->hs_cols >= 2 && (p->hs_pitch == 0 || p->hs_pitch >= p->hs_cols) &&
                                            (int64_t)p->hs_rows * (p->hs_pitch > 0 ? p->hs_pitch : p->hs_cols) < (1ll << 31)),
                 "b200_env_create: height_samples shape / pitch");
  B200_CHECK_ARG(!p->terrain_tiles || (p->has_height_samples && p->hs_pitch % 8 == 0 && p->hs_pitch > 0),
                 "b200_env_create: terrain_tiles needs a height field with hs_pitch % 8 == 0");
  B200_CHECK_ARG(!p->has_height_samples || (p->horizontal_scale > 0.0f && (float)(p->hs_rows + p->hs_cols) * p->horizontal_scale < 8388608.0f),
                 "b200_env_create: height field extent must stay below 2^23 m");
  B200_CHECK_ARG(p->resample_interval > 0 && p->push_interval > 0, "b200_env_create: intervals must be > 0");
  B200_CHECK_ARG(!p->command_curriculum || p->max_episode_length >= 2, "b200_env_create: command curriculum needs max_episode_length >= 2");
  B200_CHECK_ARG(p->alias_outputs == 0 || p->alias_outputs == 1, "b200_env_create: alias_outputs must be 0 or 1");
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess) {
    b200_set_error("b200_env_create: %s", cudaGetErrorString(err));
    return (int)err;
  }
  B200_CHECK_ARG(device >= 0 && device < ndev, "b200_env_create: device %d of %d", device, ndev);
  B200Env* env = new B200Env;
  env->p = *p;
  env->device = device;
  env->force_generic_layout = 0;
  env->phase_trace = nullptr;
  env->prefetch_history = 0;
  *out = env;
  return 0;
}

int b200_env_destroy(B200Env* env) {
  delete env;
  return 0;
}

int b200_env_set_phase_trace(B200Env* env, unsigned long long* trace) {
  B200_CHECK_ARG(env, "b200_env_set_phase_trace: null handle");
  env->phase_trace = trace;
  return 0;
}

int b200_env_force_generic_layout(B200Env* env, int on) {
  B200_CHECK_ARG(env, "b200_env_force_generic_layout: null handle");
  env->force_generic_layout = on ? 1 : 0;
  return 0;
}

int b200_env_set_prefetch(B200Env* env, int on) {
  B200_CHECK_ARG(env, "b200_env_set_prefetch: null handle");
  env->prefetch_history = on ? 1 : 0;
  return 0;
}

// 2-D tensor map of the height field for the terrain-tile mode: int16 [hs_rows][hs_pitch], box 32 x 32, no swizzle
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_hs_map(CUtensorMap* map, const B200EnvParams& p, const int16_t* hs) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      b200_set_error("cuTensorMapEncodeTiled is unavailable");
      return -2;
    }
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  cuuint64_t dims[2] = {(cuuint64_t)p.hs_cols, (cuuint64_t)p.hs_rows};
  cuuint64_t strides[1] = {(cuuint64_t)p.hs_pitch * 2};
  cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<int16_t*>(hs), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b200_set_error("cuTensorMapEncodeTiled(height_samples) failed with %d", (int)r);
    return -3;
  }
  return 0;
}

static int post_physics_attr() {
  static bool done = false;
  if (!done) {
    cudaError_t et = cudaFuncSetAttribute(post_physics_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmem);
    if (et == cudaSuccess) et = cudaFuncSetAttribute(post_physics_tile_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (et != cudaSuccess) {
      b200_set_error("post_physics_tile_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(et));
      return (int)et;
    }
    for (int fixed = 0; fixed < 2; ++fixed) {
      cudaError_t e = cudaFuncSetAttribute(fixed ? post_physics_kernel<true> : post_physics_kernel<false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kEnvsPerCta * sizeof(EnvScratch)));
      if (e != cudaSuccess) {
        b200_set_error("post_physics_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
      }
    }
    done = true;
  }
  return 0;
}

// the go2 layout (history 10, scan 12 x 11) runs the variant with the layout baked in
static int launch_post_physics(const B200Env* env, const B200EnvBuffers* bufs, int64_t step, const int64_t* step_dev, cudaStream_t st,
                                int dry = 0) {
  const int ctas = (env->p.num_envs + kEnvsPerCta - 1) / kEnvsPerCta;
  const size_t smem = kEnvsPerCta * sizeof(EnvScratch);
  if (env->p.alias_outputs && env_layout_is_go2(env->p) && !env->force_generic_layout && !dry && !bufs->height_index) {
    CUtensorMap hs_map = {};
    const int tiles = env->p.terrain_tiles && env->p.has_height_samples && ((uintptr_t)bufs->height_samples & 15) == 0;
    if (tiles)
      if (int rc = make_hs_map(&hs_map, env->p, bufs->height_samples)) return rc;
    post_physics_tile_kernel<<<ctas, kEnvsPerCta * 32, kTileSmem, st>>>(env->p, *bufs, step, step_dev, env->phase_trace, hs_map, tiles);
  }
  else if (env_layout_is_go2(env->p) && !env->force_generic_layout)
    post_physics_kernel<true><<<ctas, kEnvsPerCta * 32, smem, st>>>(env->p, *bufs, step, step_dev, dry ? nullptr : env->phase_trace, env->prefetch_history, dry);
  else
    post_physics_kernel<false><<<ctas, kEnvsPerCta * 32, smem, st>>>(env->p, *bufs, step, step_dev, dry ? nullptr : env->phase_trace,
                                                                     env->prefetch_history, dry);
  return 0;
}

// Command curriculum (go2.py:80-107, :222-223), one CTA per step when P.command_curriculum:
//   command_ranges[0:2] <- command_ranges[2:4]       (the move decided on the previous curriculum step is now in force)
//   on a step with common_step_counter % max_episode_length == 0:
//   command_ranges[2:4] <- rule(mean of cc_value over the envs with cc_reset)      (fixed-order sums: deterministic)
__global__ void __launch_bounds__(256)
command_curriculum_kernel(const __grid_constant__ B200EnvParams P, const __grid_constant__ B200EnvBuffers B, int64_t step,
                          const int64_t* __restrict__ step_dev) {
  __shared__ double dsm[256];
  __shared__ int ism[256];
  if (step_dev) step += *step_dev;
  double* cr = B.command_ranges;
  if (threadIdx.x == 0) {
    cr[0] = cr[2];
    cr[1] = cr[3];
  }
  if (step % P.max_episode_length != 0) return;
  int cnt = 0;
  double acc = 0.0;
  for (int e = threadIdx.x; e < P.num_envs; e += 256)
    if (B.cc_reset[e]) {
      ++cnt;
      acc += (double)B.cc_value[e];
    }
  const int count = cta_sum_256<int>(cnt, ism);
  const double sum = cta_sum_256<double>(acc, dsm);
  if (threadIdx.x == 0) command_curriculum_rule(P, count, sum, cr + 2, cr + 2);
}

static int launch_command_curriculum(const B200Env* env, const B200EnvBuffers* bufs, int64_t step, const int64_t* step_dev, cudaStream_t st) {
  if (int rc = launch_post_physics(env, bufs, step, step_dev, st, /*dry=*/1)) return rc;
  B200_CHECK_LAUNCH("post_physics_kernel (command-curriculum probe)");
  command_curriculum_kernel<<<1, 256, 0, st>>>(env->p, *bufs, step, step_dev);
  B200_CHECK_LAUNCH("command_curriculum_kernel");
  return 0;
}

static int check_bufs(const B200Env* env, const B200EnvBuffers* b, const char* who) {
  B200_CHECK_ARG(env && b, "%s: null argument", who);
  const void* const* ptrs = reinterpret_cast<const void* const*>(b);
  const int n = (int)(sizeof(B200EnvBuffers) / sizeof(void*));
  const int idx_height_samples = (int)(offsetof(B200EnvBuffers, height_samples) / sizeof(void*));
  const int idx_origins = (int)(offsetof(B200EnvBuffers, terrain_origins) / sizeof(void*));
  const int idx_hidx = (int)(offsetof(B200EnvBuffers, height_index) / sizeof(void*));
  for (int i = 0; i < n; ++i) {
    if (ptrs[i]) continue;
    if (i == idx_hidx) continue;
    if (i == idx_height_samples && !env->p.has_height_samples) continue;
    if (i == idx_origins && !env->p.curriculum) continue;
    b200_set_error("%s: buffer #%d of B200EnvBuffers is NULL", who, i);
    return -1;
  }
  return 0;
}

int b200_pd_torques(B200Env* env, const B200EnvBuffers* bufs, const float* actions_in, int clip_and_store, void* stream) {
  if (int rc = check_bufs(env, bufs, "b200_pd_torques")) return rc;
  B200_CHECK_ARG(!clip_and_store || actions_in, "b200_pd_torques: actions_in is NULL");
  const int64_t n = (int64_t)env->p.num_envs * B200_NUM_DOF;
  pd_torques_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(env->p, *bufs, actions_in, clip_and_store);
  B200_CHECK_LAUNCH("pd_torques_kernel");
  return 0;
}

int b200_post_physics_step_parts(B200Env* env, const B200EnvBuffers* bufs, int64_t common_step_counter, int parts, void* stream) {
  if (int rc = check_bufs(env, bufs, "b200_post_physics_step")) return rc;
  if (int rc = post_physics_attr()) return rc;
  if (parts & 1) {
    if (env->p.command_curriculum)
      if (int rc = launch_command_curriculum(env, bufs, common_step_counter, nullptr, (cudaStream_t)stream)) return rc;
    if (int rc = launch_post_physics(env, bufs, common_step_counter, nullptr, (cudaStream_t)stream)) return rc;
    B200_CHECK_LAUNCH("post_physics_kernel");
  }
  if (parts & 2) {
    extras_kernel<<<B200_NUM_REWARD_TERMS + 1, 256, 0, (cudaStream_t)stream>>>(env->p, *bufs, nullptr);
    B200_CHECK_LAUNCH("extras_kernel");
  }
  return 0;
}

int b200_post_physics_step(B200Env* env, const B200EnvBuffers* bufs, int64_t common_step_counter, void* stream) {
  return b200_post_physics_step_parts(env, bufs, common_step_counter, 3, stream);
}

__global__ void counter_add_kernel(int64_t* c, int64_t delta) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *c += delta;
}

int b200_counter_add(int64_t* counter, int64_t delta, void* stream) {
  B200_CHECK_ARG(counter, "b200_counter_add: null counter");
  counter_add_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counter, delta);
  B200_CHECK_LAUNCH("counter_add_kernel");
  return 0;
}

int b200_post_physics_step_dev_parts(B200Env* env, const B200EnvBuffers* bufs, int64_t* step_counter_dev, int parts, void* stream) {
  if (int rc = check_bufs(env, bufs, "b200_post_physics_step_dev")) return rc;
  B200_CHECK_ARG(step_counter_dev, "b200_post_physics_step_dev: null counter");
  // go2.py:355: every kernel of the step reads counter + 1; extras_kernel, the last launch, stores the increment (one launch
  // fewer on the rollout's critical path than a counter kernel in front)
  if (parts & 1) {
    if (int rc = post_physics_attr()) return rc;
    if (env->p.command_curriculum)
      if (int rc = launch_command_curriculum(env, bufs, 1, step_counter_dev, (cudaStream_t)stream)) return rc;
    if (int rc = launch_post_physics(env, bufs, 1, step_counter_dev, (cudaStream_t)stream)) return rc;
    B200_CHECK_LAUNCH("post_physics_kernel");
  }
  if (parts & 2) {
    extras_kernel<<<B200_NUM_REWARD_TERMS + 1, 256, 0, (cudaStream_t)stream>>>(env->p, *bufs, step_counter_dev);
    B200_CHECK_LAUNCH("extras_kernel");
  }
  return 0;
}

int b200_post_physics_step_dev(B200Env* env, const B200EnvBuffers* bufs, int64_t* step_counter_dev, void* stream) {
  return b200_post_physics_step_dev_parts(env, bufs, step_counter_dev, 3, stream);
}

int b200_reset_all(B200Env* env, const B200EnvBuffers* bufs, int64_t common_step_counter, int init_done, void* stream) {
  if (int rc = check_bufs(env, bufs, "b200_reset_all")) return rc;
  reset_all_kernel<<<(env->p.num_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(env->p, *bufs, common_step_counter, init_done);
  B200_CHECK_LAUNCH("reset_all_kernel");
  extras_kernel<<<B200_NUM_REWARD_TERMS + 1, 256, 0, (cudaStream_t)stream>>>(env->p, *bufs, nullptr);
  B200_CHECK_LAUNCH("extras_kernel");
  return 0;
}

int b200_env_init_randomisation(B200Env* env, const B200EnvBuffers* bufs, const B200InitParams* init, void* stream) {
  if (int rc = check_bufs(env, bufs, "b200_env_init_randomisation")) return rc;
  B200_CHECK_ARG(init, "b200_env_init_randomisation: null init params");
  B200_CHECK_ARG(init->num_init_levels <= 0 || (env->p.has_height_samples && init->terrain_cols > 0 && bufs->terrain_origins),
                 "b200_env_init_randomisation: terrain levels need terrain_origins / terrain_cols");
  B200_CHECK_ARG(init->num_init_levels > 0 || init->grid_cols > 0, "b200_env_init_randomisation: grid_cols must be > 0 on a plane");
  env_init_kernel<<<(env->p.num_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(env->p, *bufs, *init);
  B200_CHECK_LAUNCH("env_init_kernel");
  return 0;
}

int b200_get_heights(B200Env* env, const B200EnvBuffers* bufs, void* stream) {
  if (int rc = check_bufs(env, bufs, "b200_get_heights")) return rc;
  const int64_t n = (int64_t)env->p.num_envs * env->p.num_scan;
  heights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(env->p, *bufs);
  B200_CHECK_LAUNCH("heights_kernel");
  return 0;
}

}  // extern "C"
