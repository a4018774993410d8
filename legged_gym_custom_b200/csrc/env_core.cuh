// One env step, as pieces that one warp / one thread executes -- shared CUDA / host source.
//
// Replaces Go2Robot.post_physics_step and its callees (go2.py:345-387; full list in
// include/b200gym.h at b200_post_physics_step).  The step is cut into PIECES of different
// natural width (row stages over the 32 lanes of a warp, per-body / per-dof / per-leg items,
// reward-term parts, one reduce / reset thread per env; see "the whole env step" below).
// Pieces communicate only through the per-env scratch (shared memory on the GPU); the CUDA
// kernel (env_kernels.cu) packs them onto a CTA of 8 envs with barriers in between, and
// tests/host_emul compiles the same source with g++ and runs the pieces one after the other,
// which is how the kernel source itself is checked against the golden vectors without a GPU.
//
// Arithmetic contract (parity with the reference's torch path, SURVEY.md §7 hard part 1):
//  * this TU is compiled with -fmad=false (g++: -ffp-contract=off): every * and + rounds
//    separately, as torch's op-by-op evaluation does;
//  * torch-CPU's small-vector norm accumulates with FMA (x*x, then fma(y,y,acc), ...) for
//    2- and 3-element reductions and with separately rounded adds for the 4-element
//    quaternion norm (measured) -- norm2/norm3/quat-normalise below
//    reproduce exactly that, because termination / curriculum / zero-command masks and the
//    height-sample indices must be bit-exact;
//  * division and sqrt are IEEE (nvcc defaults -prec-div/-prec-sqrt = true).
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/b200gym.h"
#include "philox.cuh"

#if defined(__CUDA_ARCH__)
#define B200_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define B200_LDG(p) __ldg(p)
#define B200_WARP_SYNC() __syncwarp()
#else
#define B200_FMA(a, b, c) fmaf((a), (b), (c))
#define B200_LDG(p) (*(p))
#define B200_WARP_SYNC() ((void)0)
#endif
#define B200_FOR_LANES(lane) for (int lane = lane_lo; lane < lane_hi; ++lane)

// the go2 layout (every registered go2 task): kernels instantiated with FIXED = true bake it in
#define B200_GO2_HISTORY 10
#define B200_GO2_SCAN_NX 12
#define B200_GO2_SCAN_NY 11

#define B200_TWO_PI_F 6.283185307179586f   /* fp32(2*np.pi) */
#define B200_PI_F 3.141592653589793f

struct float3_ {
  float x, y, z;
};

// ---- per-warp scratch -------------------------------------------------------------------
// The three bulk rows are sized by the instantiation: EnvScratch (below) holds any supported layout; the kernel that
// assembles its output rows in a shared-memory tile (env_kernels.cu, post_physics_tile_kernel) keeps `cur` / `tail` there
// and instantiates the struct with token-sized arrays.
template <int CUR_CAP, int TAIL_CAP, int SCAN_CAP>
struct alignas(16) EnvScratchT {
  // the new proprioceptive row (unclipped); the history rows never pass through shared memory (env_hist_*)
  float cur[CUR_CAP];
  float tail[TAIL_CAP];                 // priv | est | scan  = critic tail (16 B aligned pieces for 29+3+132)
  float heights[SCAN_CAP];
  // staged inputs
  float root[16];
  float dof[24];                 // interleaved pos, vel
  float contact[B200_NUM_BODIES * 3 + 3];
  float feet_z[4];
  float act[12], tq[12], last_act[12], last_dv[12], last_tq[12];
  float cmd[4], lch[4], fat[4];
  float sums[B200_NUM_REWARD_TERMS + 2];
  float origin[4];
  float jump_flag;
  int32_t last_contacts[4];
  int32_t outliers[32];
  int64_t ep_len, level, type;
  // item stage (one thread per item): per-body contact tests, per-dof products, per-leg gait terms, the four angles
  int32_t body_hit[32];          // bit 0: |F| > 1 (termination test), bit 1: |F| > 0.1 (collision test)
  float dofv[10][12];            // per-dof contributions of the dof-summed reward terms (rows: DV_*)
  float leg_sin[4], leg_cos[4];  // sin / cos of 2 pi phase, contact order fl, fr, bl, br
  int32_t leg_stance[4];
  float ang[4];                  // roll, pitch, yaw, heading
  // termination flags (env_item_flags): known before any reward is, which lets the history rows move meanwhile
  int32_t early_reset, early_time_out, early_refill;
  uint32_t reset_draws[4 * 7];   // env_reset_draw (only when early_reset)
  // lin_vel_x command range as fp32 {lo, span}: [0] in force (periodic resampling), [1] for the resets of this step --
  // they differ only on the step a command curriculum moves the range (go2.py:222-223 runs before _resample_commands)
  float cc_lo[2], cc_span[2];    // (16 bytes: also keeps sizeof / 16 odd, see the static_assert below)
  // results of the item stage, the reward terms and env_finalize
  float blv[4], bav[4], pg[4], rpy[4], phases[8];
  float cmd_out[4], lch_out[4], fat_out[4];
  float root_out[16], dof_out[24];
  float term[B200_NUM_REWARD_TERMS + 2];
  float origin_out[4];
  float rew, jump_flag_out;
  int32_t contact_filt[4], contact_cur[4];
  int32_t reset, time_out, root_dirty, dof_dirty;
  int64_t ep_len_out, level_out;
};
typedef EnvScratchT<B200_MAX_PROPRIO, 32 + 4 + B200_MAX_SCAN, B200_MAX_SCAN> EnvScratch;

// lanes of a warp address consecutive EnvScratch objects (lane = env slot): an odd number of 16-byte units per object
// puts 8 consecutive slots on 8 different bank groups
static_assert((sizeof(EnvScratch) / 16) % 2 == 1, "EnvScratch stride would bank-conflict the per-env lanes");

// Per-dof / per-observation constants that are indexed by LANE: out of the constant bank (a lane-varying index
// serialises there) into a small table -- shared memory on the GPU, filled once per CTA by env_tables_fill.
struct EnvTables {
  float default_dof_pos[B200_NUM_DOF], dof_pos_lo[B200_NUM_DOF], dof_pos_hi[B200_NUM_DOF];
  float dof_vel_soft[B200_NUM_DOF];      // dof_vel_limits * soft_dof_vel_limit
  float torque_soft[B200_NUM_DOF];       // torque_limits * soft_torque_limit
  // cur_obs[i] = (scratch_as_floats[cur_idx[i]] - cur_off[i]) * cur_scl[i] + (2u - 1) * noise[i]   (go2.py:506-519)
  int32_t cur_idx[B200_MAX_PROPRIO];
  float cur_off[B200_MAX_PROPRIO], cur_scl[B200_MAX_PROPRIO], noise[B200_MAX_PROPRIO];
};

struct alignas(16) f4_ {
  float x, y, z, w;
};

// ---- small math, in the reference's op order ----------------------------------------------
B200_HD float norm2_fma(float x, float y) { return sqrtf(B200_FMA(y, y, x * x)); }
B200_HD float norm3_fma(float x, float y, float z) { return sqrtf(B200_FMA(z, z, B200_FMA(y, y, x * x))); }

// isaacgym.torch_utils.quat_rotate_inverse: a - b + c
B200_HD float3_ quat_rotate_inverse(const float* q, float vx, float vy, float vz) {
  const float x = q[0], y = q[1], z = q[2], w = q[3];
  const float s = 2.0f * (w * w) - 1.0f;
  const float cx = y * vz - z * vy, cy = z * vx - x * vz, cz = x * vy - y * vx;
  const float d = (x * vx + y * vy) + z * vz;
  float3_ r;
  r.x = (vx * s - cx * w * 2.0f) + x * d * 2.0f;
  r.y = (vy * s - cy * w * 2.0f) + y * d * 2.0f;
  r.z = (vz * s - cz * w * 2.0f) + z * d * 2.0f;
  return r;
}

// heading = atan2(fwd.y, fwd.x), fwd = quat_apply(q, (1,0,0))   (go2.py:400-401)
B200_HD float heading_of(const float* q) {
  const float x = q[0], y = q[1], z = q[2], w = q[3];
  const float t1 = z * 2.0f, t2 = -y * 2.0f;          // t = cross(xyz, (1,0,0)) * 2 = (0, 2z, -2y)
  const float fx = 1.0f + (y * t2 - z * t1);
  const float fy = w * t1 + (-(x * t2));
  return atan2f(fy, fx);
}

// legged_gym/utils/math.py:45-48: python-style remainder, then fold (pi, 2pi) down
B200_HD float wrap_to_pi(float a) {
  float r = fmodf(a, B200_TWO_PI_F);
  if (r != 0.0f && r < 0.0f) r += B200_TWO_PI_F;
  return r - (r > B200_PI_F ? B200_TWO_PI_F : 0.0f);
}

// torch.clip semantics (NaN passes through).  Device: max.NaN / min.NaN, two instructions.
B200_HD float clampf(float v, float lo, float hi) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(v), "f"(lo));
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(r), "f"(hi));
  return r;
#else
  return v < lo ? lo : (v > hi ? hi : v);
#endif
}
B200_HD float sq(float v) { return v * v; }

// ---- height scan: one point (legged_robot.py:1018-1025, math.py:38-42) ------------------
struct YawQuat {
  float z, w;
};
B200_HD YawQuat yaw_quat(const float* q) {
  const float n2 = q[2] * q[2] + q[3] * q[3];          // ((0+0)+z^2)+w^2, separately rounded
  float n = sqrtf(n2);
  n = n < 1e-9f ? 1e-9f : n;
  YawQuat r;
  r.z = q[2] / n;
  r.w = q[3] / n;
  return r;
}
// x / h for the height-field index, correctly rounded without the division: with y = RN(1/h), two Newton steps on the
// quotient (q += RN(x - q h) y, FMA residuals) give RN(x / h) -- the sequence div.rn itself runs, minus its reciprocal
// refinement and special-case branch, which are hoisted because h is a constant of the terrain.  Verified bit for bit
// against x / h for EVERY fp32 x with 2^-31 <= |x| < 2^24 and h in {0.05, 0.07, 0.1, 0.125, 0.2, 0.25, 0.3, 1/3}
// (tests/test_div_const.py).  The caller clamps x into that range first.
B200_HD float div_by_const(float x, float h, float y) {
  float q = x * y;
  float r = B200_FMA(-q, h, x);
  q = B200_FMA(r, y, q);
  r = B200_FMA(-q, h, x);
  return B200_FMA(r, y, q);
}
// torch: idx = clip((x / h).long(), 0, n - 2).  The numerator is clamped to [-2h, (n + 1) h] first: inside that window
// nothing changes, outside it the quotient stays on the same side of the final clip (division is monotonic), a NaN lands
// on cell 0 like (int)NaN does, and the float -> int conversion can no longer overflow.  n * h < 2^24 is checked at
// env creation, so div_by_const is exact on the clamped value.
B200_HD int height_index(const B200EnvParams& P, float x, int n, float inv_h) {
  const float h = P.horizontal_scale;
  const float lo = -2.0f * h, hi = (float)(n + 1) * h;
  x = fminf(fmaxf(x, lo), hi);
  const float q = P.index_div_mode == 0 ? div_by_const(x, h, inv_h) : x * inv_h;
  const int i = (int)q;                                  // .long() truncates toward zero
  return i < 0 ? 0 : (i > n - 2 ? n - 2 : i);
}
// scan point (vx, vy) of the x-major grid -> height-field cell
B200_HD void height_cell_pt(const B200EnvParams& P, float vx, float vy, YawQuat yq, const float* root_pos, float inv_h, int* px, int* py) {
  // quat_apply((0,0,z,w), (vx,vy,0)): t = cross * 2; b + w*t + cross(xyz, t)
  const float t0 = -(yq.z * vy) * 2.0f, t1 = (yq.z * vx) * 2.0f;
  float rx = (vx + yq.w * t0) + (-(yq.z * t1));
  float ry = (vy + yq.w * t1) + (yq.z * t0);
  rx = (rx + root_pos[0]) + P.border_size;
  ry = (ry + root_pos[1]) + P.border_size;
  *px = height_index(P, rx, P.hs_rows, inv_h);
  *py = height_index(P, ry, P.hs_cols, inv_h);
}
// point j of the x-major scan grid: ix = j / ny without an integer division ((2j+1)/(2ny) is never within 1/(2ny) of an
// integer, so the fp32 product truncates to the exact quotient for every j < 192, ny <= 24)
B200_HD void scan_point(const B200EnvParams& P, int j, float* vx, float* vy) {
  const int gx = (int)((float)(2 * j + 1) * (0.5f / (float)P.scan_ny));
  *vx = P.scan_x[gx];
  *vy = P.scan_y[j - gx * P.scan_ny];
}
B200_HD int hs_pitch_of(const B200EnvParams& P) { return P.hs_pitch > 0 ? P.hs_pitch : P.hs_cols; }
B200_HD float height_at(const B200EnvParams& P, const int16_t* hs, int px, int py) {
  const int pitch = hs_pitch_of(P);
  const int16_t* p = hs + (px * pitch + py);            // rows * pitch < 2^31 (checked at env creation)
  const int16_t a = B200_LDG(p);
  const int16_t b = B200_LDG(p + pitch);
  const int16_t c = B200_LDG(p + 1);
  int16_t m = a < b ? a : b;
  m = m < c ? m : c;
  return (float)m * P.vertical_scale;
}

// ---- command resampling for one env (go2.py:413-464) -------------------------------------
// (`lo_x`, `span_x`: the lin_vel_x range -- P.cmd_lo[0] / P.cmd_span[0] unless a command curriculum moves it)
B200_HD void resample_commands_from(const B200EnvParams& P, const uint32_t* r_v, const float* quat, float* cmd, float lo_x, float span_x);
B200_HD void resample_commands(const B200EnvParams& P, uint32_t site, uint32_t step, uint32_t e, const float* quat, float* cmd,
                               float lo_x, float span_x) {
  const Philox4 r = keyed_block(P.seed, site, step, e, 0);
  resample_commands_from(P, r.v, quat, cmd, lo_x, span_x);
}
B200_HD void resample_commands_from(const B200EnvParams& P, const uint32_t* r_v, const float* quat, float* cmd, float lo_x, float span_x) {
  struct { uint32_t v[4]; } r = {{r_v[0], r_v[1], r_v[2], r_v[3]}};
  cmd[0] = span_x * u32_to_uniform(r.v[0]) + lo_x;
  cmd[1] = P.cmd_span[1] * u32_to_uniform(r.v[1]) + P.cmd_lo[1];
  if (P.heading_command)
    cmd[3] = P.cmd_span[3] * u32_to_uniform(r.v[2]) + P.cmd_lo[3];
  else
    cmd[2] = P.cmd_span[2] * u32_to_uniform(r.v[2]) + P.cmd_lo[2];
  const float keep = norm2_fma(cmd[0], cmd[1]) > 0.2f ? 1.0f : 0.0f;
  cmd[0] *= keep;
  cmd[1] *= keep;
  if (P.zero_command && u32_to_uniform(r.v[3]) < P.zero_command_prob) {
    cmd[0] *= 0.0f;
    cmd[1] *= 0.0f;
    cmd[2] *= 0.0f;
    if (P.heading_command) cmd[3] = heading_of(quat);
  }
}

// ---- command curriculum (go2.py:80-107) ------------------------------------------------------
// fp32 {lo, span} of a double {lo, hi} pair: torch_rand_float forms (upper - lower) in double, then multiplies in fp32
B200_HD void command_range_f32(const double* lo_hi, float* lo, float* span) {
  *lo = (float)lo_hi[0];
  *span = (float)(lo_hi[1] - lo_hi[0]);
}
// `count` envs reset on a step with common_step_counter % max_episode_length == 0 and `sum` is the sum of their
// episode_sums[tracking_lin_vel] (this step's term included): new {lo, hi} from the one in force, in the reference's
// arithmetic -- the mean and its comparison in fp32 (0-dim tensor vs Python float), np.clip on Python floats in double.
B200_HD void command_curriculum_rule(const B200EnvParams& P, int count, double sum, const double* in_force, double* next) {
  double lo = in_force[0], hi = in_force[1];
  if (count > 0) {
    const float mean = (float)(sum / (double)count) / (float)P.max_episode_length;
    if (mean > P.cc_threshold) {
      const double d = P.cc_vel_increment, a = lo - d;
      // np.clip(a, a_min, a_max) = minimum(a_max, maximum(a, a_min)); go2.py:100-103 passes a_max = a itself
      const double a_min = P.cc_max_reverse_vel, a_max = P.cc_max_reverse_vel < 0.0 ? 0.0 : a;
      lo = fmin(a_max, fmax(a, a_min));
      hi = fmin(P.cc_max_forward_vel, fmax(hi + d, 0.0));
    }
  }
  next[0] = lo;
  next[1] = hi;
}

// ---- reset of one env (go2.py:207-263 with legged_robot.py:481-574) ----------------------
// Operates on the caller's register copies (no shared-memory read-modify-write).
struct ResetState {
  float root[13], dof[24], cmd[4], origin[3], lch[4], fat[4];
  float cc_lo, cc_span;          // lin_vel_x range the reset resamples from
  int32_t contact_cur[4];
  int64_t level, ep_len;
};
// The reset draws of one env are 7 Philox blocks (block id -> site, block): 0-2 RESET_DOFS 0-2 (lanes 0..11), 3-4 RESET_ROOT
// 0-1 (lanes 0..7), 5 CURRICULUM 0, 6 CMD_RESET 0.  They are independent of everything else, so the CUDA kernel draws them
// with 7 lanes in parallel (env_reset_draw -> EnvScratch::reset_draws) and reset_env consumes the 28 words.
#define B200_RESET_BLOCKS 7
B200_HD void env_reset_draw(const B200EnvParams& P, uint32_t* draws /* [28] */, uint32_t e, uint32_t step, int block) {
  const uint32_t site = block < 3 ? SITE_RESET_DOFS : (block < 5 ? SITE_RESET_ROOT : (block == 5 ? SITE_CURRICULUM : SITE_CMD_RESET));
  const uint32_t blk = block < 3 ? block : (block < 5 ? block - 3 : 0);
  const Philox4 r = keyed_block(P.seed, site, step, e, blk);
  for (int i = 0; i < 4; ++i) draws[block * 4 + i] = r.v[i];
}
B200_HD void reset_env(const B200EnvParams& P, const B200EnvBuffers& B, ResetState& R, int64_t type, const uint32_t* draws,
                       int do_curriculum) {
  if (P.curriculum && do_curriculum) {                  // legged_robot.py:543-574
    const float dist = norm2_fma(R.root[0] - R.origin[0], R.root[1] - R.origin[1]);
    const int up = dist > P.promote_dist;
    const float expected = norm2_fma(R.cmd[0], R.cmd[1]) * P.max_episode_length_s;
    const int down = dist < expected * P.demote_threshold;
    int64_t lv = R.level + up - down;
    if (lv >= P.max_terrain_level)
      lv = (int64_t)(draws[20] % (uint32_t)P.max_terrain_level);
    else
      lv = lv < 0 ? 0 : lv;
    R.level = lv;
    const float* o = B.terrain_origins + (lv * P.terrain_cols + type) * 3;
    R.origin[0] = B200_LDG(o);
    R.origin[1] = B200_LDG(o + 1);
    R.origin[2] = B200_LDG(o + 2);
  }
  for (int d = 0; d < B200_NUM_DOF; ++d) {              // _reset_dofs
    const float u = u32_to_uniform(draws[d]);
    R.dof[2 * d] = P.default_dof_pos[d] + (P.dof_reset_span * u + P.dof_reset_lo);
    R.dof[2 * d + 1] = 0.0f;
  }
  for (int i = 0; i < 13; ++i) R.root[i] = P.base_init_state[i];   // _reset_root_states
  for (int i = 0; i < 3; ++i) R.root[i] += R.origin[i];
  int lane0 = 0;
  if (P.custom_origins) {
    for (int i = 0; i < 2; ++i) R.root[i] += 2.0f * u32_to_uniform(draws[12 + i]) + -1.0f;
    lane0 = 2;
  }
  for (int i = 0; i < 6; ++i) R.root[7 + i] = 1.0f * u32_to_uniform(draws[12 + lane0 + i]) + -0.5f;
  resample_commands_from(P, draws + 24, R.root + 3, R.cmd, R.cc_lo, R.cc_span);
  for (int f = 0; f < 4; ++f) {
    R.lch[f] = 0.0f;
    R.fat[f] = 0.0f;
    R.contact_cur[f] = 0;                                // last_contacts[env_ids] = 0 (go2.py:242)
  }
  R.ep_len = 0;
}

// ---- the item stage: everything between "state loaded" and "observations assembled" that is the same formula on
// 19 bodies / 12 dofs / 4 legs / 4 angles, plus the per-env pieces the reward terms share (base-frame velocities,
// foot contacts, command update, push, termination flags).  One ITEM = one thread; the CUDA kernel packs the items of
// the CTA's 8 envs type by type into full warps (env_kernels.cu), the host emulation runs them one after the other.
// Items only read staged inputs (stage 0 / 1) and write disjoint scratch fields, so their order does not matter.
// Contact tests compare the (FMA-accumulated, as torch does) squared norm with the pre-rounded squared threshold:
// sqrt_rn(s) > t  <=>  s > contact_thr2 exactly.
enum { DV_ACTION_RATE = 0, DV_DELTA_TORQUES, DV_DOF_ACC, DV_DQ2, DV_POS_LIMITS, DV_VEL2, DV_VEL_LIMITS, DV_ABS_DQ, DV_TORQUE_LIMITS, DV_TORQUES2 };

B200_HD float gait_phase(const B200EnvParams& P, int64_t ep) { return fmodf((float)ep * P.dt, P.period) / P.period; }

template <class SC>
B200_HD void env_item_body(const B200EnvParams& P, SC& S, int b) {
  const float* c = S.contact + b * 3;
  const float n2 = B200_FMA(c[2], c[2], B200_FMA(c[1], c[1], c[0] * c[0]));
  S.body_hit[b] = (n2 > P.contact_thr2_term ? 1 : 0) | (n2 > P.contact_thr2_collision ? 2 : 0);
}

template <class SC>
B200_HD void env_item_dof(const B200EnvParams& P, const EnvTables& T, SC& S, int d) {
  const float pos = S.dof[2 * d], vel = S.dof[2 * d + 1];
  const float dq = pos - T.default_dof_pos[d];
  S.dofv[DV_ACTION_RATE][d] = sq(S.last_act[d] - S.act[d]);
  S.dofv[DV_DELTA_TORQUES][d] = sq(S.tq[d] - S.last_tq[d]);
  S.dofv[DV_DOF_ACC][d] = sq((S.last_dv[d] - vel) / P.dt);
  S.dofv[DV_DQ2][d] = sq(dq);
  const float lo = pos - T.dof_pos_lo[d], hi = pos - T.dof_pos_hi[d];
  S.dofv[DV_POS_LIMITS][d] = -(lo > 0.0f ? 0.0f : lo) + (hi < 0.0f ? 0.0f : hi);
  S.dofv[DV_VEL2][d] = sq(vel);
  S.dofv[DV_VEL_LIMITS][d] = clampf(fabsf(vel) - T.dof_vel_soft[d], 0.0f, 1.0f);
  S.dofv[DV_ABS_DQ][d] = fabsf(dq);
  const float over = fabsf(S.tq[d]) - T.torque_soft[d];
  S.dofv[DV_TORQUE_LIMITS][d] = over < 0.0f ? 0.0f : over;
  S.dofv[DV_TORQUES2][d] = sq(S.tq[d]);
}

// gait phase of one leg (go2.py:279-290), order fl, fr, bl, br
template <class SC>
B200_HD void env_item_leg(const B200EnvParams& P, SC& S, int f) {
  const float ph = gait_phase(P, S.ep_len + 1);
  const float off = f == 0 ? P.fl_offset : (f == 1 ? P.fr_offset : (f == 2 ? P.bl_offset : P.br_offset));
  const float keep = norm3_fma(S.cmd[0], S.cmd[1], S.cmd[2]) < 0.2f ? 0.0f : 1.0f;
  const float phf = fmodf(ph + off, 1.0f) * keep;
  S.phases[f == 0 ? 2 : (f == 1 ? 1 : (f + 1))] = phf;      // API order: phase, fr, fl, bl, br
  if (f == 0) S.phases[0] = ph;
  const float a = B200_TWO_PI_F * phf;
  const float sn = sinf(a);
  S.leg_sin[f] = sn;
  S.leg_cos[f] = cosf(a);
  S.leg_stance[f] = sn <= P.stance_threshold;
}

// a = 0 roll, 1 pitch, 2 yaw (quaternion_to_euler, go2.py:11-31), 3 heading.  The heading item goes on with the command
// update of _post_physics_step_callback (go2.py:390-410), which is the only consumer of the heading.
template <class SC>
B200_HD void env_item_angle(const B200EnvParams& P, SC& S, int a, uint32_t e, uint32_t step) {
  const float x = S.root[3], y = S.root[4], z = S.root[5], w = S.root[6];
  if (a == 1) {                                           // pitch (go2.py:23-25)
    const float pitch = asinf(clampf(2.0f * (w * y - z * x), -1.0f, 1.0f));
    S.ang[1] = pitch;
    S.rpy[1] = pitch;
    return;
  }
  float num, den;                                         // the three atan2 share one code path
  if (a == 0) {                                           // roll (go2.py:19-21)
    num = 2.0f * (w * x + y * z);
    den = 1.0f - 2.0f * (x * x + y * y);
  } else if (a == 2) {                                    // yaw (go2.py:27-29)
    num = 2.0f * (w * z + x * y);
    den = 1.0f - 2.0f * (y * y + z * z);
  } else {                                                // heading = atan2(fwd.y, fwd.x), fwd = quat_apply(q, (1,0,0))
    const float t1 = z * 2.0f, t2 = -y * 2.0f;
    den = 1.0f + (y * t2 - z * t1);
    num = w * t1 + (-(x * t2));
  }
  const float v = atan2f(num, den);
  S.ang[a] = v;
  if (a == 0) S.rpy[0] = v;
  if (a == 2) S.rpy[2] = v;
  if (a == 3) {
    float cmd[4] = {S.cmd[0], S.cmd[1], S.cmd[2], S.cmd[3]};
    const int64_t ep = S.ep_len + 1;                      // go2.py:354
    if ((uint32_t)ep % (uint32_t)P.resample_interval == 0u) resample_commands(P, SITE_CMD_PERIODIC, step, e, S.root + 3, cmd, S.cc_lo[0], S.cc_span[0]);
    if (P.heading_command) cmd[2] = clampf(wrap_to_pi(cmd[3] - v) * P.heading_error_gain, -1.0f, 1.0f);
    for (int i = 0; i < 4; ++i) S.cmd_out[i] = cmd[i];
  }
}

// base-frame velocities and gravity (go2.py:357-360)
template <class SC>
B200_HD void env_item_velocities(SC& S) {
  const float* q = S.root + 3;
  const float3_ blv = quat_rotate_inverse(q, S.root[7], S.root[8], S.root[9]);
  const float3_ bav = quat_rotate_inverse(q, S.root[10], S.root[11], S.root[12]);
  const float3_ pg = quat_rotate_inverse(q, 0.0f, 0.0f, -1.0f);
  S.blv[0] = blv.x; S.blv[1] = blv.y; S.blv[2] = blv.z;
  S.bav[0] = bav.x; S.bav[1] = bav.y; S.bav[2] = bav.z;
  S.pg[0] = pg.x; S.pg[1] = pg.y; S.pg[2] = pg.z;
}

// update_feet_states (go2.py:266-328; leg order of contacts / feet: fl, fr, bl, br) and _push_robots
// (legged_robot.py:535-540).  Stage 0 has already copied root -> root_out, so the push lands on the copy.
template <class SC>
B200_HD void env_item_feet_push(const B200EnvParams& P, SC& S, uint32_t e, int64_t step64) {
  for (int f = 0; f < 4; ++f) {
    const int cur = S.contact[P.feet[f] * 3 + 2] > 1.0f;
    const int filt = cur | (S.last_contacts[f] != 0);
    S.contact_cur[f] = cur;
    S.contact_filt[f] = filt;
    S.lch_out[f] = filt ? S.feet_z[f] : S.lch[f];
  }
  int root_dirty = 0;
  if (P.push_robots && ((uint32_t)step64 % (uint32_t)P.push_interval) == 0u) {
    const float span = P.max_push_vel - -P.max_push_vel;
    S.root_out[7] = span * keyed_uniform(P.seed, SITE_PUSH, (uint32_t)step64, e, 0) + -P.max_push_vel;
    S.root_out[8] = span * keyed_uniform(P.seed, SITE_PUSH, (uint32_t)step64, e, 1) + -P.max_push_vel;
    root_dirty = 1;
  }
  S.root_dirty = root_dirty;
}

// check_termination (go2.py:186-204) and the jump flags of the NEXT step (go2.py:487-494, from this step's heights):
// known before any reward is, which lets the history rows move while the rewards are computed
template <class SC>
B200_HD void env_item_flags(const B200EnvParams& P, SC& S) {
  int reset = 0;
  for (int i = 0; i < P.n_termination; ++i) {
    const float* c = S.contact + P.termination[i] * 3;
    reset |= B200_FMA(c[2], c[2], B200_FMA(c[1], c[1], c[0] * c[0])) > P.contact_thr2_term;
  }
  const int64_t ep = S.ep_len + 1;                      // go2.py:354
  const int time_out = ep > P.max_episode_length;
  reset |= time_out;
  reset |= quat_rotate_inverse(S.root + 3, 0.0f, 0.0f, -1.0f).z > 0.0f;
  if (P.parkour) reset |= S.root[2] < -1.0f;            // a push only touches root[7:9]
  S.early_reset = reset;
  S.early_time_out = time_out;
  S.early_refill = reset || ep <= 1;                    // go2.py:570-574: episode_length_buf <= 1 after the reset
  float jf = S.jump_flag;
  if (P.parkour) {
    int n = 0;
    for (int l = 0; l < 32; ++l) n += S.outliers[l];
    jf = n >= 8 ? 1.0f : 0.0f;
  }
  S.jump_flag_out = jf;
}

// item ids of one env: bodies | dofs | legs | angles | velocities | feet + push | flags
enum { ITEM_BODY0 = 0, ITEM_DOF0 = ITEM_BODY0 + B200_NUM_BODIES, ITEM_LEG0 = ITEM_DOF0 + B200_NUM_DOF, ITEM_ANGLE0 = ITEM_LEG0 + 4,
       ITEM_VEL = ITEM_ANGLE0 + 4, ITEM_FEET, ITEM_FLAGS, ITEM_COUNT };
template <class SC>
B200_HD void env_item(const B200EnvParams& P, const EnvTables& T, SC& S, int item, uint32_t e, int64_t step64) {
  if (item < ITEM_DOF0) env_item_body(P, S, item - ITEM_BODY0);
  else if (item < ITEM_LEG0) env_item_dof(P, T, S, item - ITEM_DOF0);
  else if (item < ITEM_ANGLE0) env_item_leg(P, S, item - ITEM_LEG0);
  else if (item < ITEM_VEL) env_item_angle(P, S, item - ITEM_ANGLE0, e, (uint32_t)step64);
  else if (item == ITEM_VEL) env_item_velocities(S);
  else if (item == ITEM_FEET) env_item_feet_push(P, S, e, step64);
  else env_item_flags(P, S);
}

// ---- the reward terms (legged_robot.py:216-237 and every _reward_*): term k is evaluated by part k % B200_TERM_PARTS --
// the CUDA kernel runs the parts on different warps (lane = env slot), the emulation one after the other.  A term reads
// the staged inputs and the item-stage results and writes term[k] = value * (scale * dt), 0 for a disabled term; the two
// stateful terms also publish their state (feet_air_time -> fat_out, heading_alignment -> cmd_out[3]).
#define B200_TERM_PARTS 8
template <bool FIXED, class SC>
B200_HD void env_terms_part(const B200EnvParams& P, SC& S, int part) {
  const int num_scan = FIXED ? B200_GO2_SCAN_NX * B200_GO2_SCAN_NY : P.num_scan;
  const float* sc = P.reward_scales;
  const float* cmd = S.cmd_out;                         // after resampling / heading update, before a reset
  const float* root = S.root_out;                       // world-frame rewards see the push (Appendix C.7)
  const float* blv = S.blv;
  const float* bav = S.bav;
  const float* pg = S.pg;
#define B200_MINE(NAME) ((B200_REW_##NAME % B200_TERM_PARTS) == part)
#define B200_ON(NAME) (sc[B200_REW_##NAME] != 0.0f)
#define B200_PUT(NAME, VALUE) S.term[B200_REW_##NAME] = B200_ON(NAME) ? (VALUE) * sc[B200_REW_##NAME] : 0.0f
#define B200_DOFSUM(ROW, OUT)                                   \
  {                                                             \
    OUT = 0.0f;                                                 \
    for (int d_ = 0; d_ < 12; ++d_) OUT += S.dofv[ROW][d_];      \
  }
#define B200_DOFSUM_TERM(NAME, ROW)            \
  if (B200_MINE(NAME)) {                       \
    float a = 0.0f;                            \
    if (B200_ON(NAME)) B200_DOFSUM(ROW, a)     \
    B200_PUT(NAME, a);                         \
  }
  B200_DOFSUM_TERM(action_rate, DV_ACTION_RATE)
  if (B200_MINE(ang_vel_xy)) B200_PUT(ang_vel_xy, sq(bav[0]) + sq(bav[1]));
  if (B200_MINE(base_height)) {
    float a = 0.0f;
    if (B200_ON(base_height)) {
      for (int j = 0; j < num_scan; ++j) a += root[2] - S.heights[j];
      a = sq(a / (float)num_scan - P.base_height_target);
    }
    B200_PUT(base_height, a);
  }
  if (B200_MINE(calf_collision)) {
    float a = 0.0f;
    if (B200_ON(calf_collision))
      for (int f = 0; f < 4; ++f) a += (S.body_hit[P.calves[f]] & 2) ? 1.0f : 0.0f;
    B200_PUT(calf_collision, a);
  }
  if (B200_MINE(calf_pos)) {
    float a = 0.0f;
    if (B200_ON(calf_pos))
      for (int f = 0; f < 4; ++f) a += S.dofv[DV_DQ2][P.calf_joints[f]];
    B200_PUT(calf_pos, a);
  }
  if (B200_MINE(calf_symmetry))
    B200_PUT(calf_symmetry, fabsf(S.dof[2 * P.calf_joints[0]] - S.dof[2 * P.calf_joints[1]]) +
                                fabsf(S.dof[2 * P.calf_joints[2]] - S.dof[2 * P.calf_joints[3]]));
  if (B200_MINE(collision)) {
    float a = 0.0f;
    if (B200_ON(collision))
      for (int i = 0; i < P.n_penalised; ++i) a += (S.body_hit[P.penalised[i]] & 2) ? 1.0f : 0.0f;
    B200_PUT(collision, a);
  }
  B200_DOFSUM_TERM(delta_torques, DV_DELTA_TORQUES)
  B200_DOFSUM_TERM(dof_acc, DV_DOF_ACC)
  B200_DOFSUM_TERM(dof_error, DV_DQ2)
  B200_DOFSUM_TERM(dof_pos_limits, DV_POS_LIMITS)
  B200_DOFSUM_TERM(dof_vel, DV_VEL2)
  B200_DOFSUM_TERM(dof_vel_limits, DV_VEL_LIMITS)
  if (B200_MINE(feet_air_time)) {                       // go2.py:819-831 (stateful)
    float a = 0.0f;
    if (B200_ON(feet_air_time)) {
      // update_feet_states has already overwritten last_contacts with the CURRENT contacts (go2.py:307-310), so the
      // "filtered" contact of this reward (go2.py:824-825) is just the current one
      for (int f = 0; f < 4; ++f) {
        const int now = S.contact_cur[f];
        const float first = (S.fat[f] > 0.0f && now) ? 1.0f : 0.0f;
        const float t = S.fat[f] + P.dt;
        a += (t - 0.5f) * first;
        S.fat_out[f] = t * (now ? 0.0f : 1.0f);
      }
      a *= norm2_fma(cmd[0], cmd[1]) > 0.1f ? 1.0f : 0.0f;
    }
    B200_PUT(feet_air_time, a);
  }
  if (B200_MINE(feet_contact_forces)) {
    float a = 0.0f;
    if (B200_ON(feet_contact_forces))
      for (int f = 0; f < 4; ++f) {
        const float* c = S.contact + P.feet[f] * 3;
        const float over = norm3_fma(c[0], c[1], c[2]) - P.max_contact_force;
        a += over < 0.0f ? 0.0f : over;
      }
    B200_PUT(feet_contact_forces, a);
  }
  if (B200_MINE(heading_alignment)) {                   // go2.py:734-756; wrap_to_pi mutates commands[:,3]
    float a = 0.0f;
    if (B200_ON(heading_alignment)) {
      float desired = 0.0f;
      if (P.heading_command) {
        desired = wrap_to_pi(cmd[3]);
        S.cmd_out[3] = desired;
      }
      const float moving = norm3_fma(cmd[0], cmd[1], cmd[2]) >= 0.2f ? 1.0f : 0.0f;
      a = sq(wrap_to_pi(desired - S.ang[3])) * moving;
    }
    B200_PUT(heading_alignment, a);
  }
  if (B200_MINE(hip_pos)) {
    float a = 0.0f;
    if (B200_ON(hip_pos))
      for (int f = 0; f < 4; ++f) a += S.dofv[DV_DQ2][P.hip_joints[f]];
    B200_PUT(hip_pos, a);
  }
  if (B200_MINE(jump_zone_forward_vel) || B200_MINE(jump_zone_upward_vel) || B200_MINE(min_height)) {
    const float moving = norm3_fma(cmd[0], cmd[1], cmd[2]) >= 0.2f ? 1.0f : 0.0f;
    const float jumping = S.jump_flag > 0.0f ? 1.0f : 0.0f;
    if (B200_MINE(jump_zone_forward_vel)) B200_PUT(jump_zone_forward_vel, (root[7] < 0.0f ? 0.0f : root[7]) * jumping * moving);
    if (B200_MINE(jump_zone_upward_vel)) B200_PUT(jump_zone_upward_vel, (root[9] < 0.0f ? 0.0f : root[9]) * jumping * moving);
    if (B200_MINE(min_height)) B200_PUT(min_height, clampf(P.base_height_target - root[2], 0.0f, P.base_height_target) * jumping);
  }
  if (B200_MINE(lin_vel_z)) B200_PUT(lin_vel_z, sq(blv[2]));
  if (B200_MINE(orientation)) B200_PUT(orientation, sq(pg[0]) + sq(pg[1]));
  if (B200_MINE(phase_contact_match)) {                 // go2.py:621-644
    float a = 0.0f;
    for (int f = 0; f < 4; ++f) a += (S.contact_filt[f] == S.leg_stance[f]) ? 0.25f : -0.25f;
    B200_PUT(phase_contact_match, a);
  }
  if (B200_MINE(phase_foot_lifting)) {                  // go2.py:647-678
    float a = 0.0f;
    for (int f = 0; f < 4; ++f) {
      const float h = clampf(S.feet_z[f] - S.lch_out[f], 0.0f, P.max_foot_height) / P.max_foot_height;
      a += S.leg_stance[f] ? -h : h;
    }
    B200_PUT(phase_foot_lifting, a / 2.0f);
  }
  if (B200_MINE(reverse_penalty)) B200_PUT(reverse_penalty, -(root[7] > 0.0f ? 0.0f : root[7]));
  if (B200_MINE(stand_still)) {
    float a = 0.0f;
    if (B200_ON(stand_still)) {
      B200_DOFSUM(DV_ABS_DQ, a)
      a *= norm2_fma(cmd[0], cmd[1]) < 0.1f ? 1.0f : 0.0f;
    }
    B200_PUT(stand_still, a);
  }
  if (B200_MINE(stumble_calves)) {
    int any = 0;
    if (B200_ON(stumble_calves))
      for (int f = 0; f < 4; ++f) {
        const float* c = S.contact + P.calves[f] * 3;
        any |= norm2_fma(c[0], c[1]) > 5.0f * fabsf(c[2]);
      }
    B200_PUT(stumble_calves, any ? 1.0f : 0.0f);
  }
  if (B200_MINE(stumble_feet)) {
    int any = 0;
    if (B200_ON(stumble_feet))
      for (int f = 0; f < 4; ++f) {
        const float* c = S.contact + P.feet[f] * 3;
        any |= norm2_fma(c[0], c[1]) > 5.0f * fabsf(c[2]);
      }
    B200_PUT(stumble_feet, any ? 1.0f : 0.0f);
  }
  if (B200_MINE(thigh_pos)) {
    float a = 0.0f;
    if (B200_ON(thigh_pos))
      for (int f = 0; f < 4; ++f) a += S.dofv[DV_DQ2][P.thigh_joints[f]];
    B200_PUT(thigh_pos, a);
  }
  if (B200_MINE(thigh_symmetry))
    B200_PUT(thigh_symmetry, fabsf(S.dof[2 * P.thigh_joints[0]] - S.dof[2 * P.thigh_joints[1]]) +
                                 fabsf(S.dof[2 * P.thigh_joints[2]] - S.dof[2 * P.thigh_joints[3]]));
  B200_DOFSUM_TERM(torque_limits, DV_TORQUE_LIMITS)
  B200_DOFSUM_TERM(torques, DV_TORQUES2)
  if (B200_MINE(tracking_ang_vel)) B200_PUT(tracking_ang_vel, expf(-sq(cmd[2] - bav[2]) / P.tracking_sigma));
  if (B200_MINE(tracking_lin_vel)) B200_PUT(tracking_lin_vel, expf(-(sq(cmd[0] - blv[0]) + sq(cmd[1] - blv[1])) / P.tracking_sigma));
  if (B200_MINE(tracking_pitch)) B200_PUT(tracking_pitch, expf(-sq(S.ang[1] * 57.29577951308232f - P.pitch_deg_target) / P.tracking_sigma));
  if (B200_MINE(tracking_roll)) B200_PUT(tracking_roll, expf(-sq(S.ang[0] * 57.29577951308232f - P.roll_deg_target) / P.tracking_sigma));
  if (B200_MINE(zero_cmd_dof_error)) {
    float a = 0.0f;
    if (B200_ON(zero_cmd_dof_error)) {
      B200_DOFSUM(DV_DQ2, a)
      a = a * (norm3_fma(cmd[0], cmd[1], cmd[2]) < 0.2f ? 1.0f : 0.0f);
    }
    B200_PUT(zero_cmd_dof_error, a);
  }
#undef B200_DOFSUM_TERM
#undef B200_DOFSUM
#undef B200_PUT
#undef B200_MINE
}

// ---- dry pass of a command-curriculum step (go2.py:222-223, :87): update_command_curriculum needs the mean of
// episode_sums[tracking_lin_vel] over the envs that reset on this step BEFORE any of them resamples its command.  After
// stage 0 and the item stage of env e: evaluate the part that holds the term, publish (sum so far + this step's term,
// reset flag); nothing else leaves the scratch.
template <bool FIXED, class SC>
B200_HD void env_cc_probe(const B200EnvParams& P, const B200EnvBuffers& B, SC& S, int e) {
  env_terms_part<FIXED>(P, S, B200_REW_tracking_lin_vel % B200_TERM_PARTS);
  B.cc_value[e] = S.sums[B200_REW_tracking_lin_vel] + S.term[B200_REW_tracking_lin_vel];
  B.cc_reset[e] = (uint8_t)S.early_reset;
}

// ---- compute_reward's sum (alphabetical, legged_robot.py:216-237), the termination reward, and reset_idx on this env if
// flagged (go2.py:375-376).  One thread per env, after every part of env_terms_part.
// (needs EnvScratch::reset_draws when the env resets: env_reset_draw, blocks 0..6)
template <class SC>
B200_HD void env_finalize(const B200EnvParams& P, const B200EnvBuffers& B, SC& S) {
  const float* sc = P.reward_scales;
  const int reset = S.early_reset, time_out = S.early_time_out;
  float rew = 0.0f;
  for (int k = 0; k < B200_REW_termination; ++k)
    if (sc[k] != 0.0f) rew += S.term[k];
  if (P.only_positive_rewards) rew = rew < 0.0f ? 0.0f : rew;
  {                                                     // legged_robot.py:234-237
    float r_ = 0.0f;
    if (B200_ON(termination)) {
      r_ = ((reset && !time_out) ? 1.0f : 0.0f) * sc[B200_REW_termination];
      rew += r_;
    }
    S.term[B200_REW_termination] = r_;
  }
#undef B200_ON
  S.rew = rew;
  S.reset = reset;
  S.time_out = time_out;
  S.dof_dirty = reset;
  S.level_out = S.level;
  S.ep_len_out = S.ep_len + 1;
  if (reset) {                                          // rare: everything the reset touches goes through registers
    ResetState R;
    for (int i = 0; i < 13; ++i) R.root[i] = S.root_out[i];
    for (int i = 0; i < 24; ++i) R.dof[i] = S.dof_out[i];
    for (int i = 0; i < 4; ++i) R.cmd[i] = S.cmd_out[i];
    for (int i = 0; i < 3; ++i) R.origin[i] = S.origin_out[i];
    R.level = S.level;
    R.ep_len = S.ep_len + 1;
    R.cc_lo = S.cc_lo[1];
    R.cc_span = S.cc_span[1];
    reset_env(P, B, R, S.type, S.reset_draws, 1);
    for (int i = 0; i < 13; ++i) S.root_out[i] = R.root[i];
    for (int i = 0; i < 24; ++i) S.dof_out[i] = R.dof[i];
    for (int i = 0; i < 4; ++i) {
      S.cmd_out[i] = R.cmd[i];
      S.lch_out[i] = R.lch[i];
      S.fat_out[i] = R.fat[i];
      S.contact_cur[i] = R.contact_cur[i];
    }
    for (int i = 0; i < 3; ++i) S.origin_out[i] = R.origin[i];
    S.level_out = R.level;
    S.ep_len_out = R.ep_len;
    S.root_dirty = 1;
  }
}


// ---- cur_obs (go2.py:506-519) as a table: element i = (one float of the scratch - offset) * scale ----------------
//   [0:3) base_ang_vel * 0.25 | [3:5) roll, pitch | [5:8) commands * (2, 2, 0.25) | [8:20) (dof_pos - default) * 1
//   [20:32) dof_vel * 0.05 | [32:44) actions | [44:52) sin, cos of the fr, fl, bl, br phases (go2.py:476-481)
// (x - 0) * 1 is x bit for bit, so one formula serves every segment.  Entry i of every table: env_tables_fill(P, T, i).
#define B200_SOFF(field) ((int)(offsetof(SC, field) / sizeof(float)))
template <class SC = EnvScratch>
B200_HD void env_tables_fill(const B200EnvParams& P, EnvTables& T, int i) {
  if (i < B200_NUM_DOF) {
    T.default_dof_pos[i] = P.default_dof_pos[i];
    T.dof_pos_lo[i] = P.dof_pos_lo[i];
    T.dof_pos_hi[i] = P.dof_pos_hi[i];
    T.dof_vel_soft[i] = P.dof_vel_limits[i] * P.soft_dof_vel_limit;
    T.torque_soft[i] = P.torque_limits[i] * P.soft_torque_limit;
  }
  if (i >= B200_MAX_PROPRIO) return;
  int idx = 0;
  float off = 0.0f, scl = 1.0f;
  if (i < 3) { idx = B200_SOFF(bav) + i; scl = P.obs_ang_vel; }
  else if (i < 5) { idx = B200_SOFF(rpy) + (i - 3); }
  else if (i < 8) { idx = B200_SOFF(cmd_out) + (i - 5); scl = i < 7 ? P.obs_lin_vel : P.obs_ang_vel; }
  else if (i < 20) { idx = B200_SOFF(dof_out) + 2 * (i - 8); off = P.default_dof_pos[i - 8]; scl = P.obs_dof_pos; }
  else if (i < 32) { idx = B200_SOFF(dof_out) + 2 * (i - 20) + 1; scl = P.obs_dof_vel; }
  else if (i < 44) { idx = B200_SOFF(act) + (i - 32); }
  else if (i < B200_PROPRIO) {                          // scratch order of the legs is fl, fr, bl, br
    const int leg = (i - 44) >> 1;
    const int f = leg == 0 ? 1 : (leg == 1 ? 0 : leg);
    idx = (((i - 44) & 1) ? B200_SOFF(leg_cos) : B200_SOFF(leg_sin)) + f;
  }
  T.cur_idx[i] = idx;
  T.cur_off[i] = off;
  T.cur_scl[i] = scl;
  T.noise[i] = i < B200_PROPRIO ? P.noise_vec[i] : 0.0f;
}
template <class SC>
B200_HD float cur_obs_element(const B200EnvParams& P, const EnvTables& T, const SC& S, float u, int i) {
  float v = (reinterpret_cast<const float*>(&S)[T.cur_idx[i]] - T.cur_off[i]) * T.cur_scl[i];
  if (P.add_noise) v += (2.0f * u - 1.0f) * T.noise[i];
  return v;
}

B200_HD f4_ clamp4(f4_ v, float c) {
  v.x = clampf(v.x, -c, c);
  v.y = clampf(v.y, -c, c);
  v.z = clampf(v.z, -c, c);
  v.w = clampf(v.w, -c, c);
  return v;
}

// ---- the whole env step for env `e` --------------------------------------------------------------
// All bulk rows are multiples of 4 floats (52, 520, 572, 736, 132, critic tail 164; b200_env_create
// rejects a scan size that is not) and move as 16-byte vectors.  The env step of env `e` is, in dependency order:
//   env_warp_pre      one warp, lanes [lane_lo, lane_hi): stage 0 (load the small rows), 1 (height scan)
//   env_item          ITEM_COUNT independent items (bodies, dofs, legs, angles + command update, velocities, feet + push,
//                     termination flags), one thread each
//   env_hist_load / env_hist_store   the history rows, one 16-byte vector per call: history -> clip -> obs / critic
//                     and history shifted by one slot, straight from global to global; need only the flags item
//   env_terms_part    the reward terms, B200_TERM_PARTS independent parts, one thread per (env, part)
//   env_finalize      reward sum, termination reward, reset_idx; one thread per env
//   env_warp_post     one warp: stage 3 (current observation, critic tail), 4 (write-back)
// The CUDA kernel (env_kernels.cu) maps these onto a CTA of 8 envs so that every phase fills its warps.
// FIXED = true bakes the go2 layout (history 10, scan 12 x 11) into the code: trip counts and row offsets
// become literals, loops unroll, loads batch.  FIXED = false reads them from P (any other layout).
// pt_x / pt_y: the num_scan scan points (scan_point), shared-memory tables on the GPU
#define B200_ENV_DIMS                                                                 \
  constexpr int NP = B200_PROPRIO, NP4 = NP / 4;                                      \
  const int H = FIXED ? B200_GO2_HISTORY : P.history_len;                             \
  const int NS = FIXED ? B200_GO2_SCAN_NX * B200_GO2_SCAN_NY : P.num_scan;            \
  const int NY = FIXED ? B200_GO2_SCAN_NY : P.scan_ny;                                \
  constexpr int NPRIV = 29, NEST = 3, PE = NPRIV + NEST;                              \
  const int HN = H * NP, HN4 = HN / 4, OBS = HN + NP, OBS4 = OBS / 4;                 \
  const int TAIL = PE + NS, TAIL4 = TAIL / 4, NS4 = NS / 4;                           \
  const int CRIT = OBS + TAIL;                                                        \
  const int64_t N = P.num_envs;                                                       \
  (void)H; (void)NS; (void)NY; (void)HN; (void)HN4; (void)OBS; (void)OBS4; (void)TAIL; \
  (void)TAIL4; (void)NS4; (void)CRIT; (void)N; (void)NP4; (void)PE

B200_HD bool env_layout_is_go2(const B200EnvParams& P) {
  return P.history_len == B200_GO2_HISTORY && P.scan_nx == B200_GO2_SCAN_NX && P.scan_ny == B200_GO2_SCAN_NY &&
         P.num_scan == B200_GO2_SCAN_NX * B200_GO2_SCAN_NY;
}

template <bool FIXED, class SC>
B200_HD void env_warp_scan(const B200EnvParams& P, const B200EnvBuffers& B, SC& S, const float* pt_x, const float* pt_y, int e, int lane_lo,
                           int lane_hi);
template <bool FIXED, class SC>
B200_HD void env_warp_pre(const B200EnvParams& P, const B200EnvBuffers& B, SC& S, const float* pt_x, const float* pt_y,
                          int e, int lane_lo, int lane_hi, float* tail_dst = nullptr, bool do_scan = true) {
  B200_ENV_DIMS;

  // ---- stage 0: stage the env's small rows into scratch (coalesced: consecutive lanes, consecutive floats).
  // Every load is unconditional on a clamped index and issued before the first store, so the ~25 rows cost ONE
  // memory round trip instead of one per predicated block.  The privileged statics of stage 3 are fetched here too.
  B200_FOR_LANES(lane) {
    const int64_t E = e;
    const int l4 = lane & 3, l12 = lane < 12 ? lane : 11, l13 = lane < 13 ? lane : 12, l24 = lane < 24 ? lane : 23;
    const int lc = lane + 32 < B200_NUM_BODIES * 3 ? lane + 32 : B200_NUM_BODIES * 3 - 1;
    const int ls = lane + 32 < B200_NUM_REWARD_TERMS ? lane + 32 : B200_NUM_REWARD_TERMS - 1;
    const float r_root = B.root_states[E * 13 + l13];
    const float r_dof = B.dof_state[E * 24 + l24];
    const float r_c0 = B200_LDG(B.contact_forces + E * B200_NUM_BODIES * 3 + lane);
    const float r_c1 = B200_LDG(B.contact_forces + E * B200_NUM_BODIES * 3 + lc);
    const float r_fz = B200_LDG(B.rigid_body_states + (E * B200_NUM_BODIES + P.feet[l4]) * 13 + 2);
    const float r_cmd = B.commands[E * 4 + l4];
    const float r_lch = B.last_contact_heights[E * 4 + l4];
    const float r_fat = B.feet_air_time[E * 4 + l4];
    const uint8_t r_lc = B.last_contacts[E * 4 + l4];
    const float r_act = B.actions[E * 12 + l12];
    const float r_tq = B.torques[E * 12 + l12];
    const float r_la = B.last_actions[E * 12 + l12];
    const float r_ldv = B.last_dof_vel[E * 12 + l12];
    const float r_ltq = B.last_torques[E * 12 + l12];
    const float r_s0 = B.episode_sums[E * B200_NUM_REWARD_TERMS + lane];
    const float r_s1 = B.episode_sums[E * B200_NUM_REWARD_TERMS + ls];
    const float r_org = B.env_origins[E * 3 + (lane < 3 ? lane : 2)];
    const int64_t r_ep = B.episode_length_buf[e];
    const float r_jf = B.jump_flags[e];
    const int64_t r_lv = B.terrain_levels[e];
    const int64_t r_ty = B200_LDG(B.terrain_types + e);
    // privileged statics (go2.py:528-532): mass/com 4 | friction 1 | kp - 1 (12) | kd - 1 (12); lanes 29-31 unused
    float r_priv;
    {
      const int i = lane < NPRIV ? lane : NPRIV - 1;
      const float* src = i < 4 ? B.priv_mass_params + E * 4 + i
                               : (i < 5 ? B.priv_friction + E : (i < 17 ? B.kp_kd_multipliers + E * 12 + (i - 5) : B.kp_kd_multipliers + (N + E) * 12 + (i - 17)));
      r_priv = B200_LDG(src);
      if (i >= 5) r_priv = r_priv - 1.0f;
    }
    if (lane < 13) S.root[lane] = S.root_out[lane] = r_root;      // *_out: what the step leaves behind unless a push /
    if (lane < 24) S.dof[lane] = S.dof_out[lane] = r_dof;         // reset / reward term overwrites it
    S.contact[lane] = r_c0;
    if (lane + 32 < B200_NUM_BODIES * 3) S.contact[lane + 32] = r_c1;
    if (lane < 4) {
      S.feet_z[lane] = r_fz;
      S.cmd[lane] = r_cmd;
      S.lch[lane] = r_lch;
      S.fat[lane] = S.fat_out[lane] = r_fat;
      S.last_contacts[lane] = r_lc;
    }
    if (lane < 12) {
      S.act[lane] = r_act;
      S.tq[lane] = r_tq;
      S.last_act[lane] = r_la;
      S.last_dv[lane] = r_ldv;
      S.last_tq[lane] = r_ltq;
    }
    S.sums[lane] = r_s0;
    if (lane + 32 < B200_NUM_REWARD_TERMS) S.sums[lane + 32] = r_s1;
    if (lane < 3) S.origin[lane] = S.origin_out[lane] = r_org;
    if (lane < NPRIV) (tail_dst ? tail_dst : S.tail)[lane] = clampf(r_priv, -P.clip_obs, P.clip_obs);
    if (lane == 0) {
      S.ep_len = r_ep;
      S.jump_flag = r_jf;
      S.level = r_lv;
      S.type = r_ty;
    }
    if (lane < 2) {
      S.cc_lo[lane] = P.cmd_lo[0];
      S.cc_span[lane] = P.cmd_span[0];
      if (P.command_curriculum) command_range_f32(B.command_ranges + 2 * lane, &S.cc_lo[lane], &S.cc_span[lane]);
    }
  }
  B200_WARP_SYNC();

  // ---- stage 1: height scan (env_warp_scan)
  if (do_scan) env_warp_scan<FIXED>(P, B, S, pt_x, pt_y, e, lane_lo, lane_hi);
}

// ---- stage 1: height scan, points strided over lanes (legged_robot.py:997-1032): all cells of the lane first, then
// all gathers (3 per point, independent), then the minima -- one round trip to the height field per lane.  Needs stage 0
// (the root state) of its env only.
template <bool FIXED, class SC>
B200_HD void env_warp_scan(const B200EnvParams& P, const B200EnvBuffers& B, SC& S, const float* pt_x, const float* pt_y, int e, int lane_lo,
                           int lane_hi) {
  B200_ENV_DIMS;
  B200_FOR_LANES(lane) {
    int n_out = 0;
    if (P.has_height_samples) {
      const YawQuat yq = yaw_quat(S.root + 3);
      const float inv_h = 1.0f / P.horizontal_scale;
      const int hs_pitch = hs_pitch_of(P);
      constexpr int kMaxPerLane = (B200_MAX_SCAN + 31) / 32;
      const int per_lane = FIXED ? (B200_GO2_SCAN_NX * B200_GO2_SCAN_NY + 31) / 32 : kMaxPerLane;
      int px[kMaxPerLane], py[kMaxPerLane];
      int16_t ha[kMaxPerLane], hb[kMaxPerLane], hc[kMaxPerLane];
#pragma unroll
      for (int k = 0; k < per_lane; ++k) {
        const int j = lane + 32 * k;
        const int jc = j < NS ? j : NS - 1;
        height_cell_pt(P, pt_x[jc], pt_y[jc], yq, S.root, inv_h, &px[k], &py[k]);
      }
#pragma unroll
      for (int k = 0; k < per_lane; ++k) {
        const int16_t* p = B.height_samples + (px[k] * hs_pitch + py[k]);    // rows * pitch < 2^31 (checked at env creation)
        ha[k] = B200_LDG(p);
        hb[k] = B200_LDG(p + hs_pitch);
        hc[k] = B200_LDG(p + 1);
      }
#pragma unroll
      for (int k = 0; k < per_lane; ++k) {
        const int j = lane + 32 * k;
        if (j < NS) {
          int16_t m = ha[k] < hb[k] ? ha[k] : hb[k];
          m = m < hc[k] ? m : hc[k];
          const float h = (float)m * P.vertical_scale;
          S.heights[j] = h;
          n_out += fabsf(h) > 0.1f;
          if (B.height_index) {
            B.height_index[((int64_t)e * NS + j) * 2] = px[k];
            B.height_index[((int64_t)e * NS + j) * 2 + 1] = py[k];
          }
        }
      }
    } else {
      for (int j = lane; j < NS; j += 32) S.heights[j] = 0.0f;
    }
    S.outliers[lane] = n_out;
  }
  B200_WARP_SYNC();
}

// ---- the history rows (go2.py:566-576): vector i of env e's [H x 52] history.
//   obs[e, 0:HN] = critic[e, 0:HN] = clip(history)   (zeros when the env reset: obs_history_buf[env_ids] = 0, go2.py:238)
//   history      = history shifted left by one slot   (unless it is refilled with the new row: env_warp_post)
// A caller that moves a row with several lanes must finish ALL its loads of the row before the first store (the shift
// is in place): the CUDA kernel puts a barrier between the two, the host emulation loads the whole row first.
template <bool FIXED>
B200_HD f4_ env_hist_load(const B200EnvParams& P, const B200EnvBuffers& B, int e, int i) {
  B200_ENV_DIMS;
  return reinterpret_cast<const f4_*>(B.obs_history_buf + (int64_t)e * HN)[i];
}
template <bool FIXED>
B200_HD void env_hist_store(const B200EnvParams& P, const B200EnvBuffers& B, int e, int i, f4_ v, int reset, int refill) {
  B200_ENV_DIMS;
  f4_ o = clamp4(v, P.clip_obs);
  if (reset) o.x = o.y = o.z = o.w = 0.0f;
  if (!P.alias_outputs) reinterpret_cast<f4_*>(B.obs_buf + (int64_t)e * OBS)[i] = o;
  reinterpret_cast<f4_*>(B.critic_obs_buf + (int64_t)e * CRIT)[i] = o;
  if (!refill && i >= NP4) reinterpret_cast<f4_*>(B.obs_history_buf + (int64_t)e * HN)[i - NP4] = v;
}

template <bool FIXED, class SC>
B200_HD void env_warp_post(const B200EnvParams& P, const B200EnvBuffers& B, const EnvTables& T, SC& S, int e, int64_t step64, int lane_lo,
                           int lane_hi) {
  B200_ENV_DIMS;

  // ---- stage 3: current observation + critic tail, elements strided over lanes
  B200_FOR_LANES(lane) {
    const float c = P.clip_obs;
    {
      // observation noise (go2.py:519): elements lane and lane + 32 take words (lane >> 4) and (lane >> 4) + 2 of ONE
      // Philox block (block = lane & 15) -- the keyed lane numbering of oracle/philox.py::noise_lane
      Philox4 r;
      if (P.add_noise) r = keyed_block(P.seed, SITE_OBS_NOISE, (uint32_t)step64, (uint32_t)e, (uint32_t)(lane & 15));
      for (int i = lane; i < NP; i += 32) {
        float u = 0.0f;
        if (P.add_noise) {
          const uint32_t w = (uint32_t)(i >> 4);
          u = u32_to_uniform(w == 0 ? r.v[0] : (w == 1 ? r.v[1] : (w == 2 ? r.v[2] : r.v[3])));
        }
        S.cur[i] = cur_obs_element(P, T, S, u, i);
      }
    }
    if (lane >= NPRIV) S.tail[lane] = clampf(S.blv[lane - NPRIV] * P.obs_lin_vel, -c, c);   // priv part: stage 0
    for (int j = lane; j < NS; j += 32)                    // go2.py:538, root z AFTER a possible reset
      S.tail[PE + j] = clampf((S.root_out[2] - 0.3f) - S.heights[j], -1.0f, 1.0f);
  }
  B200_WARP_SYNC();

  // ---- stage 4: write everything back (row-contiguous, lanes over consecutive 16-byte vectors)
  B200_FOR_LANES(lane) {
    const float c = P.clip_obs;
    const bool refill = S.ep_len_out <= 1;               // go2.py:570-574 (== S.early_refill)
    f4_* obs4 = reinterpret_cast<f4_*>(B.obs_buf + (int64_t)e * OBS);
    f4_* crit4 = reinterpret_cast<f4_*>(B.critic_obs_buf + (int64_t)e * CRIT);
    f4_* hist4 = reinterpret_cast<f4_*>(B.obs_history_buf + (int64_t)e * HN);
    const f4_* cur4 = reinterpret_cast<const f4_*>(S.cur);
    if (lane < NP4) {                                     // obs / critic end with clip(cur); history's newest slot is cur
      const f4_ v = cur4[lane];
      const f4_ o = clamp4(v, c);
      if (!P.alias_outputs) obs4[HN4 + lane] = o;
      crit4[HN4 + lane] = o;
      if (!refill) hist4[HN4 - NP4 + lane] = v;
    }
    if (refill) {                                         // cur x H right after a reset (go2.py:570-574)
      for (int i = lane; i < HN4; i += 32) hist4[i] = cur4[i % NP4];
    }
    const f4_* t4 = reinterpret_cast<const f4_*>(S.tail);
    for (int i = lane; i < TAIL4; i += 32) crit4[OBS4 + i] = t4[i];
    const f4_* s4 = reinterpret_cast<const f4_*>(S.tail + PE);
    const f4_* h4 = reinterpret_cast<const f4_*>(S.heights);
    for (int i = lane; i < NS4; i += 32) {
      if (!P.alias_outputs) reinterpret_cast<f4_*>(B.scan_obs_buf + (int64_t)e * NS)[i] = s4[i];
      reinterpret_cast<f4_*>(B.measured_heights + (int64_t)e * NS)[i] = h4[i];
    }
    if (!P.alias_outputs) {
      if (lane < NPRIV) B.privileged_obs_buf[(int64_t)e * NPRIV + lane] = S.tail[lane];
      else B.estimated_obs_buf[(int64_t)e * NEST + (lane - NPRIV)] = S.tail[lane];
    }
    // persistent state (go2.py:380-384 and the in-place updates of reset_idx)
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
    for (int k = lane; k < B200_NUM_REWARD_TERMS; k += 32) {   // episode_sums += term; zeroed on reset after the
      const float total = S.sums[k] + S.term[k];               // pre-zeroing value is parked for the extras means
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
    }
  }
}

// ---- reset_idx(arange(N)) outside a step (base_task.py:131-133): one env per call, lane-invariant ----
B200_HD void env_reset_only(const B200EnvParams& P, const B200EnvBuffers& B, int e, int64_t step64, int init_done) {
  ResetState R;
  for (int i = 0; i < 13; ++i) R.root[i] = B.root_states[(int64_t)e * 13 + i];
  for (int i = 0; i < 24; ++i) R.dof[i] = B.dof_state[(int64_t)e * 24 + i];
  for (int i = 0; i < 4; ++i) R.cmd[i] = B.commands[(int64_t)e * 4 + i];
  for (int i = 0; i < 3; ++i) R.origin[i] = B.env_origins[(int64_t)e * 3 + i];
  R.level = B.terrain_levels[e];
  R.ep_len = 0;
  R.cc_lo = P.cmd_lo[0];
  R.cc_span = P.cmd_span[0];
  if (P.command_curriculum) command_range_f32(B.command_ranges, &R.cc_lo, &R.cc_span);   // the range in force
  uint32_t draws[4 * B200_RESET_BLOCKS];
  for (int b = 0; b < B200_RESET_BLOCKS; ++b) env_reset_draw(P, draws, (uint32_t)e, (uint32_t)step64, b);
  reset_env(P, B, R, B.terrain_types[e], draws, init_done);
  for (int i = 0; i < 13; ++i) B.root_states[(int64_t)e * 13 + i] = R.root[i];
  for (int i = 0; i < 24; ++i) B.dof_state[(int64_t)e * 24 + i] = R.dof[i];
  for (int i = 0; i < 4; ++i) {
    B.commands[(int64_t)e * 4 + i] = R.cmd[i];
    B.last_contact_heights[(int64_t)e * 4 + i] = 0.0f;
    B.feet_air_time[(int64_t)e * 4 + i] = 0.0f;
    B.last_contacts[(int64_t)e * 4 + i] = 0;
  }
  for (int i = 0; i < 3; ++i) {
    B.env_origins[(int64_t)e * 3 + i] = R.origin[i];
    B.last_base_lin_vel[(int64_t)e * 3 + i] = 0.0f;
  }
  for (int i = 0; i < 12; ++i) {
    B.last_actions[(int64_t)e * 12 + i] = 0.0f;
    B.last_dof_vel[(int64_t)e * 12 + i] = 0.0f;
    B.last_torques[(int64_t)e * 12 + i] = 0.0f;
  }
  for (int i = 0; i < 6; ++i) B.last_root_vel[(int64_t)e * 6 + i] = 0.0f;
  const int HN = P.history_len * B200_PROPRIO;
  for (int i = 0; i < HN; ++i) B.obs_history_buf[(int64_t)e * HN + i] = 0.0f;
  for (int k = 0; k < B200_NUM_REWARD_TERMS; ++k) {
    B.reset_episode_sums[(int64_t)e * B200_NUM_REWARD_TERMS + k] = B.episode_sums[(int64_t)e * B200_NUM_REWARD_TERMS + k];
    B.episode_sums[(int64_t)e * B200_NUM_REWARD_TERMS + k] = 0.0f;
  }
  B.terrain_levels[e] = R.level;
  B.episode_length_buf[e] = 0;
  B.reset_buf[e] = 1;
}

// ---- env-creation-time randomisation of one env (legged_robot.py:306-380, :696-701, :897-930) ----
B200_HD void env_init_one(const B200EnvParams& P, const B200EnvBuffers& B, const B200InitParams& I, int e) {
  const uint32_t ue = (uint32_t)e;
  float fr = I.dynamic_friction;
  if (I.randomize_friction) {                            // 64 buckets ~ U(lo, hi); every env picks one
    const uint32_t bucket = keyed_u32(P.seed, SITE_INIT_FRICTION_PICK, 0, ue, 0) % 64u;
    fr = (I.friction_hi - I.friction_lo) * keyed_uniform(P.seed, SITE_INIT_FRICTION_BUCKET, 0, bucket, 0) + I.friction_lo;
  }
  const_cast<float*>(B.priv_friction)[e] = fr;
  float* mp = const_cast<float*>(B.priv_mass_params) + (int64_t)e * 4;
  mp[0] = I.randomize_base_mass ? (I.mass_hi - I.mass_lo) * keyed_uniform(P.seed, SITE_INIT_MASS, 0, ue, 0) + I.mass_lo : 0.0f;
  for (int i = 1; i < 4; ++i)
    mp[i] = I.randomize_com ? (I.com_hi - I.com_lo) * keyed_uniform(P.seed, SITE_INIT_MASS, 0, ue, i) + I.com_lo : 0.0f;
  float* kpkd = const_cast<float*>(B.kp_kd_multipliers);
  for (int d = 0; d < B200_NUM_DOF; ++d) {
    kpkd[(int64_t)e * B200_NUM_DOF + d] = (I.kp_kd_hi - I.kp_kd_lo) * keyed_uniform(P.seed, SITE_INIT_KPKD, 0, ue, d) + I.kp_kd_lo;
    kpkd[((int64_t)P.num_envs + e) * B200_NUM_DOF + d] =
        (I.kp_kd_hi - I.kp_kd_lo) * keyed_uniform(P.seed, SITE_INIT_KPKD, 0, ue, B200_NUM_DOF + d) + I.kp_kd_lo;
  }
  if (I.num_init_levels > 0) {                           // _get_env_origins with a height field (legged_robot.py:904-917)
    const int64_t level = (int64_t)(keyed_u32(P.seed, SITE_INIT_LEVEL, 0, ue, 0) % (uint32_t)I.num_init_levels);
    // torch.div(arange(N), N / num_cols, rounding_mode='floor') in fp32 (legged_robot.py:909): torch's floor division is
    // fmod-based -- (a - fmod(a, b)) / b, then floor with a half-ulp guard -- NOT floor(a / b): at an exact multiple of the
    // rounded-up fp32 divisor (env 1024 of 4096 with 20 columns: 1024 / 204.8000031) a / b rounds UP to 5.0 while the
    // reference assigns column 4.  Pinned to the reference's own terrain_types (tests/test_env_init.py).
    const float a_ = (float)e, b_ = (float)((double)P.num_envs / (double)I.terrain_cols);
    const float mod_ = fmodf(a_, b_);
    float div_ = (a_ - mod_) / b_;
    if (mod_ != 0.0f && ((b_ < 0.0f) != (mod_ < 0.0f))) div_ -= 1.0f;
    float fl_ = floorf(div_);
    if (div_ - fl_ > 0.5f) fl_ += 1.0f;
    const int64_t type = (int64_t)fl_;
    B.terrain_levels[e] = level;
    const_cast<int64_t*>(B.terrain_types)[e] = type;
    const float* o = B.terrain_origins + (level * I.terrain_cols + type) * 3;
    for (int i = 0; i < 3; ++i) B.env_origins[(int64_t)e * 3 + i] = B200_LDG(o + i);
  } else {                                               // plane: grid of env_spacing (legged_robot.py:919-930)
    B.env_origins[(int64_t)e * 3 + 0] = I.env_spacing * (float)(e / I.grid_cols);
    B.env_origins[(int64_t)e * 3 + 1] = I.env_spacing * (float)(e % I.grid_cols);
    B.env_origins[(int64_t)e * 3 + 2] = 0.0f;
    B.terrain_levels[e] = 0;
    const_cast<int64_t*>(B.terrain_types)[e] = 0;
  }
}

// ---- K1: action clip + PD torques, one (env, dof) element (legged_robot.py:74-75, :440-478) ----
B200_HD void pd_torque_element(const B200EnvParams& P, const B200EnvBuffers& B, const float* actions_in, int clip_and_store,
                               int64_t idx) {
  const int d = (int)(idx % B200_NUM_DOF);
  const int64_t e = idx / B200_NUM_DOF;
  float a;
  if (clip_and_store) {
    a = clampf(actions_in[idx], -P.clip_actions, P.clip_actions);
    B.actions[idx] = a;
  } else {
    a = B.actions[idx];
  }
  const float scaled = a * P.action_scale;
  const float pos = B.dof_state[2 * idx], vel = B.dof_state[2 * idx + 1];
  float tq;
  if (P.control_type == 0) {
    if (P.randomize_kp_kd) {
      const float kp = B200_LDG(B.kp_kd_multipliers + idx), kd = B200_LDG(B.kp_kd_multipliers + (int64_t)P.num_envs * B200_NUM_DOF + idx);
      tq = kp * P.p_gains[d] * ((scaled + P.default_dof_pos[d]) - pos) - kd * P.d_gains[d] * vel;
    } else {
      tq = P.p_gains[d] * ((scaled + P.default_dof_pos[d]) - pos) - P.d_gains[d] * vel;
    }
  } else if (P.control_type == 1) {
    tq = P.p_gains[d] * (scaled - vel) - P.d_gains[d] * (vel - B.last_dof_vel[idx]) / P.sim_dt;
  } else {
    tq = scaled;
  }
  B.torques[idx] = clampf(tq, -P.torque_limits[d], P.torque_limits[d]);
  (void)e;
}
