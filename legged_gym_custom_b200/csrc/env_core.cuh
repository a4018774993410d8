// One env step for ONE env, executed by one warp -- shared CUDA / host source.
//
// Replaces Go2Robot.post_physics_step and its callees (go2.py:345-387; full list in
// include/b200gym.h at b200_post_physics_step).  The routine is written as STAGES over
// lanes: inside a stage lanes are independent; lanes communicate only through the per-warp
// scratch (shared memory on the GPU) across stage boundaries.  On the GPU a lane is a
// thread and a boundary is __syncwarp(); tests/host_emul compiles the same source with g++
// and runs the lanes of a stage sequentially, which is how the kernel source itself is
// checked against the golden vectors without a GPU.
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

#define B200_TWO_PI_F 6.283185307179586f   /* fp32(2*np.pi) */
#define B200_PI_F 3.141592653589793f

struct float3_ {
  float x, y, z;
};

// ---- per-warp scratch -------------------------------------------------------------------
struct alignas(16) EnvScratch {
  // [history | cur] contiguous: obs_buf is a clipped copy of it, the new history is it shifted by one slot
  float histcur[B200_MAX_HIST + B200_MAX_PROPRIO];
  float tail[32 + 4 + B200_MAX_SCAN];   // priv | est | scan  = critic tail (16 B aligned pieces for 29+3+132)
  float heights[B200_MAX_SCAN];
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
  // element stage (lane-parallel): per-body contact tests, per-dof products, per-leg gait terms, the four angles
  int32_t body_hit[32];          // bit 0: |F| > 1 (termination test), bit 1: |F| > 0.1 (collision test)
  float dofv[10][12];            // per-dof contributions of the dof-summed reward terms (rows: DV_*)
  float leg_sin[4], leg_cos[4];  // sin / cos of 2 pi phase, contact order fl, fr, bl, br
  int32_t leg_stance[4];
  float ang[4];                  // roll, pitch, yaw, heading
  // results of the scalar stage
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
// point j of the x-major scan grid: ix = j / ny without an integer division ((2j+1)/(2ny) is never within 1/(2ny) of an
// integer, so the fp32 product truncates to the exact quotient for every j < 192, ny <= 24)
B200_HD void height_cell(const B200EnvParams& P, const float* scan_x, const float* scan_y, YawQuat yq, const float* root_pos, int j,
                         int* px, int* py) {
  const int gx = (int)((float)(2 * j + 1) * (0.5f / (float)P.scan_ny));
  const float vx = scan_x[gx], vy = scan_y[j - gx * P.scan_ny];
  // quat_apply((0,0,z,w), (vx,vy,0)): t = cross * 2; b + w*t + cross(xyz, t)
  const float t0 = -(yq.z * vy) * 2.0f, t1 = (yq.z * vx) * 2.0f;
  float rx = (vx + yq.w * t0) + (-(yq.z * t1));
  float ry = (vy + yq.w * t1) + (yq.z * t0);
  rx = (rx + root_pos[0]) + P.border_size;
  ry = (ry + root_pos[1]) + P.border_size;
  if (P.index_div_mode == 0) {
    rx = rx / P.horizontal_scale;
    ry = ry / P.horizontal_scale;
  } else {
    const float inv = 1.0f / P.horizontal_scale;
    rx = rx * inv;
    ry = ry * inv;
  }
  // .long() truncates toward zero; clamp first so the 32-bit conversion cannot overflow (the clip to [0, n-2] that
  // follows makes the result identical to the int64 path for every finite input)
  const float hx = (float)(P.hs_rows - 1), hy = (float)(P.hs_cols - 1);
  int ix = (int)(rx < -1.0f ? -1.0f : (rx > hx ? hx : rx)), iy = (int)(ry < -1.0f ? -1.0f : (ry > hy ? hy : ry));
  ix = ix < 0 ? 0 : (ix > P.hs_rows - 2 ? P.hs_rows - 2 : ix);
  iy = iy < 0 ? 0 : (iy > P.hs_cols - 2 ? P.hs_cols - 2 : iy);
  *px = ix;
  *py = iy;
}
B200_HD float height_at(const B200EnvParams& P, const int16_t* hs, int px, int py) {
  const int16_t* p = hs + (px * P.hs_cols + py);        // rows * cols < 2^31 (checked at env creation)
  const int16_t a = B200_LDG(p);
  const int16_t b = B200_LDG(p + P.hs_cols);
  const int16_t c = B200_LDG(p + 1);
  int16_t m = a < b ? a : b;
  m = m < c ? m : c;
  return (float)m * P.vertical_scale;
}

// ---- command resampling for one env (go2.py:413-464) -------------------------------------
B200_HD void resample_commands(const B200EnvParams& P, uint32_t site, uint32_t step, uint32_t e, const float* quat, float* cmd) {
  const Philox4 r = keyed_block(P.seed, site, step, e, 0);
  cmd[0] = P.cmd_span[0] * u32_to_uniform(r.v[0]) + P.cmd_lo[0];
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

// ---- reset of one env (go2.py:207-263 with legged_robot.py:481-574) ----------------------
// Operates on the caller's register copies (no shared-memory read-modify-write).
struct ResetState {
  float root[13], dof[24], cmd[4], origin[3], lch[4], fat[4];
  int32_t contact_cur[4];
  int64_t level, ep_len;
};
B200_HD void reset_env(const B200EnvParams& P, const B200EnvBuffers& B, ResetState& R, int64_t type, uint32_t step, uint32_t e,
                       int do_curriculum) {
  if (P.curriculum && do_curriculum) {                  // legged_robot.py:543-574
    const float dist = norm2_fma(R.root[0] - R.origin[0], R.root[1] - R.origin[1]);
    const int up = dist > P.promote_dist;
    const float expected = norm2_fma(R.cmd[0], R.cmd[1]) * P.max_episode_length_s;
    const int down = dist < expected * P.demote_threshold;
    int64_t lv = R.level + up - down;
    if (lv >= P.max_terrain_level)
      lv = (int64_t)(keyed_u32(P.seed, SITE_CURRICULUM, step, e, 0) % (uint32_t)P.max_terrain_level);
    else
      lv = lv < 0 ? 0 : lv;
    R.level = lv;
    const float* o = B.terrain_origins + (lv * P.terrain_cols + type) * 3;
    R.origin[0] = B200_LDG(o);
    R.origin[1] = B200_LDG(o + 1);
    R.origin[2] = B200_LDG(o + 2);
  }
  for (int d = 0; d < B200_NUM_DOF; ++d) {              // _reset_dofs
    const float u = keyed_uniform(P.seed, SITE_RESET_DOFS, step, e, d);
    R.dof[2 * d] = P.default_dof_pos[d] + (P.dof_reset_span * u + P.dof_reset_lo);
    R.dof[2 * d + 1] = 0.0f;
  }
  for (int i = 0; i < 13; ++i) R.root[i] = P.base_init_state[i];   // _reset_root_states
  for (int i = 0; i < 3; ++i) R.root[i] += R.origin[i];
  int lane0 = 0;
  if (P.custom_origins) {
    for (int i = 0; i < 2; ++i) R.root[i] += 2.0f * keyed_uniform(P.seed, SITE_RESET_ROOT, step, e, i) + -1.0f;
    lane0 = 2;
  }
  for (int i = 0; i < 6; ++i) R.root[7 + i] = 1.0f * keyed_uniform(P.seed, SITE_RESET_ROOT, step, e, lane0 + i) + -0.5f;
  resample_commands(P, SITE_CMD_RESET, step, e, R.root + 3, R.cmd);
  for (int f = 0; f < 4; ++f) {
    R.lch[f] = 0.0f;
    R.fat[f] = 0.0f;
    R.contact_cur[f] = 0;                                // last_contacts[env_ids] = 0 (go2.py:242)
  }
  R.ep_len = 0;
}

B200_HD void publish_reset_state(EnvScratch& S, const ResetState& R, int root_dirty, int dof_dirty, int reset) {
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
  S.root_dirty = root_dirty;
  S.dof_dirty = dof_dirty;
  S.reset = reset;
}

// ---- the element stage: everything that is "the same formula on 19 bodies / 12 dofs / 4 legs / 3 angles" runs with
// one lane per element instead of a serial loop in the scalar stage.  Contact tests compare the (FMA-accumulated, as
// torch does) squared norm with the pre-rounded squared threshold: sqrt_rn(s) > t  <=>  s > contact_thr2 exactly.
enum { DV_ACTION_RATE = 0, DV_DELTA_TORQUES, DV_DOF_ACC, DV_DQ2, DV_POS_LIMITS, DV_VEL2, DV_VEL_LIMITS, DV_ABS_DQ, DV_TORQUE_LIMITS, DV_TORQUES2 };

B200_HD float gait_phase(const B200EnvParams& P, int64_t ep) { return fmodf((float)ep * P.dt, P.period) / P.period; }

B200_HD void env_element_stage(const B200EnvParams& P, EnvScratch& S, int lane) {
  if (lane < B200_NUM_BODIES) {
    const float* c = S.contact + lane * 3;
    const float n2 = B200_FMA(c[2], c[2], B200_FMA(c[1], c[1], c[0] * c[0]));
    S.body_hit[lane] = (n2 > P.contact_thr2_term ? 1 : 0) | (n2 > P.contact_thr2_collision ? 2 : 0);
  }
  if (lane < B200_NUM_DOF) {
    const int d = lane;
    const float pos = S.dof[2 * d], vel = S.dof[2 * d + 1];
    const float dq = pos - P.default_dof_pos[d];
    S.dofv[DV_ACTION_RATE][d] = sq(S.last_act[d] - S.act[d]);
    S.dofv[DV_DELTA_TORQUES][d] = sq(S.tq[d] - S.last_tq[d]);
    S.dofv[DV_DOF_ACC][d] = sq((S.last_dv[d] - vel) / P.dt);
    S.dofv[DV_DQ2][d] = sq(dq);
    const float lo = pos - P.dof_pos_lo[d], hi = pos - P.dof_pos_hi[d];
    S.dofv[DV_POS_LIMITS][d] = -(lo > 0.0f ? 0.0f : lo) + (hi < 0.0f ? 0.0f : hi);
    S.dofv[DV_VEL2][d] = sq(vel);
    S.dofv[DV_VEL_LIMITS][d] = clampf(fabsf(vel) - P.dof_vel_limits[d] * P.soft_dof_vel_limit, 0.0f, 1.0f);
    S.dofv[DV_ABS_DQ][d] = fabsf(dq);
    const float over = fabsf(S.tq[d]) - P.torque_limits[d] * P.soft_torque_limit;
    S.dofv[DV_TORQUE_LIMITS][d] = over < 0.0f ? 0.0f : over;
    S.dofv[DV_TORQUES2][d] = sq(S.tq[d]);
  }
  if (lane >= 12 && lane < 16) {                          // gait phase of one leg (go2.py:279-290), order fl, fr, bl, br
    const int f = lane - 12;
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
  if (lane >= 16 && lane < 19) {                          // the three atan2 (roll, yaw, heading) share one code path
    const float x = S.root[3], y = S.root[4], z = S.root[5], w = S.root[6];
    float num, den;
    if (lane == 16) {                                     // roll (go2.py:19-21)
      num = 2.0f * (w * x + y * z);
      den = 1.0f - 2.0f * (x * x + y * y);
    } else if (lane == 17) {                              // yaw (go2.py:27-29)
      num = 2.0f * (w * z + x * y);
      den = 1.0f - 2.0f * (y * y + z * z);
    } else {                                              // heading = atan2(fwd.y, fwd.x), fwd = quat_apply(q, (1,0,0))
      const float t1 = z * 2.0f, t2 = -y * 2.0f;
      den = 1.0f + (y * t2 - z * t1);
      num = w * t1 + (-(x * t2));
    }
    S.ang[lane == 16 ? 0 : (lane == 17 ? 2 : 3)] = atan2f(num, den);
  }
  if (lane == 19) {                                       // pitch (go2.py:23-25)
    const float x = S.root[3], y = S.root[4], z = S.root[5], w = S.root[6];
    S.ang[1] = asinf(clampf(2.0f * (w * y - z * x), -1.0f, 1.0f));
  }
}

// ---- the scalar stage: everything between "state loaded" and "observations assembled" ---
// Pure function of the staged inputs: reads S.<in>, computes in registers, writes S.<out> once.
// Every lane computes the same values (one warp's issue slots, no shuffles, lane-invariant stores).
B200_HD void env_scalar_stage(const B200EnvParams& P, const B200EnvBuffers& B, EnvScratch& S, uint32_t e, int64_t step64) {
  const uint32_t step = (uint32_t)step64;
  const float* q = S.root + 3;
  const int64_t ep = S.ep_len + 1;                      // go2.py:354
  const float3_ blv = quat_rotate_inverse(q, S.root[7], S.root[8], S.root[9]);
  const float3_ bav = quat_rotate_inverse(q, S.root[10], S.root[11], S.root[12]);
  const float3_ pg = quat_rotate_inverse(q, 0.0f, 0.0f, -1.0f);
  S.blv[0] = blv.x; S.blv[1] = blv.y; S.blv[2] = blv.z;
  S.bav[0] = bav.x; S.bav[1] = bav.y; S.bav[2] = bav.z;
  S.pg[0] = pg.x; S.pg[1] = pg.y; S.pg[2] = pg.z;

  ResetState R;
  for (int i = 0; i < 13; ++i) R.root[i] = S.root[i];
  for (int i = 0; i < 24; ++i) R.dof[i] = S.dof[i];
  for (int i = 0; i < 4; ++i) R.cmd[i] = S.cmd[i];
  for (int i = 0; i < 3; ++i) R.origin[i] = S.origin[i];
  R.level = S.level;
  R.ep_len = ep;

  // update_feet_states (go2.py:266-328); leg order of contacts/feet: fl, fr, bl, br
  int filt[4];
  for (int f = 0; f < 4; ++f) {
    const int cur = S.contact[P.feet[f] * 3 + 2] > 1.0f;
    filt[f] = cur | (S.last_contacts[f] != 0);
    R.contact_cur[f] = cur;
    S.contact_filt[f] = filt[f];
    R.lch[f] = filt[f] ? S.feet_z[f] : S.lch[f];
    R.fat[f] = S.fat[f];
  }

  const float roll = S.ang[0], pitch = S.ang[1];        // quaternion_to_euler (go2.py:11-31), element stage
  S.rpy[0] = roll;
  S.rpy[1] = pitch;
  S.rpy[2] = S.ang[2];

  // _post_physics_step_callback (go2.py:390-410)
  float* cmd = R.cmd;
  if ((uint32_t)ep % (uint32_t)P.resample_interval == 0u) resample_commands(P, SITE_CMD_PERIODIC, step, e, q, cmd);
  const float heading = S.ang[3];
  if (P.heading_command) cmd[2] = clampf(wrap_to_pi(cmd[3] - heading) * P.heading_error_gain, -1.0f, 1.0f);
  int root_dirty = 0;
  if (P.push_robots && ((uint32_t)step64 % (uint32_t)P.push_interval) == 0u) {   // legged_robot.py:535-540
    const float span = P.max_push_vel - -P.max_push_vel;
    R.root[7] = span * keyed_uniform(P.seed, SITE_PUSH, step, e, 0) + -P.max_push_vel;
    R.root[8] = span * keyed_uniform(P.seed, SITE_PUSH, step, e, 1) + -P.max_push_vel;
    root_dirty = 1;
  }
  const float* root = R.root;                           // world-frame rewards see the push (Appendix C.7)

  // check_termination (go2.py:186-204)
  int reset = 0;
  for (int i = 0; i < P.n_termination; ++i) reset |= S.body_hit[P.termination[i]] & 1;
  const int time_out = ep > P.max_episode_length;
  reset |= time_out;
  reset |= pg.z > 0.0f;
  if (P.parkour) reset |= root[2] < -1.0f;
  S.time_out = time_out;

  // compute_reward (legged_robot.py:216-237): alphabetical accumulation, term * (scale*dt)
  const float* sc = P.reward_scales;
  float rew = 0.0f;
  const float cmd_n3 = norm3_fma(cmd[0], cmd[1], cmd[2]);
  const float moving = cmd_n3 >= 0.2f ? 1.0f : 0.0f;
  const float jumping = S.jump_flag > 0.0f ? 1.0f : 0.0f;
  const int stance[4] = {S.leg_stance[0], S.leg_stance[1], S.leg_stance[2], S.leg_stance[3]};
#define B200_DOFSUM(ROW, OUT)                                   \
  {                                                             \
    OUT = 0.0f;                                                 \
    for (int d_ = 0; d_ < 12; ++d_) OUT += S.dofv[ROW][d_];      \
  }
#define B200_TERM(NAME, EXPR)                         \
  {                                                   \
    float r_ = 0.0f;                                  \
    if (sc[B200_REW_##NAME] != 0.0f) {                \
      r_ = (EXPR) * sc[B200_REW_##NAME];              \
      rew += r_;                                      \
    }                                                 \
    S.term[B200_REW_##NAME] = r_;                     \
  }
  {
    float a = 0.0f;
    if (sc[B200_REW_action_rate] != 0.0f) B200_DOFSUM(DV_ACTION_RATE, a)
    B200_TERM(action_rate, a)
  }
  B200_TERM(ang_vel_xy, sq(bav.x) + sq(bav.y))
  {
    float a = 0.0f;
    if (sc[B200_REW_base_height] != 0.0f) {
      for (int j = 0; j < P.num_scan; ++j) a += root[2] - S.heights[j];
      a = sq(a / (float)P.num_scan - P.base_height_target);
    }
    B200_TERM(base_height, a)
  }
  {
    float a = 0.0f;
    if (sc[B200_REW_calf_collision] != 0.0f)
      for (int f = 0; f < 4; ++f) a += (S.body_hit[P.calves[f]] & 2) ? 1.0f : 0.0f;
    B200_TERM(calf_collision, a)
  }
  {
    float a = 0.0f;
    if (sc[B200_REW_calf_pos] != 0.0f)
      for (int f = 0; f < 4; ++f) a += S.dofv[DV_DQ2][P.calf_joints[f]];
    B200_TERM(calf_pos, a)
  }
  B200_TERM(calf_symmetry, fabsf(S.dof[2 * P.calf_joints[0]] - S.dof[2 * P.calf_joints[1]]) +
                               fabsf(S.dof[2 * P.calf_joints[2]] - S.dof[2 * P.calf_joints[3]]))
  {
    float a = 0.0f;
    if (sc[B200_REW_collision] != 0.0f)
      for (int i = 0; i < P.n_penalised; ++i) a += (S.body_hit[P.penalised[i]] & 2) ? 1.0f : 0.0f;
    B200_TERM(collision, a)
  }
  {
    float a = 0.0f;
    if (sc[B200_REW_delta_torques] != 0.0f) B200_DOFSUM(DV_DELTA_TORQUES, a)
    B200_TERM(delta_torques, a)
  }
  {
    float a = 0.0f;
    if (sc[B200_REW_dof_acc] != 0.0f) B200_DOFSUM(DV_DOF_ACC, a)
    B200_TERM(dof_acc, a)
  }
  float dof_err;
  B200_DOFSUM(DV_DQ2, dof_err)
  B200_TERM(dof_error, dof_err)
  {
    float a = 0.0f;
    if (sc[B200_REW_dof_pos_limits] != 0.0f) B200_DOFSUM(DV_POS_LIMITS, a)
    B200_TERM(dof_pos_limits, a)
  }
  {
    float a = 0.0f;
    if (sc[B200_REW_dof_vel] != 0.0f) B200_DOFSUM(DV_VEL2, a)
    B200_TERM(dof_vel, a)
  }
  {
    float a = 0.0f;
    if (sc[B200_REW_dof_vel_limits] != 0.0f) B200_DOFSUM(DV_VEL_LIMITS, a)
    B200_TERM(dof_vel_limits, a)
  }
  {                                                     // go2.py:819-831 (stateful)
    float a = 0.0f;
    if (sc[B200_REW_feet_air_time] != 0.0f) {
      // update_feet_states has already overwritten last_contacts with the CURRENT contacts (go2.py:307-310), so the
      // "filtered" contact of this reward (go2.py:824-825) is just the current one
      for (int f = 0; f < 4; ++f) {
        const int now = R.contact_cur[f];
        const float first = (S.fat[f] > 0.0f && now) ? 1.0f : 0.0f;
        const float t = S.fat[f] + P.dt;
        a += (t - 0.5f) * first;
        R.fat[f] = t * (now ? 0.0f : 1.0f);
      }
      a *= norm2_fma(cmd[0], cmd[1]) > 0.1f ? 1.0f : 0.0f;
    }
    B200_TERM(feet_air_time, a)
  }
  {
    float a = 0.0f;
    if (sc[B200_REW_feet_contact_forces] != 0.0f)
      for (int f = 0; f < 4; ++f) {
        const float* c = S.contact + P.feet[f] * 3;
        const float over = norm3_fma(c[0], c[1], c[2]) - P.max_contact_force;
        a += over < 0.0f ? 0.0f : over;
      }
    B200_TERM(feet_contact_forces, a)
  }
  {                                                     // go2.py:734-756; wrap_to_pi mutates commands[:,3]
    float a = 0.0f;
    if (sc[B200_REW_heading_alignment] != 0.0f) {
      float desired = 0.0f;
      if (P.heading_command) {
        cmd[3] = wrap_to_pi(cmd[3]);
        desired = cmd[3];
      }
      a = sq(wrap_to_pi(desired - heading)) * moving;
    }
    B200_TERM(heading_alignment, a)
  }
  {
    float a = 0.0f;
    if (sc[B200_REW_hip_pos] != 0.0f)
      for (int f = 0; f < 4; ++f) a += S.dofv[DV_DQ2][P.hip_joints[f]];
    B200_TERM(hip_pos, a)
  }
  B200_TERM(jump_zone_forward_vel, (root[7] < 0.0f ? 0.0f : root[7]) * jumping * moving)
  B200_TERM(jump_zone_upward_vel, (root[9] < 0.0f ? 0.0f : root[9]) * jumping * moving)
  B200_TERM(lin_vel_z, sq(blv.z))
  B200_TERM(min_height, clampf(P.base_height_target - root[2], 0.0f, P.base_height_target) * jumping)
  B200_TERM(orientation, sq(pg.x) + sq(pg.y))
  {                                                     // go2.py:621-644
    float a = 0.0f;
    for (int f = 0; f < 4; ++f) a += (filt[f] == stance[f]) ? 0.25f : -0.25f;
    B200_TERM(phase_contact_match, a)
  }
  {                                                     // go2.py:647-678
    float a = 0.0f;
    for (int f = 0; f < 4; ++f) {
      const float h = clampf(S.feet_z[f] - R.lch[f], 0.0f, P.max_foot_height) / P.max_foot_height;
      a += stance[f] ? -h : h;
    }
    B200_TERM(phase_foot_lifting, a / 2.0f)
  }
  B200_TERM(reverse_penalty, -(root[7] > 0.0f ? 0.0f : root[7]))
  {
    float a = 0.0f;
    if (sc[B200_REW_stand_still] != 0.0f) {
      B200_DOFSUM(DV_ABS_DQ, a)
      a *= norm2_fma(cmd[0], cmd[1]) < 0.1f ? 1.0f : 0.0f;
    }
    B200_TERM(stand_still, a)
  }
  {
    int any = 0;
    if (sc[B200_REW_stumble_calves] != 0.0f)
      for (int f = 0; f < 4; ++f) {
        const float* c = S.contact + P.calves[f] * 3;
        any |= norm2_fma(c[0], c[1]) > 5.0f * fabsf(c[2]);
      }
    B200_TERM(stumble_calves, any ? 1.0f : 0.0f)
  }
  {
    int any = 0;
    if (sc[B200_REW_stumble_feet] != 0.0f)
      for (int f = 0; f < 4; ++f) {
        const float* c = S.contact + P.feet[f] * 3;
        any |= norm2_fma(c[0], c[1]) > 5.0f * fabsf(c[2]);
      }
    B200_TERM(stumble_feet, any ? 1.0f : 0.0f)
  }
  {
    float a = 0.0f;
    if (sc[B200_REW_thigh_pos] != 0.0f)
      for (int f = 0; f < 4; ++f) a += S.dofv[DV_DQ2][P.thigh_joints[f]];
    B200_TERM(thigh_pos, a)
  }
  B200_TERM(thigh_symmetry, fabsf(S.dof[2 * P.thigh_joints[0]] - S.dof[2 * P.thigh_joints[1]]) +
                                fabsf(S.dof[2 * P.thigh_joints[2]] - S.dof[2 * P.thigh_joints[3]]))
  {
    float a = 0.0f;
    if (sc[B200_REW_torque_limits] != 0.0f) B200_DOFSUM(DV_TORQUE_LIMITS, a)
    B200_TERM(torque_limits, a)
  }
  {
    float a = 0.0f;
    if (sc[B200_REW_torques] != 0.0f) B200_DOFSUM(DV_TORQUES2, a)
    B200_TERM(torques, a)
  }
  B200_TERM(tracking_ang_vel, expf(-sq(cmd[2] - bav.z) / P.tracking_sigma))
  B200_TERM(tracking_lin_vel, expf(-(sq(cmd[0] - blv.x) + sq(cmd[1] - blv.y)) / P.tracking_sigma))
  B200_TERM(tracking_pitch, expf(-sq(pitch * 57.29577951308232f - P.pitch_deg_target) / P.tracking_sigma))
  B200_TERM(tracking_roll, expf(-sq(roll * 57.29577951308232f - P.roll_deg_target) / P.tracking_sigma))
  B200_TERM(zero_cmd_dof_error, dof_err * (cmd_n3 < 0.2f ? 1.0f : 0.0f))
  if (P.only_positive_rewards) rew = rew < 0.0f ? 0.0f : rew;
  {                                                     // legged_robot.py:234-237
    float r_ = 0.0f;
    if (sc[B200_REW_termination] != 0.0f) {
      r_ = ((reset && !time_out) ? 1.0f : 0.0f) * sc[B200_REW_termination];
      rew += r_;
    }
    S.term[B200_REW_termination] = r_;
  }
#undef B200_TERM
#undef B200_DOFSUM
  S.rew = rew;

  // reset_idx on this env if flagged (go2.py:375-376)
  int dof_dirty = 0;
  if (reset) {
    reset_env(P, B, R, S.type, step, e, 1);
    root_dirty = 1;
    dof_dirty = 1;
  }
  publish_reset_state(S, R, root_dirty, dof_dirty, reset);

  // jump flags for the NEXT step's rewards (go2.py:487-494), from this step's heights
  float jf = S.jump_flag;
  if (P.parkour) {
    int n = 0;
    for (int l = 0; l < 32; ++l) n += S.outliers[l];
    jf = n >= 8 ? 1.0f : 0.0f;
  }
  S.jump_flag_out = jf;
}

// ---- one element of cur_obs (go2.py:506-519) ----------------------------------------------
B200_HD float cur_obs_element(const B200EnvParams& P, const EnvScratch& S, float u, int i) {
  float v;
  if (i < 3) v = S.bav[i] * P.obs_ang_vel;
  else if (i < 5) v = S.rpy[i - 3];
  else if (i < 8) v = S.cmd_out[i - 5] * (i < 7 ? P.obs_lin_vel : P.obs_ang_vel);
  else if (i < 20) v = (S.dof_out[2 * (i - 8)] - P.default_dof_pos[i - 8]) * P.obs_dof_pos;
  else if (i < 32) v = S.dof_out[2 * (i - 20) + 1] * P.obs_dof_vel;
  else if (i < 44) v = S.act[i - 32];
  else {                                                // sin/cos of fr, fl, bl, br (go2.py:476-481); scratch order fl, fr, bl, br
    const int leg = (i - 44) >> 1;
    const int f = leg == 0 ? 1 : (leg == 1 ? 0 : leg);
    v = ((i - 44) & 1) ? S.leg_cos[f] : S.leg_sin[f];
  }
  if (P.add_noise) v += (2.0f * u - 1.0f) * P.noise_vec[i];
  return v;
}

B200_HD f4_ clamp4(f4_ v, float c) {
  v.x = clampf(v.x, -c, c);
  v.y = clampf(v.y, -c, c);
  v.z = clampf(v.z, -c, c);
  v.w = clampf(v.w, -c, c);
  return v;
}

// ---- the whole env step for env `e`, lanes [lane_lo, lane_hi) -------------------------------
// Row sizes are multiples of 4 floats for the go2 layout (52, 520, 572, 736, 132, critic tail 164),
// so the bulk rows move as 16-byte vectors; `vec_ok` (warp-uniform) falls back to scalars otherwise.
// The env step of env `e` is three calls:
//   env_warp_pre    one warp, lanes [lane_lo, lane_hi): stage 0 (load), 1 (height scan), 2a (element stage)
//   env_scalar_stage one THREAD per env (the CUDA kernel runs it on warp 0 of the CTA, lane = env slot)
//   env_warp_post   one warp: stage 3 (observation assembly), 4 (write-back)
// scan_x / scan_y: the scan-point tables (shared-memory copies on the GPU, P.scan_x / P.scan_y on the host)
#define B200_ENV_DIMS                                                      \
  const int NP = B200_PROPRIO, H = P.history_len, NS = P.num_scan;         \
  const int HN = H * NP, OBS = HN + NP;                                    \
  const int TAIL = P.num_priv + P.num_est + NS;                            \
  const int CRIT = OBS + TAIL;                                             \
  const int64_t N = P.num_envs;                                            \
  const bool vec_ok = (HN % 4 == 0) && (TAIL % 4 == 0) && (NS % 4 == 0) && ((P.num_priv + P.num_est) % 4 == 0); \
  float* hist = B.obs_history_buf + (int64_t)e * HN;                       \
  (void)OBS; (void)CRIT; (void)N; (void)TAIL

B200_HD void env_warp_pre(const B200EnvParams& P, const B200EnvBuffers& B, EnvScratch& S, const float* scan_x, const float* scan_y,
                          int e, int lane_lo, int lane_hi) {
  B200_ENV_DIMS;

  // ---- stage 0: stage the env's rows into scratch (coalesced: consecutive lanes, consecutive floats)
  B200_FOR_LANES(lane) {
    if (vec_ok) {
      const f4_* src = reinterpret_cast<const f4_*>(hist);
      f4_* dst = reinterpret_cast<f4_*>(S.histcur);
      for (int i = lane; i < HN / 4; i += 32) dst[i] = src[i];
    } else {
      for (int i = lane; i < HN; i += 32) S.histcur[i] = hist[i];
    }
    if (lane < 13) S.root[lane] = B.root_states[(int64_t)e * 13 + lane];
    if (lane < 24) S.dof[lane] = B.dof_state[(int64_t)e * 24 + lane];
    for (int i = lane; i < B200_NUM_BODIES * 3; i += 32) S.contact[i] = B200_LDG(B.contact_forces + (int64_t)e * B200_NUM_BODIES * 3 + i);
    if (lane < 4) {
      S.feet_z[lane] = B200_LDG(B.rigid_body_states + ((int64_t)e * B200_NUM_BODIES + P.feet[lane]) * 13 + 2);
      S.cmd[lane] = B.commands[(int64_t)e * 4 + lane];
      S.lch[lane] = B.last_contact_heights[(int64_t)e * 4 + lane];
      S.fat[lane] = B.feet_air_time[(int64_t)e * 4 + lane];
      S.last_contacts[lane] = B.last_contacts[(int64_t)e * 4 + lane];
    }
    if (lane < 12) {
      S.act[lane] = B.actions[(int64_t)e * 12 + lane];
      S.tq[lane] = B.torques[(int64_t)e * 12 + lane];
      S.last_act[lane] = B.last_actions[(int64_t)e * 12 + lane];
      S.last_dv[lane] = B.last_dof_vel[(int64_t)e * 12 + lane];
      S.last_tq[lane] = B.last_torques[(int64_t)e * 12 + lane];
    }
    for (int k = lane; k < B200_NUM_REWARD_TERMS; k += 32) S.sums[k] = B.episode_sums[(int64_t)e * B200_NUM_REWARD_TERMS + k];
    if (lane < 3) S.origin[lane] = B.env_origins[(int64_t)e * 3 + lane];
    if (lane == 0) {
      S.ep_len = B.episode_length_buf[e];
      S.jump_flag = B.jump_flags[e];
      S.level = B.terrain_levels[e];
      S.type = B200_LDG(B.terrain_types + e);
    }
  }
  B200_WARP_SYNC();

  // ---- stage 1: height scan, points strided over lanes (legged_robot.py:997-1032)
  B200_FOR_LANES(lane) {
    int n_out = 0;
    if (P.has_height_samples) {
      const YawQuat yq = yaw_quat(S.root + 3);
      for (int j = lane; j < NS; j += 32) {
        int px, py;
        height_cell(P, scan_x, scan_y, yq, S.root, j, &px, &py);
        const float h = height_at(P, B.height_samples, px, py);
        S.heights[j] = h;
        n_out += fabsf(h) > 0.1f;
        if (B.height_index) {
          B.height_index[((int64_t)e * NS + j) * 2] = px;
          B.height_index[((int64_t)e * NS + j) * 2 + 1] = py;
        }
      }
    } else {
      for (int j = lane; j < NS; j += 32) S.heights[j] = 0.0f;
    }
    S.outliers[lane] = n_out;
  }
  B200_WARP_SYNC();

  // ---- stage 2a: element stage (one lane per body / dof / leg / angle)
  B200_FOR_LANES(lane) { env_element_stage(P, S, lane); }
  B200_WARP_SYNC();

}

B200_HD void env_warp_post(const B200EnvParams& P, const B200EnvBuffers& B, EnvScratch& S, int e, int64_t step64, int lane_lo, int lane_hi) {
  B200_ENV_DIMS;

  // ---- stage 3: current observation + critic tail, elements strided over lanes
  B200_FOR_LANES(lane) {
    const float c = P.clip_obs;
    if (S.reset) {                                        // obs_history_buf[env_ids] = 0 (go2.py:238)
      for (int i = lane; i < HN; i += 32) S.histcur[i] = 0.0f;
    }
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
        S.histcur[HN + i] = cur_obs_element(P, S, u, i);
      }
    }
    for (int i = lane; i < P.num_priv; i += 32) {         // go2.py:528-532
      float v;
      if (i < 4) v = B200_LDG(B.priv_mass_params + (int64_t)e * 4 + i);
      else if (i < 5) v = B200_LDG(B.priv_friction + e);
      else if (i < 17) v = B200_LDG(B.kp_kd_multipliers + (int64_t)e * 12 + (i - 5)) - 1.0f;
      else v = B200_LDG(B.kp_kd_multipliers + (N + e) * 12 + (i - 17)) - 1.0f;
      S.tail[i] = clampf(v, -c, c);
    }
    for (int i = lane; i < P.num_est; i += 32) S.tail[P.num_priv + i] = clampf(S.blv[i] * P.obs_lin_vel, -c, c);
    for (int j = lane; j < NS; j += 32)                    // go2.py:538, root z AFTER a possible reset
      S.tail[P.num_priv + P.num_est + j] = clampf((S.root_out[2] - 0.3f) - S.heights[j], -1.0f, 1.0f);
  }
  B200_WARP_SYNC();

  // ---- stage 4: write everything back (row-contiguous, lanes over consecutive 16-byte vectors)
  B200_FOR_LANES(lane) {
    const float c = P.clip_obs;
    const bool refill = S.ep_len_out <= 1;               // go2.py:570-574
    float* obs = B.obs_buf + (int64_t)e * OBS;
    float* crit = B.critic_obs_buf + (int64_t)e * CRIT;
    if (vec_ok) {
      const f4_* hc = reinterpret_cast<const f4_*>(S.histcur);
      f4_* obs4 = reinterpret_cast<f4_*>(obs);
      f4_* crit4 = reinterpret_cast<f4_*>(crit);
      f4_* hist4 = reinterpret_cast<f4_*>(hist);
      for (int i = lane; i < OBS / 4; i += 32) {          // obs = clip([history | cur]); critic starts with it
        const f4_ v = clamp4(hc[i], c);
        obs4[i] = v;
        crit4[i] = v;
      }
      if (!refill) {                                      // history <- shift left one slot ...
        for (int i = lane; i < HN / 4; i += 32) hist4[i] = hc[i + NP / 4];
      } else {                                            // ... or cur x H right after a reset (go2.py:570-574)
        for (int i = lane; i < HN / 4; i += 32) hist4[i] = hc[HN / 4 + i % (NP / 4)];
      }
      const f4_* t4 = reinterpret_cast<const f4_*>(S.tail);
      for (int i = lane; i < TAIL / 4; i += 32) crit4[OBS / 4 + i] = t4[i];
      const f4_* s4 = reinterpret_cast<const f4_*>(S.tail + P.num_priv + P.num_est);
      const f4_* h4 = reinterpret_cast<const f4_*>(S.heights);
      for (int i = lane; i < NS / 4; i += 32) {
        reinterpret_cast<f4_*>(B.scan_obs_buf + (int64_t)e * NS)[i] = s4[i];
        reinterpret_cast<f4_*>(B.measured_heights + (int64_t)e * NS)[i] = h4[i];
      }
    } else {
      for (int i = lane; i < OBS; i += 32) {
        const float v = clampf(S.histcur[i], -c, c);
        obs[i] = v;
        crit[i] = v;
      }
      for (int i = lane; i < HN; i += 32) hist[i] = refill ? S.histcur[HN + i % NP] : S.histcur[i + NP];
      for (int i = lane; i < TAIL; i += 32) crit[OBS + i] = S.tail[i];
      for (int j = lane; j < NS; j += 32) {
        B.scan_obs_buf[(int64_t)e * NS + j] = S.tail[P.num_priv + P.num_est + j];
        B.measured_heights[(int64_t)e * NS + j] = S.heights[j];
      }
    }
    for (int i = lane; i < P.num_priv; i += 32) B.privileged_obs_buf[(int64_t)e * P.num_priv + i] = S.tail[i];
    for (int i = lane; i < P.num_est; i += 32) B.estimated_obs_buf[(int64_t)e * P.num_est + i] = S.tail[P.num_priv + i];
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
  reset_env(P, B, R, B.terrain_types[e], (uint32_t)step64, (uint32_t)e, init_done);
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
