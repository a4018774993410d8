// Keyed Philox4x32-10 uniforms -- CUDA/host twin of oracle/philox.py (same keying, same
// bits).  counter = (env, step, site, lane >> 2), key = (seed_lo, seed_hi), word = lane & 3;
// u = (word >> 8) * 2^-24 in [0,1).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

enum B200DrawSite {
  SITE_CMD_PERIODIC = 0,  // go2.py:393-396 -> _resample_commands; lanes vx, vy, heading|yaw, zero-mask
  SITE_PUSH = 1,          // legged_robot.py:539; lanes x, y
  SITE_CURRICULUM = 2,    // legged_robot.py:572 randint_like; lane 0 (raw u32 % max_level)
  SITE_RESET_DOFS = 3,    // legged_robot.py:491; lanes 0..11
  SITE_RESET_ROOT = 4,    // legged_robot.py:520 (xy, only with custom origins) then :526 (6 velocities)
  SITE_CMD_RESET = 5,     // go2.py:230 -> _resample_commands
  SITE_OBS_NOISE = 6,     // go2.py:519 rand_like; lanes 0..num_proprio-1
  SITE_ACTION_NOISE = 7,  // actor_critic.py:204 Normal.sample; lanes 2a, 2a+1
  // env-creation-time draws (step = 0)
  SITE_INIT_FRICTION_BUCKET = 8,   // legged_robot.py:318-320: env = bucket id (64 buckets), lane 0
  SITE_INIT_FRICTION_PICK = 9,     // legged_robot.py:319 randint(0, 64): lane 0 (raw u32 % 64)
  SITE_INIT_MASS = 10,             // legged_robot.py:363-372: lane 0 added mass, lanes 1..3 centre-of-mass shift
  SITE_INIT_KPKD = 11,             // legged_robot.py:696-701: lanes 0..11 kp, 12..23 kd
  SITE_INIT_LEVEL = 12             // legged_robot.py:909 randint(0, max_init_level + 1): lane 0 (raw u32 % count)
};

struct Philox4 {
  uint32_t v[4];
};

B200_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

B200_HD Philox4 keyed_block(uint64_t seed, uint32_t site, uint32_t step, uint32_t env, uint32_t block) {
  return philox4x32_10(env, step, site, block, (uint32_t)seed, (uint32_t)(seed >> 32));
}

B200_HD float u32_to_uniform(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }

B200_HD uint32_t keyed_u32(uint64_t seed, uint32_t site, uint32_t step, uint32_t env, uint32_t lane) {
  const Philox4 b = keyed_block(seed, site, step, env, lane >> 2);
  const uint32_t w = lane & 3u;
  return w == 0 ? b.v[0] : (w == 1 ? b.v[1] : (w == 2 ? b.v[2] : b.v[3]));
}

B200_HD float keyed_uniform(uint64_t seed, uint32_t site, uint32_t step, uint32_t env, uint32_t lane) {
  return u32_to_uniform(keyed_u32(seed, site, step, env, lane));
}
