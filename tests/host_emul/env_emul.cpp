// Host emulation of the env kernels: compiles legged_gym_custom_b200/csrc/env_core.cuh -- the
// SAME source the CUDA kernels execute -- with g++ (-ffp-contract=off mirrors -fmad=false) and
// runs each warp's lanes sequentially stage by stage.  TEST INFRASTRUCTURE: lets the CPU test
// suite check the kernel source against the golden vectors without a GPU.  Not a product path
// (the product library has no host fallback).
#include <string.h>

#include "../../legged_gym_custom_b200/csrc/env_core.cuh"

extern "C" {

int emul_params_size(void) { return (int)sizeof(B200EnvParams); }
int emul_buffers_size(void) { return (int)sizeof(B200EnvBuffers); }
int emul_scratch_size(void) { return (int)sizeof(EnvScratch); }

int emul_pd_torques(const B200EnvParams* p, const B200EnvBuffers* b, const float* actions_in, int clip_and_store) {
  const int64_t n = (int64_t)p->num_envs * B200_NUM_DOF;
  for (int64_t i = 0; i < n; ++i) pd_torque_element(*p, *b, actions_in, clip_and_store, i);
  return 0;
}

static void extras(const B200EnvParams& P, const B200EnvBuffers& B) {
  const int T = B200_NUM_REWARD_TERMS, N = P.num_envs;
  int count = 0;
  for (int e = 0; e < N; ++e) count += B.reset_buf[e] != 0;
  B.reset_count[0] = count;
  if (count == 0) return;
  for (int k = 0; k < T; ++k) {
    if (P.reward_scales[k] == 0.0f) continue;
    float s = 0.0f;
    for (int e = 0; e < N; ++e)
      if (B.reset_buf[e]) s += B.reset_episode_sums[(int64_t)e * T + k];
    B.extras_episode[k] = (s / (float)count) / P.max_episode_length_s;
  }
  if (P.curriculum) {
    float s = 0.0f;
    for (int e = 0; e < N; ++e) s += (float)B.terrain_levels[e];
    B.extras_episode[T] = s / (float)N;
  }
  for (int e = 0; e < N; ++e) B.extras_time_outs[e] = B.time_out_buf[e];
}

}  // extern "C"

// one env step, stage by stage, in the order the CUDA kernel's barriers allow: the history rows move between the
// item stage and the write-back (all loads of a row before its first store, like the kernel's barrier between B1 and B2)
template <bool FIXED>
static void step_all(const B200EnvParams& P, const B200EnvBuffers& B, int64_t step) {
  static EnvScratch S;
  static f4_ row[B200_MAX_HIST / 4];
  static EnvTables T;
  static float pt_x[B200_MAX_SCAN], pt_y[B200_MAX_SCAN];
  for (int j = 0; j < P.num_scan; ++j) scan_point(P, j, &pt_x[j], &pt_y[j]);
  for (int i = 0; i < B200_MAX_PROPRIO; ++i) env_tables_fill(P, T, i);
  const int hn4 = P.history_len * (B200_PROPRIO / 4);
  if (P.command_curriculum) {     // launch_command_curriculum (env_kernels.cu): probe pass, then the range update
    double* cr = B.command_ranges;
    cr[0] = cr[2];
    cr[1] = cr[3];
    if (step % P.max_episode_length == 0) {
      for (int e = 0; e < P.num_envs; ++e) {
        memset(&S, 0xCD, sizeof(S));
        env_warp_pre<FIXED>(P, B, S, pt_x, pt_y, e, 0, 32);
        for (int it = 0; it < ITEM_COUNT; ++it) env_item(P, T, S, it, (uint32_t)e, step);
        env_cc_probe<FIXED>(P, B, S, e);
      }
      int count = 0;
      double sum = 0.0;
      for (int e = 0; e < P.num_envs; ++e)
        if (B.cc_reset[e]) {
          ++count;
          sum += (double)B.cc_value[e];
        }
      command_curriculum_rule(P, count, sum, cr + 2, cr + 2);
    }
  }
  for (int e = 0; e < P.num_envs; ++e) {
    memset(&S, 0xCD, sizeof(S));   // poison: a stage that reads what no stage wrote shows up as garbage
    env_warp_pre<FIXED>(P, B, S, pt_x, pt_y, e, 0, 32);
    for (int it = ITEM_COUNT - 1; it >= 0; --it) env_item(P, T, S, it, (uint32_t)e, step);   // order-free (here: reversed)
    for (int i = 0; i < hn4; ++i) row[i] = env_hist_load<FIXED>(P, B, e, i);
    for (int i = 0; i < hn4; ++i) env_hist_store<FIXED>(P, B, e, i, row[i], S.early_reset, S.early_refill);
    for (int part = B200_TERM_PARTS - 1; part >= 0; --part) env_terms_part<FIXED>(P, S, part);
    if (S.early_reset)
      for (int b = 0; b < B200_RESET_BLOCKS; ++b) env_reset_draw(P, S.reset_draws, (uint32_t)e, (uint32_t)step, b);
    env_finalize(P, B, S);
    env_warp_post<FIXED>(P, B, T, S, e, step, 0, 32);
  }
}

extern "C" {

// variant: 0 = pick like the library does (layout baked in for the go2 layout), 1 = force the layout-generic code
int emul_post_physics_step_variant(const B200EnvParams* p, const B200EnvBuffers* b, int64_t step, int force_generic) {
  if (env_layout_is_go2(*p) && !force_generic) step_all<true>(*p, *b, step);
  else step_all<false>(*p, *b, step);
  extras(*p, *b);
  return 0;
}

int emul_post_physics_step(const B200EnvParams* p, const B200EnvBuffers* b, int64_t step) {
  return emul_post_physics_step_variant(p, b, step, 0);
}

// the command-curriculum rule alone (go2.py:87-107): {lo, hi} in force -> {lo, hi} after `count` resets with episode-sum `sum`
int emul_command_curriculum_rule(const B200EnvParams* p, int count, double sum, const double* in_force, double* next) {
  command_curriculum_rule(*p, count, sum, in_force, next);
  return 0;
}

int emul_env_init(const B200EnvParams* p, const B200EnvBuffers* b, const B200InitParams* init) {
  for (int e = 0; e < p->num_envs; ++e) env_init_one(*p, *b, *init, e);
  return 0;
}

int emul_reset_all(const B200EnvParams* p, const B200EnvBuffers* b, int64_t step, int init_done) {
  for (int e = 0; e < p->num_envs; ++e) env_reset_only(*p, *b, e, step, init_done);
  extras(*p, *b);
  return 0;
}

}  // extern "C"
