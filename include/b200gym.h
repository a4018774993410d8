/* libb200gym.so -- C ABI of the B200-native legged_gym_custom hot path.
 *
 * The reference (JustinMLu/legged_gym_custom) has no FFI layer: its boundary is Python
 * duck-typing between OnPolicyRunner, the env object and the PPO object.
 * This header is the boundary a maintainer binds UNDER those Python classes (ctypes stub in
 * INTEGRATION.md).  Each entry point names the reference code it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is DEVICE memory unless it says "host";
 *  - the caller (PyTorch on the Python side) owns all memory; the library allocates nothing
 *    but small per-handle scratch, and never frees caller memory;
 *  - every call is asynchronous on the passed cudaStream_t (as void*), never synchronises,
 *    never reads back to the host;
 *  - return 0 = OK, <0 = argument error, >0 = cudaError_t; text via b200_last_error();
 *  - one host thread per process/GPU; not re-entrant per handle.
 */
#ifndef B200GYM_H_
#define B200GYM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 3
#define B200_NUM_DOF 12          /* go2 family: {FL,FR,RL,RR}_{hip,thigh,calf}_joint */
#define B200_NUM_BODIES 19       /* base, Head_upper, Head_lower, 4 x {hip,thigh,calf,foot} */
#define B200_NUM_FEET 4          /* FL, FR, RL, RR (reference order, go2.py:295-298) */
#define B200_MAX_SCAN_AXIS 24
#define B200_MAX_PROPRIO 64
#define B200_MAX_SCAN 192         /* scan_nx * scan_ny upper bound (base LeggedRobot uses 17 x 11 = 187) */
#define B200_MAX_HIST 640          /* history_len * num_proprio upper bound */
#define B200_PROPRIO 52          /* go2 cur-obs layout (go2.py:506-515); other widths are rejected */

/* Reward terms: every `_reward_*` of legged_robot.py:1036-1148 and go2.py:578-831, in the
 * ALPHABETICAL order in which the reference sums them (class_to_dict iterates dir(),
 * helpers.py:45; legged_robot.py:735-750).  A zero scale disables a term, exactly like
 * _prepare_reward_function pops it.  `termination` is applied after the positive clip
 * (legged_robot.py:234-237) and is kept last. */
enum B200RewardTerm {
  B200_REW_action_rate = 0,
  B200_REW_ang_vel_xy,
  B200_REW_base_height,
  B200_REW_calf_collision,
  B200_REW_calf_pos,
  B200_REW_calf_symmetry,
  B200_REW_collision,
  B200_REW_delta_torques,
  B200_REW_dof_acc,
  B200_REW_dof_error,
  B200_REW_dof_pos_limits,
  B200_REW_dof_vel,
  B200_REW_dof_vel_limits,
  B200_REW_feet_air_time,
  B200_REW_feet_contact_forces,
  B200_REW_heading_alignment,
  B200_REW_hip_pos,
  B200_REW_jump_zone_forward_vel,
  B200_REW_jump_zone_upward_vel,
  B200_REW_lin_vel_z,
  B200_REW_min_height,
  B200_REW_orientation,
  B200_REW_phase_contact_match,
  B200_REW_phase_foot_lifting,
  B200_REW_reverse_penalty,
  B200_REW_stand_still,
  B200_REW_stumble_calves,
  B200_REW_stumble_feet,
  B200_REW_thigh_pos,
  B200_REW_thigh_symmetry,
  B200_REW_torque_limits,
  B200_REW_torques,
  B200_REW_tracking_ang_vel,
  B200_REW_tracking_lin_vel,
  B200_REW_tracking_pitch,
  B200_REW_tracking_roll,
  B200_REW_zero_cmd_dof_error,
  B200_REW_termination,
  B200_NUM_REWARD_TERMS
};

/* Constants the reference bakes at init from its cfg classes (legged_robot.py:933-955 _parse_cfg,
 * :625-727 _init_buffers, :730-754 _prepare_reward_function; go2.py:110-129).  All "python
 * scalars" are stored as the fp32 value torch would broadcast them to. */
typedef struct B200EnvParams {
  int32_t abi_version;
  int32_t num_envs;
  int32_t num_proprio;        /* 52 */
  int32_t history_len;        /* 10 */
  int32_t num_priv;           /* 29 = mass/com 4 + friction 1 + kp 12 + kd 12 */
  int32_t num_est;            /* 3 */
  int32_t num_scan;           /* scan_nx * scan_ny = 132 */
  int32_t control_type;       /* 0 'P', 1 'V', 2 'T' (legged_robot.py:456-471) */
  int32_t randomize_kp_kd;
  int32_t decimation;
  float sim_dt;
  float dt;                   /* decimation * sim_dt */
  float action_scale;
  float clip_actions;
  float clip_obs;
  float p_gains[B200_NUM_DOF];
  float d_gains[B200_NUM_DOF];
  float default_dof_pos[B200_NUM_DOF];
  float torque_limits[B200_NUM_DOF];
  float dof_pos_lo[B200_NUM_DOF];      /* soft limits (legged_robot.py:354-357) */
  float dof_pos_hi[B200_NUM_DOF];
  float dof_vel_limits[B200_NUM_DOF];
  /* episode / domain randomisation */
  int32_t max_episode_length;          /* ceil(episode_length_s / dt); time-out is ep_len > this */
  float max_episode_length_s;
  int32_t resample_interval;           /* int(resampling_time / dt) */
  int32_t push_robots;
  int32_t push_interval;
  float max_push_vel;
  /* gait phase (go2.py:279-283) */
  float period, fr_offset, bl_offset, fl_offset, br_offset;
  /* commands: index 0 vx, 1 vy, 2 yaw rate, 3 heading; value = span * u + lo */
  float cmd_lo[4];
  float cmd_span[4];
  int32_t heading_command;
  float heading_error_gain;
  int32_t zero_command;
  float zero_command_prob;
  /* terrain / height scan (legged_robot.py:997-1032) */
  int32_t has_height_samples;          /* 0 for mesh_type 'plane' -> heights == 0 */
  int32_t hs_rows, hs_cols;
  float border_size, horizontal_scale, vertical_scale;
  int32_t index_div_mode;              /* 0: x / h_scale (torch CPU); 1: x * fp32(1/h_scale) (torch CUDA) */
  int32_t parkour;                     /* hole termination + jump flags (go2.py:202-204, :487-494) */
  int32_t curriculum;
  int32_t custom_origins;              /* heightfield/trimesh -> xy randomised on reset */
  float promote_dist;                  /* env_length * promote_threshold */
  float demote_threshold;
  int32_t max_terrain_level;           /* = terrain.num_rows */
  int32_t terrain_cols;
  int32_t scan_nx, scan_ny;
  float scan_x[B200_MAX_SCAN_AXIS];
  float scan_y[B200_MAX_SCAN_AXIS];
  /* reset */
  float base_init_state[13];
  float dof_reset_lo, dof_reset_span;  /* default + U(0, 0.9) (legged_robot.py:491) */
  /* observations (go2.py:467-574) */
  float obs_lin_vel, obs_ang_vel, obs_dof_pos, obs_dof_vel;
  int32_t add_noise;
  float noise_vec[B200_MAX_PROPRIO];
  /* rewards */
  float reward_scales[B200_NUM_REWARD_TERMS];   /* already multiplied by dt */
  int32_t only_positive_rewards;
  float tracking_sigma, base_height_target, max_contact_force, max_foot_height;
  float stance_threshold;              /* 2 * percent_time_on_ground - 1 */
  float soft_dof_vel_limit, soft_torque_limit, pitch_deg_target, roll_deg_target;
  /* rigid-body / joint index tables */
  int32_t feet[B200_NUM_FEET];
  int32_t calves[B200_NUM_FEET];
  int32_t n_penalised;
  int32_t penalised[B200_NUM_BODIES];
  int32_t n_termination;
  int32_t termination[B200_NUM_BODIES];
  int32_t hip_joints[4], thigh_joints[4], calf_joints[4];
  /* largest fp32 s with sqrt_rn(s) <= 1.0 / 0.1: "norm > t" is evaluated as "squared norm > s" (bit-identical masks,
   * no square root) for the termination (legged_robot.py:146) and collision (:1088) contact tests */
  float contact_thr2_term, contact_thr2_collision;
  /* != 0: the step's observation outputs exist ONCE, as the row of critic_obs_buf -- [obs | priv | est | scan] is exactly
   * what obs_buf, privileged_obs_buf, estimated_obs_buf and scan_obs_buf hold (go2.py:538-563 concatenates the same clipped
   * values), so those four pointers are ignored and the caller exposes them as column slices of the critic rows.  Saves a
   * quarter of the step's writes and lets the rows land directly in a rollout-storage slot (bufs->critic_obs_buf may point
   * there).  The go2 layout then runs post_physics_tile_kernel (rows assembled in shared memory, bulk-copied out). */
  int32_t alias_outputs;
  uint64_t seed;                       /* Philox key (oracle/philox.py, csrc/philox.cuh) */
  /* command curriculum (go2.py:80-107, :222-223; cfg.commands.curriculum, off in every shipped go2 cfg).  The reference
   * keeps the lin_vel_x range as Python floats and moves it with np.clip in double: so do we (B200EnvBuffers.command_ranges) */
  double cc_vel_increment, cc_max_forward_vel, cc_max_reverse_vel;
  double cc_range0[2];                 /* cfg.commands.ranges.lin_vel_x as doubles: initial content of command_ranges */
  float cc_threshold;                  /* fp32(0.8 * reward_scales[tracking_lin_vel]) -- the fp32 tensor comparison of go2.py:91 */
  int32_t command_curriculum;
  /* height_samples row pitch in ELEMENTS (>= hs_cols).  A multiple of 8 (16 bytes) lets the tile kernel fetch an env's scan
   * neighbourhood as ONE 2-D TMA box; 0 means hs_cols. */
  int32_t hs_pitch;
  /* != 0 (needs hs_pitch % 8 == 0, the go2 layout and alias_outputs): the height scan stages the env's terrain tile -- the
   * 32 x 32 int16 cells around its scan points, one TMA box -- in shared memory and reads the 132 x 3 cells from there
   * (north_star design choice 2) instead of gathering them from the L2-resident field. */
  int32_t terrain_tiles;
} B200EnvParams;

/* Device buffers of one env shard.  "PhysX" = written by the simulator each step and only
 * READ here except for reset/push writes (exactly the in-place writes the reference pushes
 * back with set_*_tensor_indexed).  "own" = persistent state of this library's env.
 * Shapes use N = num_envs; all float = fp32; rows are contiguous. */
typedef struct B200EnvBuffers {
  /* PhysX (legged_robot.py:632-646, go2.py:136-138) */
  float* root_states;            /* [N,13] in/out */
  float* dof_state;              /* [N*12,2] in/out */
  const float* contact_forces;   /* [N*19,3] */
  const float* rigid_body_states;/* [N*19,13] */
  /* static per-env randomisation (legged_robot.py:687-701) */
  const float* kp_kd_multipliers;    /* [2,N,12] */
  const float* priv_mass_params;     /* [N,4] */
  const float* priv_friction;        /* [N,1] */
  /* terrain */
  const int16_t* height_samples;     /* [hs_rows,hs_cols] with row pitch params.hs_pitch, or NULL */
  const float* terrain_origins;      /* [max_terrain_level,terrain_cols,3] or NULL */
  /* own persistent state */
  float* actions;                /* [N,12] clipped actions of this step (written by b200_pd_torques) */
  float* torques;                /* [N,12] */
  float* commands;               /* [N,4] */
  int64_t* episode_length_buf;   /* [N] */
  float* last_actions;           /* [N,12] */
  float* last_dof_vel;           /* [N,12] */
  float* last_root_vel;          /* [N,6] */
  float* last_base_lin_vel;      /* [N,3] */
  float* last_torques;           /* [N,12] */
  float* obs_history_buf;        /* [N,history_len,num_proprio] */
  uint8_t* last_contacts;        /* [N,4] bool */
  float* last_contact_heights;   /* [N,4] */
  float* feet_air_time;          /* [N,4] */
  float* jump_flags;             /* [N,1] */
  float* episode_sums;           /* [N,B200_NUM_REWARD_TERMS] env-major; columns of disabled terms stay 0 */
  int64_t* terrain_levels;       /* [N] */
  const int64_t* terrain_types;  /* [N] */
  float* env_origins;            /* [N,3] */
  /* derived quantities kept for the Python attribute surface (play.py:81-92) */
  float* base_lin_vel;           /* [N,3] */
  float* base_ang_vel;           /* [N,3] */
  float* projected_gravity;      /* [N,3] */
  float* rpy;                    /* [N,3] roll, pitch, yaw */
  float* measured_heights;       /* [N,num_scan] */
  int64_t* height_index;         /* [N,num_scan,2] clipped (px,py); NULL = do not record (parity tests only) */
  float* phases;                 /* [N,5] phase, fr, fl, bl, br */
  uint8_t* foot_contacts;        /* [N,4] bool fl, fr, bl, br (filtered) */
  /* step outputs (legged_robot.py:100) */
  float* obs_buf;                /* [N,(history_len+1)*num_proprio] clipped */
  float* privileged_obs_buf;     /* [N,num_priv] clipped */
  float* critic_obs_buf;         /* [N,obs+priv+est+scan] clipped */
  float* estimated_obs_buf;      /* [N,num_est] clipped */
  float* scan_obs_buf;           /* [N,num_scan] (NOT clipped to clip_obs, legged_robot.py:97) */
  float* rew_buf;                /* [N] */
  uint8_t* reset_buf;            /* [N] bool */
  uint8_t* time_out_buf;         /* [N] bool */
  /* extras (go2.py:246-263): refreshed only on steps where >=1 env reset */
  uint8_t* extras_time_outs;     /* [N] bool */
  float* extras_episode;         /* [B200_NUM_REWARD_TERMS + 1]: rew_<term> means, then terrain_level */
  int32_t* reset_count;          /* [1] number of envs reset this step */
  float* reset_episode_sums;     /* [N,B200_NUM_REWARD_TERMS] scratch: pre-zeroing sums of envs reset this step */
  /* command curriculum (only touched when params.command_curriculum != 0) */
  double* command_ranges;        /* [4] lin_vel_x {lo, hi} in force, then {lo, hi} the resets of THIS step resample from */
  float* cc_value;               /* [N] scratch: episode_sums[tracking_lin_vel] + this step's term (dry pass) */
  uint8_t* cc_reset;             /* [N] scratch: the env resets this step (dry pass) */
} B200EnvBuffers;

typedef struct B200Env B200Env;  /* opaque handle: params in device constant storage + scratch */

const char* b200_last_error(void);
int b200_abi_version(void);
int b200_env_params_size(void);      /* sizeof(B200EnvParams), for the binding's self-check */
int b200_env_buffers_size(void);

/* Env-creation-time domain randomisation and buffer initialisation: replaces the host loops of
 * LeggedRobot._process_rigid_shape_props / _process_rigid_body_props (legged_robot.py:306-380: friction buckets, added base
 * mass, centre-of-mass shift), the kp/kd multipliers of _init_buffers (:696-701), _get_env_origins (:897-930: initial
 * terrain levels / types / origins, or the plane grid).  Draws are keyed Philox (sites 8-12, env = index, step = 0). */
typedef struct B200InitParams {
  int32_t randomize_friction;    /* 0: every env gets dynamic_friction */
  float friction_lo, friction_hi, dynamic_friction;
  int32_t randomize_base_mass;
  float mass_lo, mass_hi;
  int32_t randomize_com;
  float com_lo, com_hi;
  float kp_kd_lo, kp_kd_hi;
  int32_t num_init_levels;       /* terrain: levels ~ randint(0, num_init_levels); <= 0: no height field (plane grid) */
  int32_t terrain_cols;
  float env_spacing;             /* plane grid */
  int32_t grid_cols;             /* plane grid: floor(sqrt(N)) */
} B200InitParams;
/* Writes priv_friction, priv_mass_params, kp_kd_multipliers, terrain_levels, terrain_types, env_origins (the const
 * qualifiers of those B200EnvBuffers members describe the STEP kernels; this call is their producer). */
int b200_env_init_randomisation(B200Env* env, const B200EnvBuffers* bufs, const B200InitParams* init, void* stream);

/* Handle lifetime. Copies `params`; `device` is the CUDA ordinal. */
int b200_env_create(const B200EnvParams* params, int device, B200Env** out);
int b200_env_destroy(B200Env* env);
/* The go2 layout (history 10, scan 12 x 11: every registered go2 task) runs a kernel variant with the layout
 * baked in; `on` != 0 forces the layout-generic variant instead (tests run both against the oracle). */
int b200_env_force_generic_layout(B200Env* env, int on);
/* Profiling aid: when `trace` (device, uint64 [ceil(num_envs / 8)][8]) is non-NULL, thread 0 of every CTA of the
 * post-physics kernel records %globaltimer (ns) at its phase boundaries: [0] start, [1] rows loaded + height scan,
 * [2] items, [3] reward terms + history loads, [4] reward sum / reset, [5] end (observations written). */
int b200_env_set_phase_trace(B200Env* env, unsigned long long* trace);
/* Performance switch (results are unaffected): `on` != 0 makes every CTA of the post-physics kernel issue L2 prefetches
 * for its envs' obs_history_buf rows (go2.py:566-576, the largest read of the step) at kernel entry, ahead of the two
 * dependent round trips (state rows, height gathers) that precede their use.  Off by default: at 4096 envs the rows are
 * L2-resident between steps. */
int b200_env_set_prefetch(B200Env* env, int on);

/* Terrain construction on the device (init-time).
 * b200_parkour_field: the parkour height field of Terrain.parkour_curriculum / parkour_selected_terrain (terrain.py:103-131,
 * terrain_utils.py:318-399) -- `field` int16 [rows, cols] = tile_rows x tile_cols tiles of length_px x width_px cells inside a
 * border; tile (i, j) is described by tiles[i * tile_cols + j] (device array): start platform, up to 32 obstacles in the
 * reference's drawing order (rows [row_lo, row_hi) get `height` except columns < zero_below or >= zero_from, which get 0),
 * side walls of `pad` cells at `border_height`.  The host builds the tables with the reference's rounding / slice semantics.
 * b200_heightfield_to_trimesh: convert_heightfield_to_trimesh (terrain_utils.py:401-465) -- vertices float [rows * cols, 3],
 * triangles uint32 [2 * (rows - 1) * (cols - 1), 3], with the slope-threshold correction when use_slope_threshold != 0. */
#define B200_MAX_TILE_OBSTACLES 32
typedef struct B200ParkourTile {
  int32_t platform_rows, num_obstacles, pad;
  int16_t platform_height, border_height;
  int32_t row_lo[B200_MAX_TILE_OBSTACLES], row_hi[B200_MAX_TILE_OBSTACLES];
  int32_t zero_below[B200_MAX_TILE_OBSTACLES], zero_from[B200_MAX_TILE_OBSTACLES];
  int16_t height[B200_MAX_TILE_OBSTACLES];
} B200ParkourTile;
int b200_parkour_field(int16_t* field, int rows, int cols, int border, int length_px, int width_px, int tile_rows, int tile_cols,
                       const B200ParkourTile* tiles /* device */, void* stream);
int b200_heightfield_to_trimesh(const int16_t* height_field, int rows, int cols, double horizontal_scale, double vertical_scale,
                                int use_slope_threshold, double slope_threshold, float* vertices, uint32_t* triangles, void* stream);

/* Replaces LeggedRobot.step's action clip (legged_robot.py:74-75) + _compute_torques
 * (legged_robot.py:440-478).  `actions_in` [N,12] raw policy actions; when `clip_and_store`
 * != 0 they are clipped to +-clip_actions and stored into bufs->actions (first substep of an
 * env step), otherwise bufs->actions is used.  Writes bufs->torques. */
int b200_pd_torques(B200Env* env, const B200EnvBuffers* bufs, const float* actions_in,
                    int clip_and_store, void* stream);

/* Replaces Go2Robot.post_physics_step (go2.py:345-387) and everything it calls:
 * update_feet_states, quaternion_to_euler, _post_physics_step_callback (command resample,
 * heading controller, _get_heights, _push_robots), check_termination, compute_reward (all
 * _reward_* terms), reset_idx (terrain curriculum, _reset_dofs, _reset_root_states,
 * _resample_commands, buffer zeroing, episode extras), compute_observations, the last_*
 * updates, and the observation clip of LeggedRobot.step (legged_robot.py:91-95).
 * `common_step_counter` is the value AFTER this step's increment (go2.py:355). */
int b200_post_physics_step(B200Env* env, const B200EnvBuffers* bufs, int64_t common_step_counter,
                           void* stream);
/* The same step split into its two kernels, for measurement: `parts` bit 0 = post_physics_kernel (everything per env),
 * bit 1 = extras_kernel (episode means / time_outs over the envs that reset).  parts = 3 is b200_post_physics_step. */
int b200_post_physics_step_parts(B200Env* env, const B200EnvBuffers* bufs, int64_t common_step_counter, int parts, void* stream);

/* Same, with `common_step_counter` kept in DEVICE memory: the step's kernels use *step_counter_dev + 1 and the last
 * launch of the call stores the increment, so a captured CUDA graph of the rollout replays with advancing step numbers. */
int b200_post_physics_step_dev(B200Env* env, const B200EnvBuffers* bufs, int64_t* step_counter_dev, void* stream);
/* The same split into its two kernels (`parts` as in b200_post_physics_step_parts): bit 1 (extras_kernel, which also commits
 * the counter) may be launched on another stream, ordered after bit 0 and before the next step's bit 0. */
int b200_post_physics_step_dev_parts(B200Env* env, const B200EnvBuffers* bufs, int64_t* step_counter_dev, int parts, void* stream);
int b200_counter_add(int64_t* counter_dev, int64_t delta, void* stream);

/* Replaces BaseTask.reset's reset_idx(arange(N)) (base_task.py:131-135): resets every env
 * with the curriculum update skipped when `init_done` == 0 (legged_robot.py:551-552). */
int b200_reset_all(B200Env* env, const B200EnvBuffers* bufs, int64_t common_step_counter,
                   int init_done, void* stream);

/* Standalone height scan = LeggedRobot._get_heights (legged_robot.py:997-1032) for all envs;
 * writes bufs->measured_heights and, if non-NULL, bufs->height_index. */
int b200_get_heights(B200Env* env, const B200EnvBuffers* bufs, void* stream);

/* Replaces RolloutStorage.compute_returns (rollout_storage.py:110-124): GAE reverse scan,
 * then advantages = (A - mean(A)) / (std_unbiased(A) + 1e-8) over all T*N.
 * rewards/values/returns/advantages [T,N] fp32, dones [T,N] uint8, last_values [N].
 * `scratch` >= b200_gae_scratch_bytes(T, N) bytes of device memory. */
int64_t b200_gae_scratch_bytes(int T, int N);
int b200_compute_returns(const float* rewards, const uint8_t* dones, const float* values,
                         const float* last_values, float* returns, float* advantages,
                         int T, int N, float gamma, float lam, void* scratch, void* stream);

/* PPO.process_env_step's time-out bootstrap + RolloutStorage.add_transitions scalars
 * (ppo.py:156-171, rollout_storage.py:87-105): rewards_t = rew + gamma * value * time_out;
 * dones_t = reset.  `time_outs` may be NULL ('time_outs' not in infos). */
int b200_store_step_scalars(const float* rew, const uint8_t* reset, const uint8_t* time_outs,
                            const float* values, float gamma, float* rewards_t, uint8_t* dones_t,
                            int N, void* stream);

/* ---- learner: fused Linear layers on tensor cores (csrc/linear_kernels.cu) ----------------------
 * Replace the nn.Linear (+ nn.ELU) stacks of ActorCritic / MlpEstimator / ScanEncoder /
 * PrivilegedEncoder / AdaptationEncoder (actor_critic.py:84-108, support_networks.py:9-175) and
 * their autograd backward.  Row-major fp32; every leading dimension a multiple of 4 floats and
 * X / W / dY pointers 16-byte aligned.  act: 0 none, 1 ELU.  precise: 0 = TF32 (the reference's
 * matmul precision on GPU, train.py:39), 1 = 3xTF32 (fp32-accurate, used for parity). */
int b200_linear_forward(const float* X, int ldx, const float* W, int ldw, const float* bias, float* Y, int ldy,
                        int M, int N, int K, int act, int precise, void* stream);
/* dX[M,K] (+)= (dY[M,N] . W[N,K]) * elu'(Yprev[M,K])   (Yprev may be NULL) */
int b200_linear_dgrad(const float* dY, int lddy, const float* W, int ldw, const float* Yprev, int ldyp, float* dX,
                      int lddx, int M, int N, int K, int accumulate, int precise, void* stream);
/* dW[N,K] += dY[M,N]^T . X[M,K] ;  db[N] += column sums of dY   (db may be NULL) */
int b200_linear_wgrad(const float* dY, int lddy, const float* X, int ldx, float* dW, int ldw, float* db, int M, int N,
                      int K, int precise, void* stream);

/* ---- learner: the same three GEMMs on the tcgen05 / TMEM / TMA path (csrc/mlp_tcgen05.cu, kind::tf32) ----------
 * Same argument meaning as b200_linear_*; bias gradients stay with b200_linear_wgrad / the caller. */
/* Large forward / dgrad problems run on CTA pairs (tcgen05 cta_group::2, 256-row tiles, each CTA stages half of B);
 * `on` = 0 forces the single-CTA kernels everywhere (A/B measurements and tests).  Returns 0. */
int b200_tc_set_pair_mode(int on);
/* Profiling aid: the i-th tcgen05 GEMM launched after this call min/max-es its CTAs' %globaltimer (ns) into
 * dev_buf[i] = {first CTA start, last CTA end} (device, uint64 [capacity][2]; the caller fills it with {~0, 0} before every
 * run) and describes itself in host_meta[i] = {M, N, K, mode * 1000 + tile width (+ 500 for a CTA pair)} (host, int64
 * [capacity][4], written at call time).  Capture-safe: a CUDA-graph node keeps the slot it was captured with.  NULL = off. */
int b200_tc_set_trace(unsigned long long* dev_buf, long long* host_meta, int capacity);
/* Optional: launch the tcgen05 GEMMs with programmatic stream serialisation (PDL): each triggers its dependents at entry
 * and waits for its predecessor (griddepcontrol.wait) only after its prologue, so barrier init / TMEM allocation /
 * descriptor prefetch overlap the previous kernel's tail.  Off by default: measured neutral-to-negative for the update
 * (early-resident CTAs take SMs away from the kernels of the side streams).  Returns 0. */
int b200_tc_set_pdl(int on);
/* GEMM launches issued while a cap is set size their persistent grid (and the wgrad split-K) for `sms` SMs instead of the whole
 * device; 0 = no cap.  The learner caps the low-priority side-stream chains (critic, estimator) so that their one-wave
 * persistent kernels leave SMs to the small kernels of the critical path.  Host-side state, read at launch time. */
int b200_tc_set_sm_cap(int sms);
/* The same cap as a property of ONE stream: every GEMM launched to `stream` from now on is sized for `sms` SMs (0 removes
 * the cap).  What the learner uses (set once per side stream): no process-wide state to toggle around launches. */
int b200_tc_set_stream_sm_cap(void* stream, int sms);
/* 2: forward / dgrad launches run TWO persistent CTAs per SM on tiles <= 128 columns wide (2 x 256 TMEM columns, half the
 * operand ring each), so that one CTA's epilogue overlaps the other's main loop; 1: one CTA per SM, tiles up to 256 wide,
 * CTA pairs for the large forward problems; 0 (default): by shape -- outputs up to 256 columns wide take 2, wider ones 1.
 * Host-side state, read at launch time (A/B switch). */
int b200_tc_set_ctas_per_sm(int n);
/* != 0 (default): the dgrad kernels move their epilogue tiles by TMA (stored activation in, dX out, 128-byte-swizzled 32 x 32
 * boxes) wherever the operands allow it; 0: the staging-tile epilogue everywhere (A/B switch). */
int b200_tc_set_tma_epilogue(int on);
/* != 0: weight gradients with >= 256 output rows run on CTA pairs (cta_group::2, 256-row UMMA: X is staged once per pair);
 * (default); 0: single CTAs (A/B switch).  Measured on B200, M = 24576: 512 x 627 49.7 -> 43.9 us, 512 x 736 45.4 -> 37.9 us. */
int b200_tc_set_wgrad_pairs(int on);
/* A whole Linear / ELU chain (an MLP's forward pass) in ONE persistent launch: Y_0 = act_0(X . W_0^T + b_0), Y_l =
 * act_l(Y_{l-1} . W_l^T + b_l).  Replaces num_layers b200_tc_linear_forward launches (actor_critic.py:84-135,
 * support_networks.py:9-120 at rollout batch sizes, where launch latency, TMEM allocation and pipeline fill dominate): the
 * CTAs stay resident across the layers and meet at a grid-wide barrier in global memory between two layers.  Every Y_l must
 * be a distinct buffer with ldy % 4 == 0 (it is the next layer's TMA operand); K >= 8; any N (TMA zero-fills the weight
 * rows of a narrow head).  `sync`: 4 zero-initialised uint32 on the device, owned by this call site (concurrent chains
 * on different streams need their own); never reset by the host.  `max_ctas` (0 = all): the CTAs of the launch wait for
 * each other, so all of them must be resident -- an SM holds two -- and chains that may run CONCURRENTLY must share the
 * budget of 2 x SMs between them.  Same arithmetic as b200_tc_linear_forward (kind::tf32, fp32 accumulation), for every N. */
typedef struct B200MlpLayer {
  const float* W;      /* [N, ldw] */
  const float* bias;   /* [N] or NULL */
  float* Y;            /* [M, ldy] */
  int32_t ldw, ldy, N, K, act;   /* act: 0 none, 1 ELU */
} B200MlpLayer;
int b200_tc_mlp_forward(const B200MlpLayer* layers /* host */, int num_layers, const float* X, int ldx, int M, unsigned int* sync, int max_ctas,
                        void* stream);
/* dgrad `accumulate`: 0 = overwrite dX; 1 = add to the existing dX; n > 1 = add to the first n columns of dX only (the
 * PPO loss head leaves the ROA regulariser's gradient in the latent columns of the [latent | scan latent] gradient). */
int b200_tc_linear_supported(int M, int N, int K);
int b200_tc_linear_forward(const float* X, int ldx, const float* W, int ldw, const float* bias, float* Y, int ldy,
                           int M, int N, int K, int act, void* stream);
int b200_tc_linear_dgrad(const float* dY, int lddy, const float* W, int ldw, const float* Yprev, int ldyp, float* dX,
                         int lddx, int M, int N, int K, int accumulate, void* stream);
/* Same, and additionally dbias_prev[K] += column sums of dX: dX is d(loss)/d(pre-activation) of the layer below, so its
 * column sum is that layer's bias gradient (replaces a separate b200_colsum pass over dX).  Needs accumulate = 0. */
int b200_tc_linear_dgrad_bias(const float* dY, int lddy, const float* W, int ldw, const float* Yprev, int ldyp, float* dX, int lddx,
                              int M, int N, int K, int accumulate, float* dbias_prev, void* stream);
int b200_tc_linear_wgrad(const float* dY, int lddy, const float* X, int ldx, float* dW, int ldw, int M, int N, int K,
                         void* stream);

/* ---- learner: storage traffic, heads, optimiser (csrc/learner_kernels.cu) ---------------------- */
typedef struct B200CopySeg {
  const float* src;
  float* dst;
  int32_t width, src_ld, dst_ld, _pad;
} B200CopySeg;
/* RolloutStorage.add_transitions' copies (rollout_storage.py:87-105): up to 8 strided 2-D copies, one launch. */
int b200_copy_segments(const B200CopySeg* segs /* host */, int nseg, int rows, void* stream);
/* mini_batch_generator's gathers (rollout_storage.py:134-181): dst[i,:] = src[idx[i],:] */
int b200_gather_rows(const float* src, int src_ld, const int64_t* idx, float* dst, int dst_ld, int width, int64_t rows, void* stream);
int b200_gather_bytes(const uint8_t* src, const int64_t* idx, uint8_t* dst, int64_t rows, void* stream);
/* ActorCritic.act + get_actions_log_prob (actor_critic.py:190-226): a = mu + std*z, z keyed by (seed, step, env, action). */
int b200_sample_actions(const float* mu, int ldmu, const float* std, uint64_t seed, int64_t step, float* actions, float* logp,
                        float* mu_out, float* sigma_out, int N, int A, void* stream);

/* Same with the noise step counter in device memory: `step_counter_dev` points at TWO int64 words -- [0] the counter (read by
 * the launch, then incremented by its last CTA), [1] a ticket word the caller zero-initialises once and the kernel leaves at
 * zero.  One launch; replays of a captured graph advance the counter. */
int b200_sample_actions_dev(const float* mu, int ldmu, const float* std, uint64_t seed, int64_t* step_counter_dev, float* actions,
                            float* logp, float* mu_out, float* sigma_out, int N, int A, void* stream);

/* PPO.update's loss head, forward + backward (ppo.py:199-270). `sums` receives SUMS over the minibatch of
 * surrogate, value, regularisation and entropy terms (divide by M for the reference's means). */
typedef struct B200PpoLossArgs {
  const float* mu;        int32_t ldmu;
  const float* std;
  const float* actions;
  const float* old_logp;
  const float* adv;
  const float* returns;
  const float* target_values;
  const float* value;     int32_t ldv;
  const float* latent_p;  int32_t ldlp;
  const float* latent_a;  int32_t ldla;
  float* dmu;             int32_t lddmu;
  float* dvalue;          int32_t lddv;
  float* dlatent_p;       int32_t lddlp;
  float* dstd;
  float* sums;
  int32_t M, A, L;
  float clip, value_coef, entropy_coef, reg_coef;
  int32_t use_clipped_value_loss;
  const float* reg_coef_dev;   /* if non-NULL, the ROA coefficient is read from device memory (graph replay) */
} B200PpoLossArgs;
int b200_ppo_loss(const B200PpoLossArgs* args /* host */, void* stream);
/* estimator loss mean ||pred - target||_2^2 (ppo.py:224-226) and DAgger loss mean ||target - pred||_2 (ppo.py:330-333) */
int b200_mse_rows_loss(const float* pred, int ldp, const float* target, int ldt, float* dpred, int lddp, float* sum, int M, int D, void* stream);
int b200_l2_rows_loss(const float* pred, int ldp, const float* target, int ldt, float* dpred, int lddp, float* sum, int M, int D, void* stream);
int b200_elu_backward(float* dY, int lddy, const float* Y, int ldy, int M, int N, void* stream);
/* schedule == 'adaptive' (ppo.py:233-246), kept on the device so that minibatches stay graph-replayable.
 * b200_kl_sum adds sum_i KL(old_i || new_i) of the M samples to acc[0] (`acc` = 2 doubles on the device, zero before the first
 * call; with several ranks all-reduce acc[0] between the two calls).  b200_adaptive_lr forms kl_mean = acc[0] / count, applies
 * lr /= 1.5 (floor 1e-5) if kl_mean > 2 desired_kl, lr *= 1.5 (cap 1e-2) if 0 < kl_mean < desired_kl / 2, to the main
 * optimiser's `adam_state[4]` (see b200_clip_adam), stores kl_mean in acc[1] and clears acc[0]. */
int b200_kl_sum(const float* mu, int ldmu, const float* std, const float* old_mu, int ldom, const float* old_sigma, int ldos, int M, int A,
                double* acc, void* stream);
int b200_adaptive_lr(double* acc, int64_t count, double desired_kl, double* adam_state, void* stream);
/* clip_grad_norm_ + Adam.step on flat buffers (ppo.py:228-231, :273-276, :336-339); zeroes `grads`.
 * grad_scale = 1/world_size after the NCCL sum all-reduce of `grads`.
 * `state` = 8 doubles on the device: [0] scratch, [1] step, [2] beta1^step, [3] beta2^step, [4] lr -- advanced
 * by the call itself so a captured CUDA graph replays correctly (initialise to {0, 0, 1, 1, lr, 0, 0, 0}).
 * [7] (in) squared norm of gradients OUTSIDE `grads` that the reference's clip_grad_norm_ call also covers (ppo.py:274
 * clips actor_critic.parameters(): the adaptation encoder's stale post-clip .grad of the last update_dagger is part of
 * the norm and is rescaled with it); the call consumes it as [5] and leaves the rescaled value in [7] for the next step.
 * [6] (out) post-clip squared norm of `grads` -- what update_dagger hands to the main optimiser's [7]. */
int b200_clip_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double* state,
                   float grad_scale, float max_norm, float beta1, float beta2, float eps, void* stream);

/* The data-parallel form of the same step (section 8 row e): gradient exchange + clip_grad_norm_ + Adam as
 * ONE kernel per optimiser step over NVLink / NVSwitch peer memory -- replaces all_reduce(grads) + b200_clip_adam on every
 * rank.  `grads`, `params` and the sync area are SYMMETRIC allocations (one per rank, every rank's mapped into every rank:
 * `*_peer[p]` = rank p's buffer as seen from here, own rank included; host arrays of `world` device pointers); `*_mc` are
 * their NVSwitch multicast addresses or NULL (then the kernel loops over the peers).  Rank r owns shard r of the flat
 * buffers: it reduces that shard of all ranks' gradients (multimem.ld_reduce: the switch adds), contributes the shard's sum
 * of squares to the global norm (W doubles exchanged through the sync area, summed in rank order), applies clip + Adam to
 * its shard only -- `exp_avg` / `exp_avg_sq` are maintained for the own shard only -- and broadcasts the updated parameters
 * (multimem.st); the shard of every rank's `grads` is left zeroed.  Start / end barriers between the ranks are inside the
 * kernel (flags in the sync area: 4 * world uint64, zero-initialised once).  `local`: 8 uint32 of zero-initialised device
 * scratch (barrier epoch, grid counters, partial sum); `gsum`: ceil(n / 4 / world) * 4 floats of scratch.  `state` as for
 * b200_clip_adam (the gradient is the MEAN over ranks, i.e. grad_scale = 1 / world).  Every rank must launch the call the
 * same number of times, in the same order per sync area. */
typedef struct B200DistAdam {
  float* grads;
  float* params;
  float* const* grads_peer;                 /* host [world] */
  float* const* params_peer;                /* host [world] */
  unsigned long long* const* sync_peer;     /* host [world] */
  float* grads_mc;
  float* params_mc;
  float* exp_avg;
  float* exp_avg_sq;
  float* gsum;
  double* state;
  unsigned int* local;
  int64_t n;
  int32_t world, rank;
  float max_norm, beta1, beta2, eps;
} B200DistAdam;
int b200_dist_adam(const B200DistAdam* args /* host */, void* stream);
/* AdaptationEncoder.forward (support_networks.py:128-175) fused into one fp32 kernel.  X [M, >=520] = observation rows
 * whose first 10*52 columns are the proprio history.  Weights in the kernel layouts of networks.py: W1 [30][52],
 * W2 [20][4*32] (tap*32 + channel), W3 [10][2*20], W4 [20][3*12] (step*12 + channel).  proj/c1/c2: optional
 * [M,320] / [M,80] / [M,36] copies of the hidden activations for the backward pass. */
int b200_adaptation_forward(const float* X, int ldx, const float* W1, const float* b1, const float* W2, const float* b2,
                            const float* W3, const float* b3, const float* W4, const float* b4, float* out, int ldo,
                            float* proj, float* c1, float* c2, int M, void* stream);
/* db[N] += column sums of dY[M,N] (bias gradient, companion of b200_tc_linear_wgrad) */
int b200_colsum(const float* dY, int lddy, float* db, int M, int N, void* stream);
int b200_fill(float* p, float value, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200GYM_H_ */
