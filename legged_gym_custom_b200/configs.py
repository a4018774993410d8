"""Task constants for the go2 family, as plain class namespaces with the reference's
attribute names, so `Go2Env` accepts either these or the reference's own cfg objects
(legged_gym/envs/go2/go2_config.py, go2_parkour_config.py, go2_parkour_finetune_config.py
on top of envs/base/legged_robot_config.py).  Only what the hot path reads is kept; values
are the ones the reference resolves at runtime.
"""
import numpy as np


class NS:
    """class-as-namespace; `cfg.section.key` works on the class itself."""


def _dofs(hip_l, hip_r, thigh_f, thigh_r, calf):
    return {'FL_hip_joint': hip_l, 'FL_thigh_joint': thigh_f, 'FL_calf_joint': calf,
            'FR_hip_joint': hip_r, 'FR_thigh_joint': thigh_f, 'FR_calf_joint': calf,
            'RL_hip_joint': hip_l, 'RL_thigh_joint': thigh_r, 'RL_calf_joint': calf,
            'RR_hip_joint': hip_r, 'RR_thigh_joint': thigh_r, 'RR_calf_joint': calf}


SCAN_X = [-0.45, -0.3, -0.15, 0.0, 0.15, 0.3, 0.45, 0.6, 0.75, 0.9, 1.05, 1.2]
SCAN_Y = [-0.75, -0.6, -0.45, -0.3, -0.15, 0.0, 0.15, 0.3, 0.45, 0.6, 0.75]


class Go2Cfg(NS):
    """`go2` flat task (mesh_type plane, yaw-rate commands, 15 reward terms)."""
    seed = 1

    class env(NS):
        num_envs = 4096
        num_proprio = 52
        num_scan_obs = 132
        num_estimated_obs = 3
        num_privileged_obs = 29
        history_buffer_length = 10
        num_actions = 12
        num_observations = 52 * 11
        num_critic_obs = 52 * 11 + 29 + 3 + 132
        env_spacing = 3.
        send_timeouts = True
        episode_length_s = 20
        period, fr_offset, bl_offset, fl_offset, br_offset = 0.45, 0.0, 0.0, 0.5, 0.5

    class terrain(NS):
        mesh_type = 'plane'
        horizontal_scale, vertical_scale, border_size = 0.1, 0.005, 25
        curriculum, parkour, selected = False, False, False
        measure_heights = False
        measured_points_x, measured_points_y = SCAN_X, SCAN_Y
        terrain_length, terrain_width, num_rows, num_cols = 8., 8., 10, 20
        promote_threshold, demote_threshold, max_init_terrain_level = 0.5, 0.4, 1
        static_friction = dynamic_friction = 1.0

    class commands(NS):
        curriculum = False
        max_forward_vel, max_reverse_vel, vel_increment = 1.0, -1.0, 0.10      # command curriculum (go2_config.py:183-187)
        num_commands = 4
        resampling_time = 10.
        heading_command = False
        heading_error_gain = 0.5
        zero_command, zero_command_prob = True, 0.10
        user_command = []

        class ranges(NS):
            lin_vel_x, lin_vel_y, ang_vel_yaw, heading = [-1.0, 1.0], [-0.75, 0.75], [-1.0, 1.0], [-0.2, 0.2]

    class init_state(NS):
        pos, rot, lin_vel, ang_vel = [0.0, 0.0, 0.42], [0.0, 0.0, 0.0, 1.0], [0.0] * 3, [0.0] * 3
        default_joint_angles = _dofs(0.1, -0.1, 0.8, 1.0, -1.5)

    class control(NS):
        control_type = 'P'
        stiffness, damping = {'joint': 40.}, {'joint': 1.}
        action_scale, decimation = 0.25, 4

    class asset(NS):
        foot_name = "foot"
        penalize_contacts_on = ["base", "hip", "thigh", "calf", "Head"]
        terminate_after_contacts_on = ["base"]

    class domain_rand(NS):
        randomize_friction, friction_range = True, [0.3, 1.2]
        randomize_base_mass, added_mass_range = True, [0.0, 3.0]
        randomize_center_of_mass, added_com_range = True, [-0.15, 0.15]
        randomize_kp_kd, kp_kd_range = True, [0.8, 1.2]
        push_robots, push_interval_s, max_push_vel_xy = True, 8, 0.5

    class normalization(NS):
        clip_observations, clip_actions = 100., 3.14

        class obs_scales(NS):
            lin_vel, ang_vel, dof_pos, dof_vel, height_measurements = 2.0, 0.25, 1.0, 0.05, 5.0

    class noise(NS):
        add_noise, noise_level = True, 1.0

        class noise_scales(NS):
            lin_vel, dof_pos, dof_vel, ang_vel, gravity, imu, height_measurements = 0.1, 0.01, 0.05, 0.05, 0.02, 0.02, 0.02

    class rewards(NS):
        only_positive_rewards = True
        tracking_sigma = 0.25
        soft_dof_pos_limit, soft_dof_vel_limit, soft_torque_limit = 0.9, 1., 1.
        base_height_target = 0.25
        pitch_deg_target = roll_deg_target = 0.0
        max_foot_height, percent_time_on_ground, max_contact_force = 0.08, 0.50, 100

        class scales(NS):
            tracking_lin_vel, tracking_ang_vel = 1.5, 1.0
            phase_contact_match, phase_foot_lifting = 1.0, 0.25
            lin_vel_z, action_rate, ang_vel_xy = -2.0, -0.1, -0.01
            torques, dof_acc, delta_torques = -0.00001, -2.5e-7, -1.0e-7
            collision, orientation = -10.0, -5.0
            base_height = -20.0
            dof_error, hip_pos = -0.04, -0.75

    class sim(NS):
        dt = 0.005


class Go2ParkourCfg(Go2Cfg):
    """`go2_parkour`: trimesh gap terrain with curriculum, heading commands, 23 reward terms."""

    class env(Go2Cfg.env):
        period, fr_offset, bl_offset, fl_offset, br_offset = 0.40, 0.0, 0.5, 0.0, 0.5

    class terrain(Go2Cfg.terrain):
        mesh_type = 'trimesh'
        measure_heights = True
        num_rows, num_cols, terrain_length, terrain_width = 12, 20, 28., 10.
        parkour, curriculum, selected = True, True, False
        promote_threshold, demote_threshold, max_init_terrain_level = 0.60, 0.40, 2
        terrain_proportions = [1.0, 0.0]
        gap_x_start, gap_dx, gap_n = 5.0, 3.5, 7
        parkour_kwargs = dict(start_platform_length=3., start_platform_height=0.,
                              x_positions=list(np.arange(5.0, 5.0 + 7 * 3.5, 3.5)), y_positions=[0.0] * 7,
                              obstacle_heights=[-2.0] * 7, obstacle_lengths=[0.2, 0.4, 0.6, 0.8, 1.0, 1.1, 1.2],
                              half_valid_width=5.0, border_width=0.50, border_height=-2.0)

    class commands(Go2Cfg.commands):
        heading_command = True
        max_forward_vel, max_reverse_vel, vel_increment = 1.75, 0.5, 0.10      # go2_parkour_config.py:127-131

        class ranges(NS):
            lin_vel_x, lin_vel_y, ang_vel_yaw, heading = [0.75, 1.5], [0.0, 0.0], [-0.0, 0.0], [-0.2, 0.2]

    class init_state(Go2Cfg.init_state):
        pos = [2.0, 0.0, 0.50]

    class asset(Go2Cfg.asset):
        terminate_after_contacts_on = ["base", "Head"]

    class domain_rand(Go2Cfg.domain_rand):
        friction_range = [0.1, 1.0]

    class rewards(Go2Cfg.rewards):
        base_height_target, max_contact_force = 0.27, 75.0

        class scales(NS):
            tracking_lin_vel, tracking_ang_vel = 2.25, 2.25
            phase_contact_match, phase_foot_lifting = 1.0, 1.0
            action_rate, lin_vel_z, ang_vel_xy = -0.1, -1.0, -0.01
            torques, dof_acc, delta_torques = -0.00001, -2.5e-7, -1.0e-7
            collision, orientation, stumble_feet = -10.0, -1.0, -1.0
            dof_error, hip_pos, thigh_pos = -0.04, -0.5, -0.5
            thigh_symmetry, calf_symmetry = -0.2, -0.2
            heading_alignment, reverse_penalty = -4.5, -1.0
            jump_zone_forward_vel, jump_zone_upward_vel = 1.75, 3.75
            zero_cmd_dof_error = -1.0


class Go2FinetuneCfg(Go2ParkourCfg):
    """`go2_parkour_finetune`: fixed 18-obstacle course, curriculum off, wider vx, +feet_contact_forces."""

    class terrain(Go2ParkourCfg.terrain):
        curriculum = False
        parkour_kwargs = dict(start_platform_length=3., start_platform_height=0.,
                              x_positions=[x0 + d for x0 in (6.0, 10.0, 14.0, 18.0, 22.0, 26.0) for d in (0.0, 0.3, 0.7)],
                              y_positions=[0.0] * 18,
                              obstacle_heights=[h for up in (0.10, 0.15, 0.20, 0.25, 0.30, 0.35) for h in (-2.0, up, -2.0)],
                              obstacle_lengths=[0.3, 0.2, 0.4] * 6,
                              half_valid_width=5.0, border_width=0.50, border_height=-2.0)

    class commands(Go2ParkourCfg.commands):
        class ranges(Go2ParkourCfg.commands.ranges):
            lin_vel_x = [0.5, 2.0]

    class rewards(Go2ParkourCfg.rewards):
        class scales(Go2ParkourCfg.rewards.scales):
            feet_contact_forces = -0.01


class Go2ParkourCfgPPO(NS):
    """PPO / network hyper-parameters (go2_parkour_config.py Go2ParkourCfgPPO)."""
    seed = 1

    class policy(NS):
        actor_hidden_dims = [512, 256, 128]
        critic_hidden_dims = [512, 256, 128]
        init_noise_std = 1.0
        priv_encoder_hidden_dims = [64, 20]
        latent_encoder_output_dim = 20
        scan_encoder_hidden_dims = [128, 64]
        scan_encoder_output_dim = 32
        estimator_hidden_dims = [256, 128]
        use_history = True
        activation = 'elu'

    class algorithm(NS):
        value_loss_coef, use_clipped_value_loss, clip_param, entropy_coef = 1.0, True, 0.2, 0.01
        num_learning_epochs, num_mini_batches = 5, 4
        estimator_learning_rate, learning_rate, schedule = 1e-4, 2e-4, 'fixed'
        gamma, lam, desired_kl, max_grad_norm = 0.99, 0.95, 0.01, 1.
        dagger_update_freq = 20

    class runner(NS):
        policy_class_name, algorithm_class_name = 'ActorCritic', 'PPO'
        num_steps_per_env, max_iterations, save_interval = 24, 5000, 50
        run_name, experiment_name = 'parkour', 'go2_parkour'
        resume, load_run, checkpoint, resume_path = False, -1, -1, None


class Go2FinetuneCfgPPO(Go2ParkourCfgPPO):
    class runner(Go2ParkourCfgPPO.runner):
        run_name = 'parkour_finetune'
        resume = True


TASKS = {
    "go2": (Go2Cfg, Go2ParkourCfgPPO),   # the reference's Go2CfgPPO lacks required keys
    "go2_parkour": (Go2ParkourCfg, Go2ParkourCfgPPO),
    "go2_parkour_finetune": (Go2FinetuneCfg, Go2FinetuneCfgPPO),
}
