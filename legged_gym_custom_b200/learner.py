"""`RolloutStorage` and `PPO`: drop-ins for rsl_rl/storage/rollout_storage.py and
rsl_rl/algorithms/ppo.py with the tensor math done by libb200gym.so.

Same constructor arguments, methods and public tensor attributes as the reference:
`PPO.init_storage / act / process_env_step / compute_returns / update / update_dagger`,
`RolloutStorage.add_transitions / clear / compute_returns / mini_batch_generator` and
`.observations / .values / .returns / .advantages / ...`.

What changes underneath:
  * storage rows are written by one fused multi-segment copy per env step; 29- and 3-wide tensors
    are kept 32- / 4-wide internally (16-byte rows) and exposed as sliced views;
  * compute_returns is the GAE scan + normalisation kernels;
  * the reference draws ONE permutation per update and reuses it for all epochs
    (rollout_storage.py:142, :159-164), so the storage is gathered ONCE per update into permuted,
    contiguous minibatch slabs -- every epoch then reads dense slices, and the actor-input slab
    [T*N, 628] doubles as the concat buffer ([obs | latent | scan latent | estimated obs]);
  * forward / backward of all networks, the loss head, clip + Adam are kernel launches with no host
    synchronisation inside an update; the four logged means come back in one copy at the end.
There is no CPU fallback.
"""
import ctypes as C

import torch

from . import _lib
from .networks import ActorCritic, MlpEstimator, Workspace, _p, ceil4, chain_backward


class RolloutStorage:
    class Transition:
        def __init__(self):
            self.clear()

        def clear(self):
            self.observations = self.privileged_observations = self.critic_observations = None
            self.true_estimated_observations = self.scan_observations = None
            self.actions = self.rewards = self.dones = self.values = None
            self.actions_log_prob = self.action_mean = self.action_sigma = None

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, critic_obs_shape,
                 estimated_obs_shape, scan_obs_shape, actions_shape, device="cuda:0", alias_critic_rows=False):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("RolloutStorage lives on the GPU: the hot path has no CPU fallback")
        self.lib = _lib.lib()
        T, N = num_transitions_per_env, num_envs
        self.num_transitions_per_env, self.num_envs = T, N
        self.obs_shape, self.privileged_obs_shape, self.critic_obs_shape = obs_shape, privileged_obs_shape, critic_obs_shape
        self.estimated_obs_shape, self.actions_shape = estimated_obs_shape, actions_shape
        self.d_obs, self.d_priv, self.d_crit = obs_shape[0], privileged_obs_shape[0], critic_obs_shape[0]
        self.d_est, self.d_scan, self.d_act = estimated_obs_shape[0], scan_obs_shape[0], actions_shape[0]
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=self.device)
        # alias_critic_rows: a critic row IS [obs | priv | est | scan] (go2.py:538-563), so the four observation tensors are
        # column slices of `critic_observations` instead of four more tensors -- and an env that writes its rows straight
        # into slot t + 1 (Go2Env.bind_output_rows) leaves nothing to copy.  One extra slot holds the observation AFTER the
        # last transition (the bootstrap value's input and the next rollout's first row).
        self.alias_critic_rows = bool(alias_critic_rows)
        if self.alias_critic_rows:
            if self.d_crit != self.d_obs + self.d_priv + self.d_est + self.d_scan:
                raise ValueError("alias_critic_rows: critic width must be obs + priv + est + scan")
            self.rows = z(T + 1, N, self.d_crit)
            c0, c1, c2 = self.d_obs, self.d_obs + self.d_priv, self.d_obs + self.d_priv + self.d_est
            self.critic_observations = self.rows[:T]
            self.observations, self._priv = self.rows[:T, :, :c0], self.rows[:T, :, c0:c1]
            self._est, self.scan_observations = self.rows[:T, :, c1:c2], self.rows[:T, :, c2:]
        else:
            self.rows = None
            self.observations = z(T, N, self.d_obs)
            self._priv = z(T, N, ceil4(self.d_priv))           # 29 / 3 wide tensors are kept 32 / 4 wide (16-byte rows)
            self.critic_observations = z(T, N, self.d_crit)
            self._est = z(T, N, ceil4(self.d_est))
            self.scan_observations = z(T, N, self.d_scan)
        # row strides (floats) and copy widths of the five observation tensors
        self.ld_obs, self.ld_priv, self.ld_crit = self.observations.stride(1), self._priv.stride(1), self.critic_observations.stride(1)
        self.ld_est, self.ld_scan = self._est.stride(1), self.scan_observations.stride(1)
        self.w_priv, self.w_est = self._priv.shape[2], self._est.shape[2]
        self.rewards, self.values, self.returns, self.advantages, self.actions_log_prob = (z(T, N, 1) for _ in range(5))
        self.actions, self.mu, self.sigma = z(T, N, self.d_act), z(T, N, self.d_act), z(T, N, self.d_act)
        self.dones = z(T, N, 1, dt=torch.uint8)
        self.saved_hidden_states_a = self.saved_hidden_states_c = None
        self.step = 0
        self._gae_scratch = torch.zeros(max(16, int(self.lib.b200_gae_scratch_bytes(T, N))), dtype=torch.uint8, device=self.device)

    @property
    def privileged_observations(self):
        return self._priv[:, :, :self.d_priv]

    @property
    def true_estimated_observations(self):
        return self._est[:, :, :self.d_est]

    def add_transitions(self, transition):
        """rollout_storage.py:87-105 for callers that hold a Transition (PPO writes rows directly)."""
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        t = self.step
        self.critic_observations[t].copy_(transition.critic_observations)
        self.observations[t].copy_(transition.observations)
        self._priv[t, :, :self.d_priv].copy_(transition.privileged_observations)
        self._est[t, :, :self.d_est].copy_(transition.true_estimated_observations)
        self.scan_observations[t].copy_(transition.scan_observations)
        self.actions[t].copy_(transition.actions)
        self.rewards[t].copy_(transition.rewards.view(-1, 1))
        self.dones[t].copy_(transition.dones.view(-1, 1))
        self.values[t].copy_(transition.values)
        self.actions_log_prob[t].copy_(transition.actions_log_prob.view(-1, 1))
        self.mu[t].copy_(transition.action_mean)
        self.sigma[t].copy_(transition.action_sigma)
        self.step += 1

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam):
        """rollout_storage.py:110-124 (GAE reverse scan + advantage normalisation)."""
        lv = last_values.contiguous()
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(self.lib.b200_compute_returns(p(self.rewards), p(self.dones), p(self.values), p(lv), p(self.returns), p(self.advantages),
                                                 self.num_transitions_per_env, self.num_envs, gamma, lam, p(self._gae_scratch),
                                                 _lib.stream_ptr()))

    def get_statistics(self):
        done = self.dones.clone()
        done[-1] = 1
        flat = done.permute(1, 0, 2).reshape(-1, 1)
        idx = torch.cat((flat.new_tensor([-1], dtype=torch.int64), flat.nonzero(as_tuple=False)[:, 0]))
        return (idx[1:] - idx[:-1]).float().mean(), self.rewards.mean()

    def mini_batch_generator(self, num_mini_batches, num_epochs=8, indices=None):
        """rollout_storage.py:134-181 (same tuple).  PPO.update does not go through this generator: it uses the
        permuted slabs; this is kept for API compatibility and for the tests."""
        batch = self.num_envs * self.num_transitions_per_env
        mb = batch // num_mini_batches
        if indices is None:
            indices = torch.randperm(num_mini_batches * mb, device=self.device)
        f = lambda t: t.flatten(0, 1)
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                idx = indices[i * mb:(i + 1) * mb]
                yield (f(self.observations)[idx], f(self.privileged_observations)[idx], f(self.critic_observations)[idx],
                       f(self.true_estimated_observations)[idx], f(self.scan_observations)[idx], f(self.actions)[idx],
                       f(self.values)[idx], f(self.advantages)[idx], f(self.returns)[idx], f(self.actions_log_prob)[idx],
                       f(self.mu)[idx], f(self.sigma)[idx], (None, None), None)


class _AdamFacade:
    """Minimal torch.optim-like surface (`state_dict`, `load_state_dict`, `param_groups`) over a FlatGroup."""

    def __init__(self, group, lr):
        self.group = group
        self.param_groups = [{"lr": lr}]

    def state_dict(self):
        g = self.group
        if getattr(g, "dist", None) is not None:          # sharded moments: collect every rank's shard over the peer mappings
            m, v = g.dist.full_moments()
        else:
            m, v = g.exp_avg.clone(), g.exp_avg_sq.clone()
        return {"flat_exp_avg": m, "flat_exp_avg_sq": v, "adam_state": g.state.clone(), "param_groups": self.param_groups}

    def load_state_dict(self, sd):
        g = self.group
        if "flat_exp_avg" not in sd:
            raise ValueError("optimizer state of a reference checkpoint is per-tensor; load the model weights and restart the "
                             "optimiser (the reference itself drops the estimator / adaptation optimisers on resume)")
        g.exp_avg.copy_(sd["flat_exp_avg"])
        g.exp_avg_sq.copy_(sd["flat_exp_avg_sq"])
        g.state.copy_(sd["adam_state"])


class PPO:
    def __init__(self, actor_critic, estimator, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998,
                 lam=0.95, value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, estimator_learning_rate=1e-3,
                 max_grad_norm=1.0, use_clipped_value_loss=True, schedule="fixed", desired_kl=0.01, resume=False,
                 device="cuda:0", seed=0, process_group=None):
        if schedule not in ("fixed", "adaptive"):
            raise ValueError(f"unknown schedule '{schedule}'")
        self.device = torch.device(device)
        self.lib = _lib.lib()
        self.desired_kl, self.schedule = desired_kl, schedule
        # ppo.py:233: the KL rule runs only for schedule == 'adaptive' with a desired_kl; it lives on the device (kl_sum /
        # adaptive_lr kernels write the main optimiser's lr slot), so minibatches stay graph-replayable
        self.adaptive = schedule == "adaptive" and desired_kl is not None
        self._learning_rate, self.estimator_learning_rate = learning_rate, estimator_learning_rate
        # ROA schedule (ppo.py:40-43)
        self.start_val, self.end_val, self.start_step, self.duration = 0.0, 0.05, 5000, 10000
        if resume:
            self.start_val, self.end_val, self.start_step, self.duration = 0.0, 0.1, 0, 1
        self.actor_critic, self.estimator = actor_critic, estimator
        self.storage = None
        ac = actor_critic
        ac.main.set_lr(learning_rate)
        ac.adapt.set_lr(learning_rate)
        estimator.group.set_lr(estimator_learning_rate)
        self.optimizer = _AdamFacade(ac.main, learning_rate)
        self.adaptation_optimizer = _AdamFacade(ac.adapt, learning_rate)
        self.estimator_optimizer = _AdamFacade(estimator.group, estimator_learning_rate)
        self.transition = RolloutStorage.Transition()
        self.clip_param, self.num_learning_epochs, self.num_mini_batches = clip_param, num_learning_epochs, num_mini_batches
        self.value_loss_coef, self.entropy_coef, self.gamma, self.lam = value_loss_coef, entropy_coef, gamma, lam
        self.max_grad_norm, self.use_clipped_value_loss = max_grad_norm, use_clipped_value_loss
        self.total_updates = 0.0
        self.seed = seed
        self.act_counter = 0
        self.process_group = process_group            # torch.distributed group for the gradient all-reduce, or None
        self.world_size = 1 if process_group is None else torch.distributed.get_world_size(process_group)
        # data-parallel optimiser steps: ONE kernel does gradient exchange + clip + Adam + parameter all-gather over peer memory
        # (csrc/dist_adam.cu); NCCL all-reduce + the local kernels only where symmetric memory cannot be set up
        from .dist import enable_fused_dist_adam
        self.dist_mode = enable_fused_dist_adam([ac.main, ac.adapt, estimator.group], process_group, max_grad_norm)
        self._perm_gen = torch.Generator(device=self.device).manual_seed(seed + 12345)
        self.act_counter_dev = torch.zeros(2, dtype=torch.int64, device=self.device)     # [counter, the kernel's ticket word]
        self.use_device_counter = False
        self.use_graphs = False
        self.use_streams = True
        # SMs the low-priority side chains of a minibatch (estimator, critic) may occupy; 0 = all.  Their big GEMMs are
        # one-wave persistent kernels that hold every SM until they end, which starves the small kernels of the critical
        # path whatever the stream priorities are (b200_tc_set_stream_sm_cap, registered per side stream).  Measured on B200 (148 SMs), update of 20 minibatches:
        # no cap 14.13-14.15 ms; 132: 13.94; 124: 13.89; 116: 13.76; 108: 13.89; 100: 13.89 (profiles/r3_side_sm_cap_ab.txt).
        # Re-measured with round 2's kernels (whole iteration, ms): no cap 16.08; 132: 15.89; 116: 15.76-15.77; 104: 15.72;
        # 96: 15.67-15.74; 80: 15.83
        self.side_sm_cap = 96
        # weight-gradient GEMMs of the actor / encoder chains on a fifth (low-priority, capped) stream: the dgrad chain that
        # the encoders' backward waits for no longer queues behind them
        self.offload_wgrads = False
        self.defer_critic_join = False      # act(): leave the critic chain running until process_env_step (OnPolicyRunner sets it)
        self._pending_critic = None
        # `defer_store` (OnPolicyRunner sets it): process_env_step() runs on the critic's stream, so neither the critic chain
        # nor the bookkeeping sits on the rollout's critical path; act() makes the main stream wait for it only AFTER the next
        # actor chain is queued (before the caller's env.step can overwrite the env's reward / flag buffers)
        self.defer_store = False
        self._store_done = None
        self.critic_fork_after = 1      # actor layer index the critic chain is forked behind when the store is deferred (-1: beside the actor)
        self._graphs, self._graph_calls = {}, {}

    @property
    def learning_rate(self):
        """the main optimiser's learning rate; under schedule='adaptive' it is read back from the device (one sync)"""
        if self.adaptive:
            return float(self.actor_critic.main.state[4].item())
        return self._learning_rate

    @learning_rate.setter
    def learning_rate(self, value):
        """the lr the kernels use lives in the optimiser's device state: assigning `alg.learning_rate` (as code written
        against the reference does, ppo.py:244-246) moves it there; graph replays pick it up"""
        self._learning_rate = float(value)
        self.optimizer.param_groups[0]["lr"] = float(value)
        self.actor_critic.main.set_lr(float(value))

    # ---- storage ------------------------------------------------------------------------------------
    def init_storage(self, num_envs, num_transitions_per_env, total_obs_shape, privileged_obs_shape, critic_obs_shape,
                     estimated_obs_shape, scan_obs_shape, action_shape, alias_critic_rows=False):
        self.storage = s = RolloutStorage(num_envs, num_transitions_per_env, total_obs_shape, privileged_obs_shape, critic_obs_shape,
                                          estimated_obs_shape, scan_obs_shape, action_shape, self.device, alias_critic_rows)
        ac = self.actor_critic
        T, N = num_transitions_per_env, num_envs
        self.batch = T * N
        self.mb = self.batch // self.num_mini_batches
        z = lambda *sh, dt=torch.float32: torch.zeros(*sh, dtype=dt, device=self.device)
        # permuted slabs (one gather per update); +64 floats of slack after the actor-input slab for the conv windows
        self.p_actor_in = z(self.batch + 1, ac.ld_actor_in)
        self.p_priv, self.p_crit, self.p_scan = z(self.batch, ceil4(s.d_priv)), z(self.batch, s.d_crit), z(self.batch, s.d_scan)
        self.p_est, self.p_act = z(self.batch, ceil4(s.d_est)), z(self.batch, s.d_act)
        self.p_val, self.p_ret, self.p_logp, self.p_adv = z(self.batch), z(self.batch), z(self.batch), z(self.batch)
        # adaptation-encoder latent of every (permuted) sample: the encoder's weights do not move during a PPO update
        # (ppo.py:213-214 runs it under inference_mode; only update_dagger trains it) and every epoch revisits the same
        # samples in the same minibatch slots, so it is evaluated ONCE per update instead of once per minibatch pass
        self.p_lat_a = z(self.batch, ac.latent_dim)
        self.roll_ws = Workspace(self.device)          # rollout-time activations (M = N)
        self.upd_ws = Workspace(self.device)           # update-time activations  (M = minibatch)
        self.loss_sums = z(8)
        if self.adaptive:                              # old mu / sigma slabs + [KL sum, last kl_mean]
            self.p_mu, self.p_sigma = z(self.batch, s.d_act), z(self.batch, s.d_act)
            self.kl_acc = z(2, dt=torch.float64)
        self.reg_coef_dev = z(1)
        self.last_values = z(N, 1)

    def test_mode(self):
        self.actor_critic.test()

    def train_mode(self):
        self.actor_critic.train()

    # ---- rollout --------------------------------------------------------------------------------------
    def _copy_segments(self, segs, rows):
        arr = (_lib.CopySeg * len(segs))()
        for i, (src, sld, dst, dld, w) in enumerate(segs):
            arr[i].src, arr[i].dst, arr[i].width, arr[i].src_ld, arr[i].dst_ld = src, dst, w, sld, dld
        _lib.check(self.lib.b200_copy_segments(arr, len(segs), rows, _lib.stream_ptr()))

    def act(self, obs, privileged_obs, critic_obs, true_estimated_obs, scan_obs, adaptation_mode=False):
        """ppo.py:129-153 fused with RolloutStorage.add_transitions' observation copies."""
        s, ac, est = self.storage, self.actor_critic, self.estimator
        if s.step >= s.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        t, N, ws = s.step, s.num_envs, self.roll_ws
        ld = ac.ld_actor_in
        x = ws.get("actor_in", N, ld)
        obs, privileged_obs, critic_obs = _rows16(obs), _rows16(privileged_obs, need16=False), _rows16(critic_obs)
        true_estimated_obs, scan_obs = _rows16(true_estimated_obs, need16=False), _rows16(scan_obs)
        # `in_place`: the env wrote this step's rows straight into storage slot t (Go2Env.bind_output_rows on an aliased
        # storage): obs / priv / est / scan are column slices of that row and there is nothing to copy
        in_place = s.alias_critic_rows and critic_obs.data_ptr() == s.critic_observations[t].data_ptr()
        # estimated obs -> actor input (the rollout acts on the ESTIMATE, ppo.py:134-137); the estimator, the latent
        # encoder, the scan encoder and the critic are independent until the actor's first layer.  Estimator -> actor is the
        # critical path (high-priority stream); the critic's value is not needed before process_env_step, so with
        # `defer_critic_join` (the runner sets it) its chain keeps running under the env step's kernels.
        if self._pending_critic is not None:              # act() called twice without process_env_step
            self._join([self._pending_critic])
            self._pending_critic = None
        # The observation copies of RolloutStorage.add_transitions ride on the side streams: the estimator reads the env's
        # obs buffer in place, [obs -> actor input, priv -> storage] precede the latent encoder, the rest precedes the critic.
        s_hi, s_scan, s_lat, s_crit = self._fork(4)
        with self._on(s_hi):
            est.fwd(ws, _p(obs), obs.stride(0), _p(x, ac.col_est), ld, N)
        with self._on(s_lat):
            segs = [(_p(obs), obs.stride(0), _p(x), ld, s.d_obs)]
            if s.alias_critic_rows and not in_place:      # the row = critic_obs (assumed [obs | priv | est | scan], as the env builds it)
                segs.append((_p(critic_obs), critic_obs.stride(0), _p(s.critic_observations[t]), s.ld_crit, s.d_crit))
            elif not s.alias_critic_rows:
                segs.append((_p(privileged_obs), privileged_obs.stride(0), _p(s._priv[t]), s.ld_priv, s.d_priv))
            # in place, the latent encoder reads the privileged columns of the storage row: it goes first and the copy of the
            # proprioceptive columns (needed by the actor's first layer only) follows it
            first_priv = in_place and not adaptation_mode
            if first_priv:
                ac.fwd_priv(ws, _p(s._priv[t]), s.ld_priv, _p(x, ac.col_latent), ld, N)
            self._copy_segments(segs, N)
            if adaptation_mode:
                ac.fwd_adapt(ws, _p(x), ld, _p(x, ac.col_latent), ld, N)
            elif not first_priv:
                ac.fwd_priv(ws, _p(s._priv[t]), s.ld_priv, _p(x, ac.col_latent), ld, N)
        with self._on(s_scan):
            ac.fwd_scan(ws, _p(scan_obs), scan_obs.stride(0), _p(x, ac.col_scan), ld, N)
        mu = ws.get("mu", N, s.d_act)
        with self._on(s_hi):
            self._join([s_lat, s_scan])
            # the critic's big first layer must not take the SMs before this point; with a deferred store nothing on the
            # main stream waits for the critic, so it starts behind the actor's wide layers instead of beside them
            late = self.defer_store and self.defer_critic_join and s_crit is not None and self.critic_fork_after >= 0
            if not late:
                self._fork_onto([s_crit])
            ac.fwd_actor(ws, _p(x), ld, _p(mu), s.d_act, N,
                         after=(min(self.critic_fork_after, len(ac.actor) - 1), lambda: self._fork_onto([s_crit])) if late else None)
        with self._on(s_crit):
            if s.alias_critic_rows:
                if not in_place and s_lat is not None:    # the row copy above rides on the latent stream
                    self._join_onto(s_crit, [s_lat])
            else:
                self._copy_segments([
                    (_p(obs), obs.stride(0), _p(s.observations[t]), s.ld_obs, s.d_obs),
                    (_p(critic_obs), critic_obs.stride(0), _p(s.critic_observations[t]), s.ld_crit, s.d_crit),
                    (_p(true_estimated_obs), true_estimated_obs.stride(0), _p(s._est[t]), s.ld_est, s.d_est),
                    (_p(scan_obs), scan_obs.stride(0), _p(s.scan_observations[t]), s.ld_scan, s.d_scan),
                ], N)
            copied = None
            if s_crit is not None and not in_place:       # the env's buffers may be overwritten (env.step) once THIS has run,
                copied = torch.cuda.Event()               # even if the critic chain itself is still in flight
                copied.record()
            ac.fwd_critic(ws, _p(s.critic_observations[t]), s.ld_crit, _p(s.values[t]), 1, N)
        self._join([s_hi])
        if copied is not None:
            torch.cuda.current_stream().wait_event(copied)
        if self.defer_critic_join and s_crit is not None:
            self._pending_critic = s_crit
        else:
            self._join([s_crit])
        if self.use_device_counter:
            _lib.check(self.lib.b200_sample_actions_dev(_p(mu), s.d_act, ac.main.ptr("std"), self.seed, C.c_void_p(self.act_counter_dev.data_ptr()),
                                                        _p(s.actions[t]), _p(s.actions_log_prob[t]), _p(s.mu[t]), _p(s.sigma[t]), N, s.d_act,
                                                        _lib.stream_ptr()))
        else:
            _lib.check(self.lib.b200_sample_actions(_p(mu), s.d_act, ac.main.ptr("std"), self.seed, self.act_counter, _p(s.actions[t]),
                                                    _p(s.actions_log_prob[t]), _p(s.mu[t]), _p(s.sigma[t]), N, s.d_act, _lib.stream_ptr()))
        self.act_counter += 1
        self.transition.actions, self.transition.values = s.actions[t], s.values[t]
        self._wait_store()              # the caller's env.step may overwrite what the previous step's deferred store reads
        return s.actions[t]

    def bookkeeping_stream(self):
        """the stream the critic chain and (with `defer_store`) process_env_step run on; None without side streams"""
        return self._fork(4)[3]

    def _wait_store(self):
        if self._store_done is not None:
            torch.cuda.current_stream().wait_event(self._store_done)
            self._store_done = None

    def process_env_step(self, rewards, dones, infos):
        """ppo.py:156-171: time-out bootstrap + scalar part of add_transitions."""
        s = self.storage
        t = s.step
        side = self._pending_critic if self.defer_store else None
        if self._pending_critic is not None and side is None:   # values[t] (time-out bootstrap below) come from the critic's stream
            self._join([self._pending_critic])
        self._pending_critic = None
        tmo = infos["time_outs"] if "time_outs" in infos else None
        rewards, dones, tmo = self._as_step_scalars(rewards, dones, tmo)
        p = lambda x: C.c_void_p(x.data_ptr())
        if side is not None:
            self._wait_store()
            self._fork_onto([side])                          # the env step's outputs are queued on the main stream
        with self._on(side):
            _lib.check(self.lib.b200_store_step_scalars(p(rewards), p(dones), p(tmo) if tmo is not None else None, p(s.values[t]),
                                                        self.gamma, p(s.rewards[t]), p(s.dones[t]), s.num_envs, _lib.stream_ptr()))
            if side is not None:
                self._store_done = torch.cuda.Event()
                self._store_done.record()
        s.step += 1
        self.transition.clear()

    def _as_step_scalars(self, rewards, dones, time_outs):
        """the kernel reads fp32 rewards and 1-byte flags from raw addresses: convert what a duck-typed env hands over
        (int64 dones, float time_outs, strided views) instead of misreading it"""
        N = self.storage.num_envs

        def flag(x, name):
            if x is None:
                return None
            if x.dtype not in (torch.bool, torch.uint8):
                x = x != 0
            if not x.is_cuda or not x.is_contiguous():
                x = x.to(self.device).contiguous()
            if x.numel() != N:
                raise ValueError(f"process_env_step: {name} has {x.numel()} elements, expected {N}")
            return x
        if rewards.dtype != torch.float32 or not rewards.is_cuda or not rewards.is_contiguous():
            rewards = rewards.to(self.device, torch.float32).contiguous()
        if rewards.numel() != N:
            raise ValueError(f"process_env_step: rewards has {rewards.numel()} elements, expected {N}")
        return rewards, flag(dones, "dones"), flag(time_outs, "time_outs")

    def compute_returns(self, last_critic_obs):
        """ppo.py:174-179."""
        s, ac = self.storage, self.actor_critic
        self._wait_store()
        x = _rows16(last_critic_obs)
        ac.fwd_critic(self.roll_ws, _p(x), x.stride(0), _p(self.last_values), 1, s.num_envs)
        s.compute_returns(self.last_values, self.gamma, self.lam)

    # ---- update ---------------------------------------------------------------------------------------
    def _gather_storage(self, indices):
        """the one permutation of this update (rollout_storage.py:142) applied to every storage tensor."""
        s, ac, B = self.storage, self.actor_critic, self.batch
        st = _lib.stream_ptr
        g = lambda src, sld, dst, dld, w: _lib.check(self.lib.b200_gather_rows(src, sld, _p_i64(indices), dst, dld, w, B, st()))
        g(_p(s.observations), s.ld_obs, _p(self.p_actor_in), ac.ld_actor_in, s.d_obs)
        g(_p(s._est), s.ld_est, _p(self.p_actor_in, ac.col_est), ac.ld_actor_in, s.w_est)   # actor sees the TRUE value (ppo.py:199)
        g(_p(s._est), s.ld_est, _p(self.p_est), self.p_est.shape[1], s.w_est)
        g(_p(s._priv), s.ld_priv, _p(self.p_priv), self.p_priv.shape[1], s.w_priv)
        g(_p(s.critic_observations), s.ld_crit, _p(self.p_crit), s.d_crit, s.d_crit)
        g(_p(s.scan_observations), s.ld_scan, _p(self.p_scan), s.d_scan, s.d_scan)
        g(_p(s.actions), s.d_act, _p(self.p_act), s.d_act, s.d_act)
        for src, dst in ((s.values, self.p_val), (s.returns, self.p_ret), (s.actions_log_prob, self.p_logp), (s.advantages, self.p_adv)):
            g(_p(src), 1, _p(dst), 1, 1)
        if self.adaptive:
            g(_p(s.mu), s.d_act, _p(self.p_mu), s.d_act, s.d_act)
            g(_p(s.sigma), s.d_act, _p(self.p_sigma), s.d_act, s.d_act)
        # the adaptation-encoder latent slab (ppo.py:213-214, see init_storage): one launch over the whole batch
        ac.fwd_adapt(self.upd_ws, _p(self.p_actor_in), ac.ld_actor_in, _p(self.p_lat_a), ac.latent_dim, B)

    def _reg_coef(self):
        stage = min(max((self.total_updates - self.start_step) / self.duration, 0.0), 1.0)     # ppo.py:219-220
        return self.start_val + stage * (self.end_val - self.start_val)

    def _allreduce(self, group):
        from .dist import allreduce_flat_grads
        allreduce_flat_grads([group], self.process_group)

    def _adam(self, group):
        if getattr(group, "dist", None) is not None:      # exchange + clip + Adam + all-gather in one kernel
            group.dist.step()
            return
        self._allreduce(group)
        _lib.check(self.lib.b200_clip_adam(_p(group.params), _p(group.grads), _p(group.exp_avg), _p(group.exp_avg_sq), group.n,
                                           C.c_void_p(group.state.data_ptr()), 1.0 / self.world_size, self.max_grad_norm, 0.9, 0.999, 1e-8,
                                           _lib.stream_ptr()))

    def _adaptive_lr(self, mu, r0, M):
        """ppo.py:233-246: KL of the minibatch's old / new action distributions -> learning rate of the main optimiser."""
        A, acc = self.storage.d_act, C.c_void_p(self.kl_acc.data_ptr())
        _lib.check(self.lib.b200_kl_sum(_p(mu), A, self.actor_critic.main.ptr("std"), _p(self.p_mu) + 4 * r0 * A, A,
                                        _p(self.p_sigma) + 4 * r0 * A, A, M, A, acc, _lib.stream_ptr()))
        if self.process_group is not None:             # every rank must take the same decision: KL over all ranks' samples
            import torch.distributed as dist
            dist.all_reduce(self.kl_acc[0:1], op=dist.ReduceOp.SUM, group=self.process_group)
        _lib.check(self.lib.b200_adaptive_lr(acc, M * self.world_size, float(self.desired_kl),
                                             C.c_void_p(self.actor_critic.main.state.data_ptr()), _lib.stream_ptr()))

    def _minibatch(self, r0, M):
        """one PPO minibatch on rows [r0, r0+M) of the permuted slabs (ppo.py:194-276)."""
        ac, est, ws, k, s = self.actor_critic, self.estimator, self.upd_ws, self.actor_critic.k, self.storage
        ld = ac.ld_actor_in
        X = _p(self.p_actor_in) + 4 * r0 * ld
        priv, ldp = _p(self.p_priv) + 4 * r0 * self.p_priv.shape[1], self.p_priv.shape[1]
        crit, scan = _p(self.p_crit) + 4 * r0 * s.d_crit, _p(self.p_scan) + 4 * r0 * s.d_scan
        tgt_est, ldte = _p(self.p_est) + 4 * r0 * self.p_est.shape[1], self.p_est.shape[1]
        L, A = ac.latent_dim, s.d_act
        mu, val = ws.get("mu", M, A), ws.get("val", M, 4)
        pred, dpred = ws.get("pred", M, 4), ws.get("dpred", M, 4)
        SL = ac.scan_latent_dim
        assert ac.col_scan == ac.col_latent + L, "latent and scan-latent columns of the actor input must be adjacent"
        dmu, dval = ws.get("dmu", M, A), ws.get("dval", M, 4)
        dls = ws.get("dlatscan", M, L + SL)                  # d(loss)/d[latent | scan latent]: ONE dgrad of the actor's first layer
        lat_a_ptr = _p(self.p_lat_a) + 4 * r0 * L
        # Streams (the fork / join edges of the captured graph).  The critical path is encoders -> actor -> loss -> actor
        # backward -> encoder backward; it runs on HIGH-priority streams (s_hi; the scan encoder beside it on s_scan) and
        # is issued first, so the big estimator / critic GEMMs of the low-priority streams fill in around it instead of
        # delaying it (measured with tools/trace_update.py: the actor's first layer used to start at 169 us of 741).
        s_hi, s_scan, s_est, s_crit, s_w = self._fork(5)
        off = self._offload(s_w) if self.offload_wgrads else None
        with self._on(s_scan):
            ac.fwd_scan(ws, scan, s.d_scan, X + 4 * ac.col_scan, ld, M)
        with self._on(s_hi):
            ac.fwd_priv(ws, priv, ldp, X + 4 * ac.col_latent, ld, M)
        # estimator: forward, loss, backward, own optimiser (ppo.py:224-231) -- fully independent chain
        with self._capped(s_est, forward=True):
            est.fwd(ws, X, ld, _p(pred), 4, M)
            _lib.check(self.lib.b200_mse_rows_loss(_p(pred), 4, tgt_est, ldte, _p(dpred), 4, _p(self.loss_sums, 4), M, est.output_dim,
                                                   _lib.stream_ptr()))
        # Data parallel with schedule='adaptive': the KL all-reduce (NCCL, on the critical stream) and the estimator's fused
        # optimiser step are both kernels that WAIT for the peer ranks.  On unordered graph branches nothing stops two ranks
        # from starting them in opposite order, each then needing the other kernel co-resident to make progress; the
        # estimator's step is therefore issued behind the all-reduce, which gives the blocking collectives one total order on
        # every rank (KL all-reduce -> estimator step -> main step) whatever the GPU can co-schedule.
        est_step_after_kl = self.adaptive and self.process_group is not None
        with self._capped(s_est):
            chain_backward(est.k, est.layers, ws, "e", X + 4 * est.in_col, ld, _p(dpred), 4, M)
            if not est_step_after_kl:
                self._adam(est.group)
        with self._capped(s_crit, forward=True):
            ac.fwd_critic(ws, crit, s.d_crit, _p(val), 4, M)
        with self._on(s_hi):
            self._join([s_scan])
            ac.fwd_actor(ws, X, ld, _p(mu), A, M)
            if self.adaptive:
                self._adaptive_lr(mu, r0, M)
            if est_step_after_kl:
                self._fork_onto([s_est])
                with self._capped(s_est):
                    self._adam(est.group)
            self._join([s_crit])
            # PPO loss head (ppo.py:249-270)
            a = _lib.PpoLossArgs()
            a.mu, a.ldmu, a.std, a.actions = _p(mu), A, ac.main.ptr("std"), _p(self.p_act) + 4 * r0 * A
            a.old_logp, a.adv = _p(self.p_logp) + 4 * r0, _p(self.p_adv) + 4 * r0
            a.returns, a.target_values = _p(self.p_ret) + 4 * r0, _p(self.p_val) + 4 * r0
            a.value, a.ldv = _p(val), 4
            a.latent_p, a.ldlp, a.latent_a, a.ldla = X + 4 * ac.col_latent, ld, lat_a_ptr, L
            a.dmu, a.lddmu, a.dvalue, a.lddv, a.dlatent_p, a.lddlp = _p(dmu), A, _p(dval), 4, _p(dls), L + SL
            a.dstd, a.sums = ac.main.ptr("std", "grads"), _p(self.loss_sums)
            a.M, a.A, a.L = M, A, L
            a.clip, a.value_coef, a.entropy_coef, a.reg_coef = self.clip_param, self.value_loss_coef, self.entropy_coef, 0.0
            a.use_clipped_value_loss, a.reg_coef_dev = int(self.use_clipped_value_loss), _p(self.reg_coef_dev)
            _lib.check(self.lib.b200_ppo_loss(C.byref(a), _lib.stream_ptr()))
            # backward: critic on its own stream; actor (input gradient only for the latent / scan-latent columns) here
            self._fork_onto([s_crit])
        with self._capped(s_crit):
            chain_backward(k, ac.critic, ws, "c", crit, s.d_crit, _p(dval), 4, M)
        with self._on(s_hi):
            chain_backward(k, ac.actor, ws, "a", X, ld, _p(dmu), A, M, wgrad_on=off)
            da0, lda0 = ws.ptr("da0", M, ceil4(ac.actor[0].N)), ceil4(ac.actor[0].N)
            # columns [latent | scan latent] of the first layer's input gradient in one pass; the regulariser's gradient, which
            # the loss head left in the first L columns, is accumulated (accumulate = number of leading columns)
            k.dgrad(ac.actor[0], da0, lda0, None, 0, _p(dls), L + SL, M, accumulate=L, wcol=ac.col_latent, K=L + SL)
            self._fork_onto([s_scan])
            chain_backward(k, ac.priv, ws, "p", priv, ldp, _p(dls), L + SL, M, wgrad_on=off)
        with self._on(s_scan):
            chain_backward(k, ac.scan, ws, "s", scan, s.d_scan, _p(dls) + 4 * L, L + SL, M, wgrad_on=off)
        self._join([s_hi, s_scan, s_crit, s_est, s_w])
        self._adam(ac.main)

    # ---- side streams: estimator / critic / adaptation-encoder chains are independent of the actor chain until the loss
    #      head (and, for the backward, until Adam); forking them lets the small and medium kernels overlap.  Under
    #      CUDA-graph capture the event waits become the fork / join edges of the graph.
    def _fork(self, n):
        """side streams, in this order: high-priority critical chain, high-priority scan encoder, estimator (update) / latent
        encoder (rollout), critic"""
        if not self.use_streams:
            return [None] * n
        if not hasattr(self, "_side"):        # one pool per device for the whole process, however many PPO objects exist
            key = torch.device(self.device).index or 0
            if key not in _SIDE_STREAMS:
                lo, hi = 0, -1
                _SIDE_STREAMS[key] = [torch.cuda.Stream(device=self.device, priority=p) for p in (hi, hi, lo, lo, lo)]
            self._side = _SIDE_STREAMS[key]
        self._fork_onto(self._side[:n])
        return self._side[:n]

    def _fork_onto(self, streams):
        """make side streams wait for everything queued on the current stream so far"""
        if not self.use_streams:
            return
        ev = torch.cuda.Event()
        ev.record()
        for s in streams:
            if s is not None:
                s.wait_event(ev)

    def _on(self, stream):
        import contextlib
        return torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()

    def _offload(self, stream):
        """-> run(fn): launch `fn`'s kernels on `stream` (capped like the other low-priority chains), ordered after everything
        queued on the CURRENT stream so far; None when there is no side stream (then callers launch inline)"""
        if stream is None:
            return None

        def run(fn):
            ev = torch.cuda.Event()
            ev.record()
            stream.wait_event(ev)
            with self._capped(stream):
                fn()
        return run

    def _capped(self, stream, forward=False):
        """`_on(stream)` for a LOW-priority chain of the update.  Its GEMM launches are sized for `side_sm_cap` SMs: the cap is
        a property of the stream inside the library (b200_tc_set_stream_sm_cap), (re)registered here whenever it changed."""
        cap = self.side_sm_cap if (self.use_streams and stream is not None) else 0
        if stream is not None and _REGISTERED_CAPS.get(stream.cuda_stream) != cap:
            _lib.check(self.lib.b200_tc_set_stream_sm_cap(C.c_void_p(stream.cuda_stream), int(cap or 0)))
            _REGISTERED_CAPS[stream.cuda_stream] = cap
        return self._on(stream)

    def _join_onto(self, stream, streams):
        """make `stream` wait for everything queued on `streams` so far"""
        for s in streams:
            if s is not None and stream is not None:
                ev = torch.cuda.Event()
                ev.record(s)
                stream.wait_event(ev)

    def _join(self, streams):
        for s in streams:
            if s is not None:
                ev = torch.cuda.Event()
                ev.record(s)
                torch.cuda.current_stream().wait_event(ev)

    # ---- CUDA graphs: a minibatch is ~90 launches with fixed pointers -> capture once per minibatch slot, replay ----
    def set_device_counter(self, enabled=True):
        self.act_counter_dev[0:1].fill_(self.act_counter)
        self.use_device_counter = bool(enabled)

    def _run_captured(self, key, fn):
        """1st call eager (allocates workspaces, sets kernel attributes), 2nd call captures + replays, then replays."""
        if not self.use_graphs:
            return fn()
        n = self._graph_calls.get(key, 0)
        self._graph_calls[key] = n + 1
        if n == 0:
            return fn()
        if key not in self._graphs:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            self._graphs[key] = g
        self._graphs[key].replay()

    def update(self):
        """ppo.py:182-293 -> (value_loss, surrogate_loss, reg_loss, reg_coef, estimator_loss)."""
        self._wait_store()
        indices = torch.randperm(self.num_mini_batches * self.mb, device=self.device, generator=self._perm_gen)
        return self.update_with_indices(indices)

    def update_with_indices(self, indices):
        self._gather_storage(indices)
        reg_coef = self._reg_coef()
        self.reg_coef_dev.fill_(reg_coef)
        self.loss_sums.zero_()
        for _ in range(self.num_learning_epochs):
            for i in range(self.num_mini_batches):
                self._run_captured(("ppo", i), lambda i=i: self._minibatch(i * self.mb, self.mb))
        n = self.num_learning_epochs * self.num_mini_batches
        sums = (self.loss_sums / (n * self.mb)).tolist()              # the only host read-back of the update
        self.storage.clear()
        self.total_updates += 1
        self.enforce_max_std(1.0)
        # sums: [surrogate, value, reg, entropy, estimator]
        return sums[1], sums[0], sums[2], reg_coef, sums[4]

    def enforce_max_std(self, max_action_std=1.0):
        self.actor_critic.std.clamp_(max=max_action_std)             # ppo.py:301-307

    def increase_update_count(self):
        self.total_updates += 1

    def _dagger_minibatch(self, r0, M):
        """ppo.py:318-339: regress the adaptation encoder onto the (detached) privileged latent."""
        ac, ws, s = self.actor_critic, self.upd_ws, self.storage
        ld, L = ac.ld_actor_in, ac.latent_dim
        X = _p(self.p_actor_in) + 4 * r0 * ld
        priv, ldp = _p(self.p_priv) + 4 * r0 * self.p_priv.shape[1], self.p_priv.shape[1]
        lat_p, lat_a, dlat_a = ws.get("lat_p", M, L), ws.get("lat_a", M, L), ws.get("dlat_a", M, L)
        ac.fwd_priv(ws, priv, ldp, _p(lat_p), L, M)
        ac.fwd_adapt(ws, X, ld, _p(lat_a), L, M, save=True)
        _lib.check(self.lib.b200_l2_rows_loss(_p(lat_a), L, _p(lat_p), L, _p(dlat_a), L, _p(self.loss_sums, 5), M, L, _lib.stream_ptr()))
        ac.bwd_adapt(ws, X, ld, _p(dlat_a), L, _p(lat_a), L, M)
        self._adam(ac.adapt)

    def update_dagger(self):
        self._wait_store()
        indices = torch.randperm(self.num_mini_batches * self.mb, device=self.device, generator=self._perm_gen)
        return self.update_dagger_with_indices(indices)

    def update_dagger_with_indices(self, indices):
        self._gather_storage(indices)
        self.loss_sums.zero_()
        for _ in range(self.num_learning_epochs):
            for i in range(self.num_mini_batches):
                self._run_captured(("dagger", i), lambda i=i: self._dagger_minibatch(i * self.mb, self.mb))
        n = self.num_learning_epochs * self.num_mini_batches
        loss = float(self.loss_sums[5].item()) / (n * self.mb)
        # ppo.py:274 clips actor_critic.parameters(): the adaptation encoder's .grad left by THIS update (post-clip, never
        # zeroed by optimizer.zero_grad()) is part of the norm of every later PPO minibatch -- hand its squared norm over
        ac = self.actor_critic
        ac.main.state[7:8].copy_(ac.adapt.state[6:7])
        self.storage.clear()
        self.total_updates += 1
        return loss


def _p_i64(t):
    return C.c_void_p(t.data_ptr())


_SIDE_STREAMS = {}      # device index -> [critical, scan, estimator / latent, critic, wgrad offload] streams
_REGISTERED_CAPS = {}   # cuda stream -> SM cap registered with the library (b200_tc_set_stream_sm_cap)


def _rows16(t, need16=True):
    """a [rows, cols] fp32 CUDA tensor the kernels can read in place: unit column stride and (need16: it feeds a TMA / 16-byte
    vector path) a 16-byte aligned base and row pitch -- column slices of a wider row qualify; anything else is copied"""
    ok = t.dim() == 2 and t.dtype == torch.float32 and t.is_cuda and (t.shape[1] == 1 or t.stride(1) == 1)
    if ok and need16:
        ok = t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0
    return t if ok else t.to(torch.float32).contiguous()
