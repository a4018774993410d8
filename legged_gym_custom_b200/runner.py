"""`OnPolicyRunner`: drop-in for rsl_rl/runners/on_policy_runner.py (same constructor, `learn`, `save`,
`load`, `get_inference_policy`), driving `Go2Env` and the kernel-backed `PPO`.

Differences that are deliberate: the finished episodes of a rollout are collected on the
device and read back once per iteration instead of `.cpu().numpy()` every env step -- `rewbuffer` / `lenbuffer` still hold
the last 100 finished EPISODES in the reference's order (on_policy_runner.py:163-169) -- and checkpoints additionally
carry the estimator, the auxiliary optimisers and `total_updates` (the reference silently drops them); the
reference's keys ('model_state_dict', 'optimizer_state_dict', 'iter', 'infos') are unchanged.
"""
import os
import time
from collections import deque

import torch

from .learner import PPO
from .networks import ActorCritic, MlpEstimator


class OnPolicyRunner:
    def __init__(self, env, train_cfg, log_dir=None, device="cuda:0", process_group=None, precise=False):
        self.cfg, self.alg_cfg, self.policy_cfg = train_cfg["runner"], train_cfg["algorithm"], train_cfg["policy"]
        self.device, self.env = torch.device(device), env
        pc, ac_ = self.policy_cfg, self.alg_cfg
        seed = int(train_cfg.get("seed", 1))
        # networks are initialised from the SHARED seed (and broadcast from rank 0 below); exploration noise and the minibatch
        # permutation are keyed per rank, so env e of every rank does not draw the same action noise
        from .dist import shard_seed
        rank = torch.distributed.get_rank(process_group) if process_group is not None else 0
        rank_seed = shard_seed(seed, rank)
        actor_critic = ActorCritic(num_proprio=env.num_proprio, num_privileged_obs=env.num_privileged_obs,
                                   num_critic_obs=env.num_critic_obs, num_estimated_obs=env.num_estimated_obs,
                                   num_scan_obs=env.num_scan_obs, num_actions=env.num_actions,
                                   history_buffer_length=env.history_buffer_length, actor_hidden_dims=pc["actor_hidden_dims"],
                                   critic_hidden_dims=pc["critic_hidden_dims"], priv_encoder_hidden_dims=pc["priv_encoder_hidden_dims"],
                                   scan_encoder_hidden_dims=pc["scan_encoder_hidden_dims"],
                                   latent_encoder_output_dim=pc["latent_encoder_output_dim"],
                                   scan_encoder_output_dim=pc["scan_encoder_output_dim"], activation=pc["activation"],
                                   init_noise_std=pc["init_noise_std"], device=self.device, seed=seed, precise=precise,
                                   learning_rate=ac_["learning_rate"])
        estimator = MlpEstimator(num_proprio=env.num_proprio, history_buffer_length=env.history_buffer_length,
                                 output_dim=env.num_estimated_obs, hidden_dims=pc["estimator_hidden_dims"], activation=pc["activation"],
                                 use_history=pc["use_history"], device=self.device, seed=seed + 1, precise=precise,
                                 learning_rate=ac_["estimator_learning_rate"])
        self.alg = PPO(actor_critic=actor_critic, estimator=estimator, num_learning_epochs=ac_["num_learning_epochs"],
                       num_mini_batches=ac_["num_mini_batches"], clip_param=ac_["clip_param"], gamma=ac_["gamma"], lam=ac_["lam"],
                       value_loss_coef=ac_["value_loss_coef"], entropy_coef=ac_["entropy_coef"], learning_rate=ac_["learning_rate"],
                       estimator_learning_rate=ac_["estimator_learning_rate"], max_grad_norm=ac_["max_grad_norm"],
                       use_clipped_value_loss=ac_["use_clipped_value_loss"], schedule=ac_["schedule"], desired_kl=ac_["desired_kl"],
                       resume=self.cfg["resume"], device=self.device, seed=rank_seed, process_group=process_group)
        if process_group is not None:      # replicas start from rank 0's weights; afterwards identical reduced gradients keep them in sync
            from .dist import broadcast_parameters
            broadcast_parameters([actor_critic.main, actor_critic.adapt, estimator.group], process_group)
        self.alg.defer_critic_join = True      # every act() of this runner is followed by process_env_step()
        self.alg.defer_store = True            # ... so its bookkeeping may ride on the critic's stream (learner.PPO.defer_store)
        if hasattr(env, "extras_stream"):      # and the env's episode statistics / time-out copy with it (Go2Env.extras_stream)
            env.extras_stream = self.alg.bookkeeping_stream()
        self.dagger_update_freq = ac_["dagger_update_freq"]
        self.num_steps_per_env, self.save_interval = self.cfg["num_steps_per_env"], self.cfg["save_interval"]
        # an env that can write its observation rows anywhere (Go2Env with alias_outputs) writes them straight into the
        # rollout storage: step t's output is slot t + 1, the transition of the NEXT act() -- no observation copies at all
        self.in_place_rows = bool(getattr(env, "supports_output_binding", False))
        self.alg.init_storage(num_envs=env.num_envs, num_transitions_per_env=self.num_steps_per_env, total_obs_shape=[env.num_obs],
                              privileged_obs_shape=[env.num_privileged_obs], critic_obs_shape=[env.num_critic_obs],
                              estimated_obs_shape=[env.num_estimated_obs], scan_obs_shape=[env.num_scan_obs],
                              action_shape=[env.num_actions], alias_critic_rows=self.in_place_rows)
        self.log_dir, self.writer = log_dir, None
        self.tot_timesteps, self.tot_time, self.current_learning_iteration = 0, 0, 0
        if self.in_place_rows:                     # the reset's observation is "the row after the last transition"
            self.env.bind_output_rows(self.alg.storage.rows[self.num_steps_per_env])
        self.env.reset()
        N = env.num_envs
        self._cur_rew, self._cur_len = torch.zeros(N, device=self.device), torch.zeros(N, device=self.device)
        # finished episodes of one rollout, [T, N] with NaN where no episode ended: read back ONCE per iteration and appended
        # to the 100-episode buffers in the reference's order (step-major, env-minor; on_policy_runner.py:163-169)
        T = self.num_steps_per_env
        self._fin_rew = torch.full((T, N), float("nan"), device=self.device)
        self._fin_len = torch.full((T, N), float("nan"), device=self.device)
        self.rewbuffer, self.lenbuffer = deque(maxlen=100), deque(maxlen=100)
        self.last_losses = {}

    # ---- CUDA graphs ----------------------------------------------------------------------------------------
    def enable_graphs(self, enabled=True):
        """Capture the whole rollout (T x [policy inference, 4 PD substeps, post-physics, storage writes] + GAE) and every
        minibatch of the update as CUDA graphs and replay them: ~2 400 launches per iteration become a handful of
        graph launches.  Step counters move to device memory so replays see advancing step numbers; the PhysX frame
        ring must divide T (the replay re-reads the frames baked at capture time in the same order)."""
        ring = len(getattr(self.env.physx, "frames", [None]))
        if enabled and self.num_steps_per_env % ring != 0:
            raise ValueError("PhysX frame ring must divide num_steps_per_env for graph replay")
        self.use_graphs = bool(enabled)
        self.env.set_device_counter(enabled)
        self.alg.set_device_counter(enabled)
        self.alg.use_graphs = bool(enabled)
        self._rollout_graphs, self._rollout_calls = {}, {}

    def release_graphs(self):
        """destroy every captured graph (rollout and minibatch slots).  Call before torch.distributed.destroy_process_group():
        graphs that captured NCCL work (schedule='adaptive' under data parallelism) keep the communicator busy and the
        teardown would wait for them forever."""
        torch.cuda.synchronize()
        self._rollout_graphs, self._rollout_calls = {}, {}
        self.alg._graphs, self.alg._graph_calls = {}, {}
        import gc
        gc.collect()
        torch.cuda.synchronize()

    def capture_graphs(self):
        """Capture EVERY graph a training run replays -- both rollout graphs (adaptation_mode False / True) and every
        ("ppo", i) / ("dagger", i) minibatch graph -- now, instead of lazily on the second use of each (the adaptation-mode
        rollout would otherwise be captured at iteration `dagger_update_freq`, in the middle of a run).  Parameters, optimiser
        state and the update counter are restored afterwards, so the learner starts from where it was; the env has advanced by
        four rollouts (set-up steps, like the reference's own warm-up reset)."""
        if not getattr(self, "use_graphs", False):
            raise RuntimeError("capture_graphs: call enable_graphs() first")
        alg = self.alg
        groups = [alg.actor_critic.main, alg.actor_critic.adapt, alg.estimator.group]
        saved = [(g.params.clone(), g.exp_avg.clone(), g.exp_avg_sq.clone(), g.state.clone()) for g in groups]
        updates, perm_state = alg.total_updates, alg._perm_gen.get_state()
        for mode in (True, False):
            for _ in range(2):                      # 1st call eager (allocations, kernel attributes), 2nd captures + replays
                self.rollout(mode)
                alg.update_dagger() if mode else alg.update()
        for g, (p, m, v, s) in zip(groups, saved):
            g.params.copy_(p); g.exp_avg.copy_(m); g.exp_avg_sq.copy_(v); g.state.copy_(s)
        alg.total_updates = updates
        alg._perm_gen.set_state(perm_state)
        torch.cuda.synchronize()

    def rollout(self, use_adaptation_mode):
        if not getattr(self, "use_graphs", False):
            return self._rollout_eager(use_adaptation_mode)
        key = bool(use_adaptation_mode)
        n = self._rollout_calls.get(key, 0)
        self._rollout_calls[key] = n + 1
        if n == 0:
            return self._rollout_eager(use_adaptation_mode)          # warm-up: allocations, kernel attributes
        T = self.num_steps_per_env
        if key not in self._rollout_graphs:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._rollout_eager(use_adaptation_mode)              # host-side counters advance here, kernels do not run yet
            self._rollout_graphs[key] = g
        else:                                                         # replay: advance the host mirrors by hand
            self.env.common_step_counter += T
            self.alg.act_counter += T
            self.alg.storage.step = T
            if hasattr(self.env.physx, "cursor"):
                self.env.physx.cursor = (self.env.physx.cursor + T) % len(self.env.physx.frames)
            if hasattr(self.env.physx, "h2d_bytes"):
                self.env.physx.h2d_bytes += self.env.physx.bytes_per_step * T
        self._rollout_graphs[key].replay()

    # ---- one iteration = rollout + GAE + update (on_policy_runner.py:144-194) -----------------------------
    def _rollout_eager(self, use_adaptation_mode):
        env, alg = self.env, self.alg
        if self.in_place_rows:                     # the last row of the previous rollout is the first of this one
            rows = alg.storage.rows
            rows[0].copy_(rows[self.num_steps_per_env])
            env.bind_output_rows(rows[0])
        obs, priv, crit = env.get_observations(), env.get_privileged_observations(), env.get_critic_observations()
        est, scan = env.get_estimated_observations(), env.get_scan_observations()
        for t in range(self.num_steps_per_env):
            actions = alg.act(obs, priv, crit, est, scan, adaptation_mode=use_adaptation_mode)
            if self.in_place_rows:
                env.bind_output_rows(rows[t + 1])
            obs, priv, crit, est, scan, rewards, dones, infos = env.step(actions)
            alg.process_env_step(rewards, dones, infos)
            if self.log_dir is not None:               # book keeping of on_policy_runner.py:160-169, without its per-step read-back
                self._cur_rew += rewards
                self._cur_len += 1
                done = dones.bool()
                nan = torch.full_like(self._cur_rew, float("nan"))
                self._fin_rew[t] = torch.where(done, self._cur_rew, nan)
                self._fin_len[t] = torch.where(done, self._cur_len, nan)
                self._cur_rew.masked_fill_(done, 0.0)
                self._cur_len.masked_fill_(done, 0.0)
        alg.compute_returns(crit)
        if hasattr(env, "wait_extras"):
            env.wait_extras()

    def iteration(self, it):
        use_adaptation_mode = it % self.dagger_update_freq == 0
        self.rollout(use_adaptation_mode)
        if use_adaptation_mode:
            self.last_losses = {"adaptation": self.alg.update_dagger()}
        else:
            v, s, r, c, e = self.alg.update()
            self.last_losses = {"value": v, "surrogate": s, "regularization": r, "reg_coef": c, "estimator": e}
        return self.last_losses

    def learn(self, num_learning_iterations, init_at_random_ep_len=False):
        if self.log_dir is not None and self.writer is None:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.writer = SummaryWriter(log_dir=self.log_dir, flush_secs=10)
            except Exception:
                self.writer = None
        if init_at_random_ep_len:
            self.env.episode_length_buf = torch.randint_like(self.env.episode_length_buf, high=int(self.env.max_episode_length))
        tot = self.current_learning_iteration + num_learning_iterations
        for it in range(self.current_learning_iteration, tot):
            start = time.time()
            losses = self.iteration(it)
            torch.cuda.synchronize()
            dt = time.time() - start
            self.tot_time += dt
            self.tot_timesteps += self.num_steps_per_env * self.env.num_envs
            if self.log_dir is not None:
                self.log(it, losses, dt)
                if it % self.save_interval == 0:
                    self.save(os.path.join(self.log_dir, f"model_{it}.pt"))
        self.current_learning_iteration += num_learning_iterations
        if self.log_dir is not None:
            self.save(os.path.join(self.log_dir, f"model_{self.current_learning_iteration}.pt"))

    def log(self, it, losses, dt):
        fin = torch.stack((self._fin_rew, self._fin_len)).cpu()      # the one read-back of the iteration's episode statistics
        ended = ~torch.isnan(fin[0])
        self.rewbuffer.extend(fin[0][ended].tolist())              # row-major = step-major, env-minor: the reference's order
        self.lenbuffer.extend(fin[1][ended].tolist())
        fps = int(self.num_steps_per_env * self.env.num_envs / dt)
        if self.writer is not None:
            for k, v in losses.items():
                self.writer.add_scalar("Loss/" + k, v, it)
            self.writer.add_scalar("Loss/learning_rate", self.alg.learning_rate, it)
            self.writer.add_scalar("Perf/total_fps", fps, it)
            self.writer.add_scalar("Policy/mean_noise_std", float(self.alg.actor_critic.std.mean()), it)
            for k, v in self.env.extras.get("episode", {}).items():
                self.writer.add_scalar("Episode/" + k, float(v), it)
            if self.rewbuffer:
                self.writer.add_scalar("Train/mean_reward", sum(self.rewbuffer) / len(self.rewbuffer), it)
                self.writer.add_scalar("Train/mean_episode_length", sum(self.lenbuffer) / len(self.lenbuffer), it)
        body = " ".join(f"{k}={v:.4f}" for k, v in losses.items())
        print(f"it {it} fps {fps} {body}" + (f" mean_reward {self.rewbuffer[-1]:.3f}" if self.rewbuffer else ""))

    def save(self, path, infos=None):
        alg = self.alg
        torch.save({"model_state_dict": alg.actor_critic.state_dict(), "optimizer_state_dict": alg.optimizer.state_dict(),
                    "iter": self.current_learning_iteration, "infos": infos,
                    "estimator_state_dict": alg.estimator.state_dict(),
                    "estimator_optimizer_state_dict": alg.estimator_optimizer.state_dict(),
                    "adaptation_optimizer_state_dict": alg.adaptation_optimizer.state_dict(), "total_updates": alg.total_updates}, path)

    def load(self, path, load_optimizer=True):
        d = torch.load(path, map_location="cpu")
        alg = self.alg
        alg.actor_critic.load_state_dict(d["model_state_dict"])
        if "estimator_state_dict" in d:
            alg.estimator.load_state_dict(d["estimator_state_dict"])
        if load_optimizer and isinstance(d.get("optimizer_state_dict"), dict) and "flat_exp_avg" in d["optimizer_state_dict"]:
            alg.optimizer.load_state_dict(d["optimizer_state_dict"])
            if "estimator_optimizer_state_dict" in d:
                alg.estimator_optimizer.load_state_dict(d["estimator_optimizer_state_dict"])
                alg.adaptation_optimizer.load_state_dict(d["adaptation_optimizer_state_dict"])
        alg.total_updates = d.get("total_updates", alg.total_updates)
        self.current_learning_iteration = d["iter"]
        return d["infos"]

    def get_inference_policy(self, device=None):
        return self.alg.actor_critic.act_inference


def class_to_dict(obj):
    """helpers.py:41-56 for the class-namespace configs."""
    if not hasattr(obj, "__dict__"):
        return obj
    out = {}
    for key in dir(obj):
        if key.startswith("_"):
            continue
        val = getattr(obj, key)
        out[key] = [class_to_dict(v) for v in val] if isinstance(val, list) else class_to_dict(val)
    return out
