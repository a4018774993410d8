"""CPU oracle for the learner half of the hot path (rsl_rl): GAE, the time-out bootstrap, the
networks and the PPO / DAgger losses.

TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this file.

Plain fp32 torch-CPU restatement; each function cites the reference lines it follows.  Pinned by
tests/test_learner_oracle_vs_reference.py, which (in the authoring container, where
/root/reference exists) runs the reference's own RolloutStorage / ActorCritic / PPO on the same
tensors and requires identical results, and by tests/golden/learner_*.npz on the GPU box.
"""
import torch
import torch.nn.functional as F


def compute_returns(rewards, dones, values, last_values, gamma, lam):
    """rollout_storage.py:110-124.  rewards/values [T,N,1] fp32, dones [T,N,1] uint8, last_values [N,1]."""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    advantage = 0
    for step in reversed(range(T)):
        next_values = last_values if step == T - 1 else values[step + 1]
        not_terminal = 1.0 - dones[step].float()
        delta = rewards[step] + not_terminal * gamma * next_values - values[step]
        advantage = delta + not_terminal * gamma * lam * advantage
        returns[step] = advantage + values[step]
    adv = returns - values
    adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    return returns, adv


def bootstrap_rewards(rew, values, time_outs, gamma):
    """ppo.py:160-166: rewards += gamma * squeeze(values * time_outs.unsqueeze(1), 1)."""
    out = rew.clone()
    if time_outs is not None:
        out += gamma * torch.squeeze(values * time_outs.unsqueeze(1), 1)
    return out


# ---- networks, restated functionally over reference-layout state dicts --------------------------
import math

import numpy as np

from . import philox


# ---- numerics of the PRODUCTION kernels (precise=False): TF32 operands, fp32 accumulation ---------------------------
# The reference trains with torch.set_float32_matmul_precision('high') (scripts/train.py:39): TF32 matmuls.  The product's
# tcgen05 kernels (kind::tf32) consume the fp32 bits directly -- the tensor core ignores the low 13 mantissa bits
# (truncation); its mma.sync kernels (layers narrower than 8) convert with cvt.rna (round to nearest, ties away); the 1- /
# 3-wide heads' dgrad and the adaptation encoder's forward are plain fp32.  `numerics("tf32")` makes every Linear of this
# file reproduce exactly that (operands rounded per kernel, fp32 accumulation), so the production path has a comparator of
# its own precision; the default "fp32" is the reference's CPU arithmetic, bit for bit.
import contextlib

_NUMERICS = ["fp32"]


@contextlib.contextmanager
def numerics(mode):
    assert mode in ("fp32", "tf32")
    _NUMERICS.append(mode)
    try:
        yield
    finally:
        _NUMERICS.pop()


def tf32_trunc(x):
    """fp32 -> TF32 by dropping the low 13 mantissa bits (what tcgen05.mma kind::tf32 sees of an fp32 operand)"""
    return (x.contiguous().view(torch.int32) & -8192).view(torch.float32)


def tf32_rna(x):
    """cvt.rna.tf32.f32: round to nearest, ties away from zero (sign-magnitude: add half an ulp to the magnitude, truncate)"""
    return ((x.contiguous().view(torch.int32) + 4096) & -8192).view(torch.float32)


_ROUND = {"trunc": tf32_trunc, "rna": tf32_rna, "fp32": lambda x: x}


CHAIN_MAX_ROWS = 8192
_CHAIN = [False]      # Kernels.use_chain of the product under test (off by default): forward chains as one tcgen05 launch


def production_gemm_modes(M, N, K):
    """operand rounding of (forward, dgrad, wgrad) of a Linear [K -> N] on M rows, as legged_gym_custom_b200.networks
    dispatches it with precise=False: tcgen05 when N >= 8 and K >= 8 (wgrad: and M >= 32), else mma.sync; the forward and
    the dgrad of a head with N <= 4 are fp32 dot / outer products; with the optional one-launch chains (b200_tc_mlp_forward)
    the FORWARD of a narrow head runs on tcgen05 instead"""
    tc = N >= 8 and K >= 8
    fwd = "trunc" if (tc or (_CHAIN[0] and K >= 8 and M <= CHAIN_MAX_ROWS)) else ("fp32" if N <= 4 else "rna")
    return (fwd, "trunc" if tc else ("fp32" if N <= 4 else "rna"), "trunc" if (tc and M >= 32) else "rna")


class _LinearTF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, fwd, dgrad, wgrad):
        ctx.save_for_backward(x, w)
        ctx.modes = (dgrad, wgrad)
        y = _ROUND[fwd](x) @ _ROUND[fwd](w).t()
        return y + b if b is not None else y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dgrad, wgrad = ctx.modes
        dy = dy.contiguous()
        dx = _ROUND[dgrad](dy) @ _ROUND[dgrad](w)
        dw = _ROUND[wgrad](dy).t() @ _ROUND[wgrad](x)
        return dx, dw, dy.sum(0), None, None, None


def linear(x, w, b, forward_mode=None):
    """F.linear in the current numerics; `forward_mode` overrides the forward rounding (the adaptation encoder's fused
    forward kernel is fp32 while its backward runs on the TF32 GEMM kernels)"""
    if _NUMERICS[-1] == "fp32":
        return F.linear(x, w, b)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    f, d, g = production_gemm_modes(x2.shape[0], w.shape[0], w.shape[1])
    return _LinearTF32.apply(x2, w, b, forward_mode or f, d, g).reshape(*lead, w.shape[0])


def mlp(sd, prefix, x, n_layers, final_act=False):
    """nn.Sequential(Linear, ELU, ..., Linear) with keys `<prefix>.<2i>.{weight,bias}`
    (actor_critic.py:84-108, support_networks.py:24-35, :70-80, :108-116)."""
    for i in range(n_layers):
        x = linear(x, sd[f"{prefix}.{2 * i}.weight"], sd[f"{prefix}.{2 * i}.bias"])
        if i < n_layers - 1 or final_act:
            x = F.elu(x)
    return x


def n_layers(sd, prefix):
    return len([k for k in sd if k.startswith(prefix + ".") and k.endswith(".weight")])


def adaptation_encoder(sd, obs, num_proprio=52, hist=10):
    """AdaptationEncoder.forward on obs[:, :-num_proprio] (actor_critic.py:174-180, support_networks.py:128-175)."""
    h = obs[:, :-num_proprio].reshape(-1, hist, num_proprio)
    p = "adaptation_encoder_."
    if _NUMERICS[-1] == "tf32":
        return _adaptation_encoder_tf32(sd, h)
    x = F.elu(F.linear(h, sd[p + "fc_encoder.0.weight"], sd[p + "fc_encoder.0.bias"])).permute(0, 2, 1)
    x = F.elu(F.conv1d(x, sd[p + "conv_layers.0.weight"], sd[p + "conv_layers.0.bias"], stride=2))
    x = F.elu(F.conv1d(x, sd[p + "conv_layers.2.weight"], sd[p + "conv_layers.2.bias"], stride=1))
    return F.elu(F.linear(x.flatten(1), sd[p + "fc_final.0.weight"], sd[p + "fc_final.0.bias"]))


def _adaptation_encoder_tf32(sd, h):
    """the same network with each Conv1d written as the GEMM over its windows that the product runs (networks.py): the fused
    forward kernel is fp32, the backward (DAgger) goes through the TF32 GEMM kernels window by window"""
    p = "adaptation_encoder_."
    B = h.shape[0]
    x = F.elu(linear(h, sd[p + "fc_encoder.0.weight"], sd[p + "fc_encoder.0.bias"], forward_mode="fp32"))        # [B, 10, 30]
    w1, w2 = sd[p + "conv_layers.0.weight"], sd[p + "conv_layers.2.weight"]                                      # [20,30,4], [10,20,2]
    win = torch.stack([x[:, 2 * t:2 * t + 4, :] for t in range(4)], 1)                                           # [B, 4, k=4, 30]
    x = F.elu(linear(win.reshape(B, 4, 120), w1.permute(0, 2, 1).reshape(20, 120), sd[p + "conv_layers.0.bias"], forward_mode="fp32"))
    win = torch.stack([x[:, t:t + 2, :] for t in range(3)], 1)                                                   # [B, 3, k=2, 20]
    x = F.elu(linear(win.reshape(B, 3, 40), w2.permute(0, 2, 1).reshape(10, 40), sd[p + "conv_layers.2.bias"], forward_mode="fp32"))
    # Flatten of [B, 10 channels, 3 steps] is channel-major: column c * 3 + t
    return F.elu(linear(x.permute(0, 2, 1).reshape(B, 30), sd[p + "fc_final.0.weight"], sd[p + "fc_final.0.bias"], forward_mode="fp32"))


def privileged_encoder(sd, priv):
    pre = "privileged_encoder_.priv_encoder"
    return mlp(sd, pre, priv, n_layers(sd, pre))


def scan_encoder(sd, scan):
    pre = "scan_encoder.scan_encoder"
    return mlp(sd, pre, scan, n_layers(sd, pre))


def actor_mean(sd, obs, priv, est, scan, adaptation_mode=False):
    """ActorCritic.update_distribution (actor_critic.py:190-197)."""
    latent = adaptation_encoder(sd, obs) if adaptation_mode else privileged_encoder(sd, priv)
    x = torch.cat((obs, latent, scan_encoder(sd, scan), est), dim=-1)
    return mlp(sd, "actor", x, n_layers(sd, "actor"))


def critic_value(sd, critic_obs):
    return mlp(sd, "critic", critic_obs, n_layers(sd, "critic"))


def estimator(sd_est, obs):
    return mlp(sd_est, "estimator", obs, n_layers(sd_est, "estimator"))


def normal_log_prob(actions, mu, std):
    return (-((actions - mu) ** 2) / (2 * std ** 2) - torch.log(std) - math.log(math.sqrt(2 * math.pi))).sum(dim=-1)


def keyed_standard_normal(seed, step, num_envs, num_actions):
    """Box-Muller on keyed uniforms -- twin of sample_actions_kernel (csrc/learner_kernels.cu)."""
    lanes = np.arange(2 * num_actions)
    r = philox.keyed_u32(seed, philox.SITE_ACTION_NOISE, step, np.arange(num_envs), lanes)
    u1 = (((r[:, 0::2] >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)).astype(np.float32)
    u2 = ((r[:, 1::2] >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
    z = torch.sqrt(-2.0 * torch.log(torch.from_numpy(u1))) * torch.cos(torch.from_numpy(u2) * 6.283185307179586)
    return z


def ppo_act(sd, sd_est, obs, priv, critic_obs, scan, seed, step, adaptation_mode=False):
    """PPO.act (ppo.py:129-153): acts on the ESTIMATED obs; -> actions, values, log-prob, mu, sigma."""
    with torch.no_grad():
        est_hat = estimator(sd_est, obs)
        mu = actor_mean(sd, obs, priv, est_hat, scan, adaptation_mode)
        std = sd["std"]
        actions = mu + std * keyed_standard_normal(seed, step, obs.shape[0], mu.shape[1])
        values = critic_value(sd, critic_obs)
        return actions, values, normal_log_prob(actions, mu, mu * 0. + std), mu, (mu * 0. + std)


def ppo_losses(sd, sd_est, b, clip=0.2, value_coef=1.0, entropy_coef=0.01, reg_coef=0.0, use_clipped_value_loss=True):
    """The loss block of PPO.update for one minibatch `b` (dict of tensors), ppo.py:199-270.
    -> dict(loss, surrogate, value, reg, entropy, estimator)"""
    mu = actor_mean(sd, b["obs"], b["priv"], b["true_est"], b["scan"], False)
    sigma = mu * 0. + sd["std"]
    logp = normal_log_prob(b["actions"], mu, sigma)
    value = critic_value(sd, b["critic_obs"])
    entropy = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(sigma)).sum(dim=-1)
    lat_p = privileged_encoder(sd, b["priv"])
    with torch.no_grad():
        lat_a = adaptation_encoder(sd, b["obs"])
    reg = (lat_p - lat_a.detach()).norm(p=2, dim=1).mean()
    est_loss = (estimator(sd_est, b["obs"]) - b["true_est"]).norm(p=2, dim=1).pow(2).mean()
    ratio = torch.exp(logp - torch.squeeze(b["old_logp"]))
    adv = torch.squeeze(b["adv"])
    surrogate = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1.0 - clip, 1.0 + clip)).mean()
    if use_clipped_value_loss:
        vc = b["values"] + (value - b["values"]).clamp(-clip, clip)
        vloss = torch.max((value - b["returns"]).pow(2), (vc - b["returns"]).pow(2)).mean()
    else:
        vloss = (b["returns"] - value).pow(2).mean()
    loss = surrogate + value_coef * vloss - entropy_coef * entropy.mean() + reg_coef * reg
    return dict(loss=loss, surrogate=surrogate, value=vloss, reg=reg, entropy=entropy.mean(), estimator=est_loss,
                mu=mu, value_out=value, lat_p=lat_p, lat_a=lat_a)


def adaptive_lr(lr, mu, sigma, old_mu, old_sigma, desired_kl):
    """The schedule == 'adaptive' block of PPO.update (ppo.py:233-246) -> (new learning rate, kl_mean)."""
    with torch.no_grad():
        kl = torch.sum(torch.log(sigma / old_sigma + 1.e-5) + (torch.square(old_sigma) + torch.square(old_mu - mu)) / (2.0 * torch.square(sigma))
                       - 0.5, axis=-1)
        kl_mean = torch.mean(kl)
        if kl_mean > desired_kl * 2.0:
            lr = max(1e-5, lr / 1.5)
        elif kl_mean < desired_kl / 2.0 and kl_mean > 0.0:
            lr = min(1e-2, lr * 1.5)
    return lr, float(kl_mean)


def dagger_loss(sd, b):
    """PPO.update_dagger's loss (ppo.py:322-333)."""
    with torch.no_grad():
        lat_p = privileged_encoder(sd, b["priv"])
    lat_a = adaptation_encoder(sd, b["obs"])
    return (lat_p.detach() - lat_a).norm(p=2, dim=1).mean()


def clip_and_adam(params, grads, state, lr, max_norm, betas=(0.9, 0.999), eps=1e-8, extra_sumsq=None):
    """nn.utils.clip_grad_norm_ + torch.optim.Adam.step on lists of tensors; `state` = dict(step, m, v).
    `extra_sumsq`: squared norm of gradients the same clip_grad_norm_ call covers but this optimiser does not own
    (ppo.py:274: the adaptation encoder's stale .grad).  -> (post-clip squared norm of `grads`, rescaled extra_sumsq)"""
    own = sum((g.detach() ** 2).sum() for g in grads)
    total = torch.sqrt(own if extra_sumsq is None else own + extra_sumsq)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    state["total_norm"] = float(total)
    state["post_clip_sumsq"] = own * coef * coef
    state["extra_sumsq"] = None if extra_sumsq is None else extra_sumsq * coef * coef
    state["step"] += 1
    t = state["step"]
    bc1, bc2 = 1 - betas[0] ** t, 1 - betas[1] ** t
    for i, (p, g) in enumerate(zip(params, grads)):
        g = g.detach() * coef
        state["m"][i].lerp_(g, 1 - betas[0])
        state["v"][i].mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        denom = (state["v"][i].sqrt() / math.sqrt(bc2)).add_(eps)
        p.data.addcdiv_(state["m"][i], denom, value=-lr / bc1)


MAIN_PREFIXES = ("actor.", "critic.", "privileged_encoder_.", "std", "scan_encoder.")


class LearnerOracle:
    """PPO.update / update_dagger over explicit minibatch index lists (the reference draws them with randperm)."""

    def __init__(self, sd, sd_est, lr=2e-4, est_lr=1e-4, max_grad_norm=1.0, desired_kl=None, **loss_kw):
        self.sd = {k: v.clone().float().requires_grad_(True) for k, v in sd.items()}
        self.sd_est = {k: v.clone().float().requires_grad_(True) for k, v in sd_est.items()}
        self.lr, self.est_lr, self.max_grad_norm, self.loss_kw = lr, est_lr, max_grad_norm, loss_kw
        self.desired_kl, self.kl_log = desired_kl, []      # desired_kl set = schedule 'adaptive'
        self.main_keys = [k for k in self.sd if k.startswith(MAIN_PREFIXES)]
        self.adapt_keys = [k for k in self.sd if k.startswith("adaptation_encoder_.")]
        mk = lambda keys, d: dict(step=0, m=[torch.zeros_like(d[k]) for k in keys], v=[torch.zeros_like(d[k]) for k in keys])
        self.opt_main, self.opt_adapt = mk(self.main_keys, self.sd), mk(self.adapt_keys, self.sd)
        self.est_keys = list(self.sd_est)
        self.opt_est = mk(self.est_keys, self.sd_est)
        # squared norm of the adaptation encoder's stale .grad (left post-clip by the last update_dagger minibatch; the
        # main optimiser's zero_grad() never clears it, and clip_grad_norm_(actor_critic.parameters()) covers + rescales it)
        self.stale_adapt_sumsq = None

    def minibatch(self, b, reg_coef=0.0):
        out = ppo_losses(self.sd, self.sd_est, b, reg_coef=reg_coef, **self.loss_kw)
        g_est = torch.autograd.grad(out["estimator"], [self.sd_est[k] for k in self.est_keys])
        clip_and_adam([self.sd_est[k] for k in self.est_keys], g_est, self.opt_est, self.est_lr, self.max_grad_norm)
        g = torch.autograd.grad(out["loss"], [self.sd[k] for k in self.main_keys], allow_unused=True)
        g = [torch.zeros_like(self.sd[k]) if gi is None else gi for k, gi in zip(self.main_keys, g)]
        self.last_grads = dict(zip(self.main_keys, g), **dict(zip(self.est_keys, g_est)))
        if self.desired_kl is not None:
            mu = out["mu"].detach()
            self.lr, kl = adaptive_lr(self.lr, mu, mu * 0. + self.sd["std"].detach(), b["mu"], b["sigma"], self.desired_kl)
            self.kl_log.append((kl, self.lr))
        clip_and_adam([self.sd[k] for k in self.main_keys], g, self.opt_main, self.lr, self.max_grad_norm,
                      extra_sumsq=self.stale_adapt_sumsq)
        self.stale_adapt_sumsq = self.opt_main["extra_sumsq"]
        return {k: float(v.detach()) for k, v in out.items() if v.dim() == 0}

    def dagger_minibatch(self, b):
        loss = dagger_loss(self.sd, b)
        g = torch.autograd.grad(loss, [self.sd[k] for k in self.adapt_keys])
        self.last_grads = dict(zip(self.adapt_keys, g))
        clip_and_adam([self.sd[k] for k in self.adapt_keys], g, self.opt_adapt, self.lr, self.max_grad_norm)
        self.stale_adapt_sumsq = self.opt_adapt["post_clip_sumsq"]
        return float(loss.detach())
