"""oracle/learner_oracle.py against tests/golden/learner_small.npz (produced by the reference's own rsl_rl
PPO / ActorCritic / RolloutStorage, oracle/make_golden_learner.py).  Runs anywhere (CPU)."""
import ast
import os

import numpy as np
import torch

import learner_util as lu
from oracle import learner_oracle as lo

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "learner_small.npz"))
T, N, EPOCHS, MBS = 6, 32, 2, 2


def _sd(prefix):
    return {k[len(prefix):]: torch.from_numpy(G[k]) for k in G.files if k.startswith(prefix)}


def _storage():
    return {k[len("storage/"):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("storage/")}


def test_gae_golden():
    t = lambda k: torch.from_numpy(G["gae/" + k])
    ret, adv = lo.compute_returns(t("rewards"), t("dones"), t("values"), t("last_values"), 0.99, 0.95)
    assert torch.equal(ret, t("returns")) and torch.equal(adv, t("advantages"))


def test_act_statistics_golden():
    sd, sd_est, st = _sd("init/ac/"), _sd("init/est/"), _storage()
    b = lu.minibatch(st, torch.arange(N))
    with torch.no_grad():
        est = lo.estimator(sd_est, b["obs"])
        assert torch.equal(est, torch.from_numpy(G["act/est"]))
        for mode in (False, True):
            mu = lo.actor_mean(sd, b["obs"], b["priv"], est, b["scan"], mode)
            assert torch.equal(mu, torch.from_numpy(G[f"act/mu_{int(mode)}"]))
            assert torch.equal(lo.normal_log_prob(b["actions"], mu, mu * 0. + sd["std"]), torch.from_numpy(G[f"act/logp_{int(mode)}"]))
        assert torch.equal(lo.critic_value(sd, b["critic_obs"]), torch.from_numpy(G["act/value"]))


def test_update_golden():
    sd, sd_est, st = _sd("init/ac/"), _sd("init/est/"), _storage()
    perm = torch.from_numpy(G["perm"])
    orc = lo.LearnerOracle(sd, sd_est, lr=2e-4, est_lr=1e-4)
    v, sur, reg, coef, el = G["update/returned"]
    assert coef == 0.1
    mb, logs = T * N // MBS, []
    for _ in range(EPOCHS):
        for i in range(MBS):
            logs.append(orc.minibatch(lu.minibatch(st, perm[i * mb:(i + 1) * mb]), reg_coef=float(coef)))
    mean = lambda k: sum(l[k] for l in logs) / len(logs)
    for mine, ref in ((mean("value"), v), (mean("surrogate"), sur), (mean("reg"), reg), (mean("estimator"), el)):
        assert abs(mine - ref) <= 1e-6 * abs(ref)
    after, after_est = _sd("update/ac/"), _sd("update/est/")
    for k in orc.main_keys:
        mine = orc.sd[k].detach() if k != "std" else torch.min(orc.sd[k].detach(), torch.tensor(1.0))
        assert torch.allclose(mine, after[k], rtol=1e-6, atol=5e-8), k
    for k in orc.adapt_keys:
        assert torch.equal(orc.sd[k].detach(), after[k]), k          # PPO.update never touches the adaptation encoder
    for k in orc.est_keys:
        assert torch.allclose(orc.sd_est[k].detach(), after_est[k], rtol=1e-6, atol=5e-8), k


def test_dagger_golden():
    sd, sd_est, st = _sd("init/ac/"), _sd("init/est/"), _storage()
    perm = torch.from_numpy(G["perm"])
    orc = lo.LearnerOracle(sd, sd_est, lr=2e-4)
    mb, losses = T * N // MBS, []
    for _ in range(EPOCHS):
        for i in range(MBS):
            losses.append(orc.dagger_minibatch(lu.minibatch(st, perm[i * mb:(i + 1) * mb])))
    ref = float(G["dagger/returned"][0])
    assert abs(sum(losses) / len(losses) - ref) <= 1e-6 * abs(ref)
    after = _sd("dagger/ac/")
    for k in orc.adapt_keys:
        assert torch.allclose(orc.sd[k].detach(), after[k], rtol=1e-6, atol=5e-8), k
    for k in orc.main_keys:
        assert torch.equal(orc.sd[k].detach(), after[k]), k


def test_dagger_then_update_golden():
    """update_dagger() followed by update() on the SAME PPO object, as OnPolicyRunner.learn runs them (iteration 0 is a DAgger
    iteration): the adaptation encoder's stale post-clip .grad is part of every later clip_grad_norm_ (ppo.py:274)."""
    sd, sd_est = _sd("init/ac/"), _sd("init/est/")
    st = {k[len("seq_storage/"):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("seq_storage/")}
    perm = torch.from_numpy(G["perm"])
    orc = lo.LearnerOracle(sd, sd_est, lr=2e-4, est_lr=1e-4)
    mb = T * N // MBS
    dl = [orc.dagger_minibatch(lu.minibatch(st, perm[i * mb:(i + 1) * mb])) for _ in range(EPOCHS) for i in range(MBS)]
    assert abs(sum(dl) / len(dl) - float(G["seq/dagger_returned"][0])) <= 1e-6 * abs(float(G["seq/dagger_returned"][0]))
    assert orc.stale_adapt_sumsq is not None and float(orc.stale_adapt_sumsq) > 0.0
    v, sur, reg, coef, el = G["seq/update_returned"]
    logs, norms = [], []
    for _ in range(EPOCHS):
        for i in range(MBS):
            logs.append(orc.minibatch(lu.minibatch(st, perm[i * mb:(i + 1) * mb]), reg_coef=float(coef)))
            norms += [orc.opt_est["total_norm"], orc.opt_main["total_norm"]]
    # the norms the reference's clip_grad_norm_ calls returned (estimator, main alternating; main includes the stale term)
    assert np.allclose(norms, G["seq/clip_total_norms"], rtol=2e-5), (norms, G["seq/clip_total_norms"])
    mean = lambda k: sum(l[k] for l in logs) / len(logs)
    for mine, ref in ((mean("value"), v), (mean("surrogate"), sur), (mean("reg"), reg), (mean("estimator"), el)):
        assert abs(mine - ref) <= 2e-6 * abs(ref)
    after, after_est = _sd("seq/ac/"), _sd("seq/est/")
    for k in orc.main_keys + orc.adapt_keys:
        mine = orc.sd[k].detach() if k != "std" else torch.min(orc.sd[k].detach(), torch.tensor(1.0))
        assert torch.allclose(mine, after[k], rtol=1e-6, atol=5e-8), k
    for k in orc.est_keys:
        assert torch.allclose(orc.sd_est[k].detach(), after_est[k], rtol=1e-6, atol=5e-8), k
    # the quirk is visible in the fixture: without the stale norm the same sequence lands elsewhere
    plain = lo.LearnerOracle(sd, sd_est, lr=2e-4, est_lr=1e-4)
    for _ in range(EPOCHS):
        for i in range(MBS):
            plain.dagger_minibatch(lu.minibatch(st, perm[i * mb:(i + 1) * mb]))
    plain.stale_adapt_sumsq = None
    for _ in range(EPOCHS):
        for i in range(MBS):
            plain.minibatch(lu.minibatch(st, perm[i * mb:(i + 1) * mb]), reg_coef=float(coef))
    assert any(not torch.allclose(plain.sd[k].detach(), after[k], rtol=1e-6, atol=5e-8) for k in orc.main_keys if k != "std")


def test_tf32_numerics_mode():
    """numerics('tf32') (the comparator of the production tcgen05 path): operand rounding only -- results stay within TF32
    distance of the fp32 oracle, the custom backward equals its own definition, and fp32 mode is untouched."""
    x = torch.randn(4096, generator=torch.Generator().manual_seed(0)) * 3
    t, r = lo.tf32_trunc(x), lo.tf32_rna(x)
    assert (t.view(torch.int32) & 8191).eq(0).all() and (r.view(torch.int32) & 8191).eq(0).all()
    assert (t.abs() <= x.abs()).all() and ((t - x).abs() / x.abs()).max() < 2.0 ** -10
    assert ((r - x).abs() / x.abs()).max() <= 2.0 ** -11 and float((t != r).float().mean()) > 0.3
    assert lo.production_gemm_modes(24576, 512, 627) == ("trunc", "trunc", "trunc")
    assert lo.production_gemm_modes(24576, 1, 128) == ("fp32", "fp32", "rna") and lo.production_gemm_modes(16, 12, 128)[2] == "rna"
    lo._CHAIN[0] = True
    assert lo.production_gemm_modes(4096, 3, 128) == ("trunc", "fp32", "rna")        # optional one-launch forward chains
    lo._CHAIN[0] = False
    assert lo.production_gemm_modes(4096, 3, 128) == ("fp32", "fp32", "rna") and lo.production_gemm_modes(24576, 6, 128)[0] == "rna"
    sd, sd_est, st = _sd("init/ac/"), _sd("init/est/"), _storage()
    b = lu.minibatch(st, torch.arange(T * N))
    outs = {}
    for mode in ("fp32", "tf32"):
        o = lo.LearnerOracle(sd, sd_est)
        with lo.numerics(mode):
            logs = o.minibatch(b, reg_coef=0.1)
            o.dagger_minibatch(b)
        outs[mode] = (logs, dict(o.last_grads), o)
    for k in ("surrogate", "value", "reg", "estimator"):
        a, c = outs["fp32"][0][k], outs["tf32"][0][k]
        assert a != c and abs(a - c) <= 5e-3 * abs(a), k
    for k, g in outs["fp32"][1].items():
        h = outs["tf32"][1][k]
        assert float((g - h).abs().max() / g.pow(2).mean().sqrt().clamp_min(1e-12)) <= 3e-2, k
    # definition of the backward: dX = r(dY) r(W), dW = r(dY)^T r(X)
    g = torch.Generator().manual_seed(1)
    X, W, dY = torch.randn(64, 40, generator=g, requires_grad=True), torch.randn(16, 40, generator=g, requires_grad=True), torch.randn(64, 16, generator=g)
    with lo.numerics("tf32"):
        y = lo.linear(X, W, torch.zeros(16))
    y.backward(dY)
    tr = lo.tf32_trunc
    assert torch.equal(y.detach(), tr(X.detach()) @ tr(W.detach()).t() + 0)
    assert torch.equal(X.grad, tr(dY) @ tr(W.detach())) and torch.equal(W.grad, tr(dY).t() @ tr(X.detach()))
