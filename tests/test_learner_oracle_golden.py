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
