"""Learner kernels + host classes (through the C ABI) vs the CPU oracle / the reference golden.

Tolerances, stated per check:
  * GEMMs in 3xTF32 ('precise') mode are fp32-accurate: <= 5e-5 relative to the output scale vs torch fp32;
    in TF32 mode (the reference's own GPU matmul precision, train.py:39) <= 3e-3 (mma.sync, rounded) / 1e-2 (tcgen05, truncated);
  * network outputs / losses in precise mode: <= 1e-4 relative;
  * gradients: <= 1e-3 relative to each tensor's rms (fp32 atomics in split-K reorder the sums);
  * post-Adam parameters: 99.9% of the elements within 2% of the largest possible movement (lr * steps) and
    none beyond 2.2x that movement -- Adam's g/sqrt(v) normalisation turns rounding noise of a near-zero
    gradient into a +-lr step, so a per-element bound tighter than the movement itself is not meaningful.
"""
import ast
import ctypes as C
import os

import numpy as np
import pytest
import torch

import golden_util as gu
import learner_util as lu
from legged_gym_custom_b200 import _lib
from oracle import learner_oracle as lo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = np.load(os.path.join(gu.GOLDEN_DIR, "learner_small.npz"))
HID = ast.literal_eval(str(GOLD["hid"]))


def scale_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.pow(2).mean().sqrt().clamp_min(1e-12))


def fro_err(a, b):
    """relative Frobenius distance ||a - b|| / ||b||"""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt().clamp_min(1e-30))


@pytest.mark.parametrize("M,N,K", [(4096, 512, 627), (333, 12, 128), (1000, 1, 128), (24576, 256, 512), (129, 30, 52), (64, 20, 36),
                                   (500, 3, 128), (2048, 64, 29)])
@pytest.mark.parametrize("precise", [1, 0])
def test_linear_kernels(M, N, K, precise):
    lib = _lib.lib()
    g = torch.Generator().manual_seed(M + N + K)
    ld = lambda k: (k + 3) // 4 * 4
    X = torch.zeros(M, ld(K)); X[:, :K] = torch.randn(M, K, generator=g)
    W = torch.zeros(N, ld(K)); W[:, :K] = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    dY = torch.zeros(M, ld(N)); dY[:, :N] = torch.randn(M, N, generator=g)
    Xd, Wd, bd, dYd = X.to(DEV), W.to(DEV), b.to(DEV), dY.to(DEV)
    p = lambda t: t.data_ptr()
    st = _lib.stream_ptr()
    tol = 5e-5 if precise else 3e-3
    # forward with ELU
    Y = torch.zeros(M, ld(N), device=DEV)
    _lib.check(lib.b200_linear_forward(p(Xd), ld(K), p(Wd), ld(K), p(bd), p(Y), ld(N), M, N, K, 1, precise, st))
    Yref = torch.nn.functional.elu(X[:, :K] @ W[:, :K].t() + b)
    assert scale_err(Y[:, :N], Yref) <= tol
    assert float(Y[:, N:].abs().max()) == 0.0 if ld(N) > N else True
    # dgrad with elu'(Yprev) and accumulate
    Yprev = torch.randn(M, ld(K), generator=g)
    dX = torch.ones(M, ld(K), device=DEV)
    _lib.check(lib.b200_linear_dgrad(p(dYd), ld(N), p(Wd), ld(K), p(Yprev.to(DEV)), ld(K), p(dX), ld(K), M, N, K, 1, precise, st))
    ref = 1.0 + (dY[:, :N] @ W[:, :K]) * torch.where(Yprev[:, :K] > 0, torch.ones(()), Yprev[:, :K] + 1.0)
    assert scale_err(dX[:, :K], ref) <= tol
    # wgrad accumulates into dW / db
    dW, db = torch.full((N, ld(K)), 0.5, device=DEV), torch.full((N,), 0.25, device=DEV)
    _lib.check(lib.b200_linear_wgrad(p(dYd), ld(N), p(Xd), ld(K), p(dW), ld(K), p(db), M, N, K, precise, st))
    torch.cuda.synchronize()
    assert scale_err(dW[:, :K], 0.5 + dY[:, :N].t() @ X[:, :K]) <= tol
    assert scale_err(db, 0.25 + dY[:, :N].sum(0)) <= 1e-5


@pytest.mark.parametrize("M,N,K", [(4096, 512, 627), (333, 12, 128), (24576, 256, 512), (129, 30, 52), (64, 20, 36), (500, 32, 128),
                                   (2048, 64, 29), (1000, 128, 132), (4096, 512, 736)])
def test_tcgen05_linear_kernels(M, N, K):
    """the tcgen05 / TMEM / TMA GEMMs (kind::tf32) vs torch fp32: forward (bias + ELU), dgrad (elu' + accumulate),
    wgrad (split-K accumulation) + colsum bias gradient.  TMA zero-fill covers the K / N / M tails."""
    lib = _lib.lib()
    g = torch.Generator().manual_seed(M + N + K)
    ld = lambda k: (k + 3) // 4 * 4
    X = torch.zeros(M, ld(K)); X[:, :K] = torch.randn(M, K, generator=g)
    W = torch.zeros(N, ld(K)); W[:, :K] = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    dY = torch.zeros(M, ld(N)); dY[:, :N] = torch.randn(M, N, generator=g)
    Xd, Wd, bd, dYd = X.to(DEV), W.to(DEV), b.to(DEV), dY.to(DEV)
    p = lambda t: t.data_ptr()
    st = _lib.stream_ptr()
    tol = 1e-2      # kind::tf32 consumes the raw fp32 bits (10-bit mantissa by truncation, as cuBLAS TF32 does)
    Y = torch.zeros(M, ld(N), device=DEV)
    _lib.check(lib.b200_tc_linear_forward(p(Xd), ld(K), p(Wd), ld(K), p(bd), p(Y), ld(N), M, N, K, 1, st))
    assert scale_err(Y[:, :N], torch.nn.functional.elu(X[:, :K] @ W[:, :K].t() + b)) <= tol
    if ld(N) > N:
        assert float(Y[:, N:].abs().max()) == 0.0
    Yprev = torch.randn(M, ld(K), generator=g)
    dX = torch.ones(M, ld(K), device=DEV)
    _lib.check(lib.b200_tc_linear_dgrad(p(dYd), ld(N), p(Wd), ld(K), p(Yprev.to(DEV)), ld(K), p(dX), ld(K), M, N, K, 1, st))
    ref = 1.0 + (dY[:, :N] @ W[:, :K]) * torch.where(Yprev[:, :K] > 0, torch.ones(()), Yprev[:, :K] + 1.0)
    assert scale_err(dX[:, :K], ref) <= tol
    dW, db = torch.full((N, ld(K)), 0.5, device=DEV), torch.full((N,), 0.25, device=DEV)
    _lib.check(lib.b200_tc_linear_wgrad(p(dYd), ld(N), p(Xd), ld(K), p(dW), ld(K), M, N, K, st))
    _lib.check(lib.b200_colsum(p(dYd), ld(N), p(db), M, N, st))
    torch.cuda.synchronize()
    assert scale_err(dW[:, :K], 0.5 + dY[:, :N].t() @ X[:, :K]) <= tol
    if ld(K) > K:
        assert float((dW[:, K:] - 0.5).abs().max()) == 0.0
    assert scale_err(db, 0.25 + dY[:, :N].sum(0)) <= 1e-5


@pytest.mark.parametrize("M,N,K", [(24576, 256, 512), (24576, 128, 256), (1000, 64, 128), (333, 512, 52), (129, 30, 52)])
def test_tcgen05_dgrad_fused_bias_gradient_and_partial_accumulate(M, N, K):
    """b200_tc_linear_dgrad_bias: dbias_prev += column sums of dX (fp32 atomics: tight tolerance against the column sums of the
    kernel's own dX); dgrad `accumulate` = n > 1 adds the old contents of the first n output columns only."""
    lib = _lib.lib()
    g = torch.Generator().manual_seed(7 * M + N + K)
    ld = lambda k: (k + 3) // 4 * 4
    W = torch.zeros(N, ld(K)); W[:, :K] = torch.randn(N, K, generator=g) / K ** 0.5
    dY = torch.zeros(M, ld(N)); dY[:, :N] = torch.randn(M, N, generator=g)
    Yprev = torch.randn(M, ld(K), generator=g)
    Wd, dYd, Ypd = W.to(DEV), dY.to(DEV), Yprev.to(DEV)
    p = lambda t: t.data_ptr()
    st = _lib.stream_ptr()
    dX, db = torch.zeros(M, ld(K), device=DEV), torch.full((ld(K),), 0.5, device=DEV)
    _lib.check(lib.b200_tc_linear_dgrad_bias(p(dYd), ld(N), p(Wd), ld(K), p(Ypd), ld(K), p(dX), ld(K), M, N, K, 0, p(db), st))
    torch.cuda.synchronize()
    ref = (dY[:, :N] @ W[:, :K]) * torch.where(Yprev[:, :K] > 0, torch.ones(()), Yprev[:, :K] + 1.0)
    assert scale_err(dX[:, :K], ref) <= 1e-2
    assert scale_err(db[:K] - 0.5, dX[:, :K].double().sum(0).float()) <= 2e-5
    if ld(K) > K:
        assert float((db[K:] - 0.5).abs().max()) == 0.0
    # partial-column accumulate: the first n columns keep their old contents added, the others are overwritten
    n = max(2, K // 3)
    old = torch.randn(M, ld(K), generator=g)
    dX2 = old.to(DEV).clone()
    _lib.check(lib.b200_tc_linear_dgrad(p(dYd), ld(N), p(Wd), ld(K), p(Ypd), ld(K), p(dX2), ld(K), M, N, K, n, st))
    torch.cuda.synchronize()
    exp = dX[:, :K].cpu().clone()
    exp[:, :n] += old[:, :n]
    assert scale_err(dX2[:, :K], exp) <= 1e-5
    # the mma.sync parity path implements the same `accumulate` contract
    dX3 = old.to(DEV).clone()
    _lib.check(lib.b200_linear_dgrad(p(dYd), ld(N), p(Wd), ld(K), p(Ypd), ld(K), p(dX3), ld(K), M, N, K, n, 1, st))
    torch.cuda.synchronize()
    ref3 = ref.clone()
    ref3[:, :n] += old[:, :n]
    assert scale_err(dX3[:, :K], ref3) <= 5e-5


def test_tcgen05_pair_pdl_switches_and_launch_trace():
    """CTA pairs (cta_group::2) and PDL change how a GEMM is scheduled, not what it computes: bit-identical outputs.  The
    launch trace records one {start, end} slot and one {M, N, K, code} row per launch."""
    lib = _lib.lib()
    M, N, K = 24576, 512, 627
    g = torch.Generator().manual_seed(5)
    ld = lambda k: (k + 3) // 4 * 4
    X = torch.zeros(M, ld(K)); X[:, :K] = torch.randn(M, K, generator=g)
    W = torch.zeros(N, ld(K)); W[:, :K] = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    Xd, Wd, bd = X.to(DEV), W.to(DEV), b.to(DEV)
    p = lambda t: t.data_ptr()
    st = _lib.stream_ptr()
    outs = []
    try:
        for pair, pdl in ((1, 0), (0, 0), (1, 1), (0, 1)):
            lib.b200_tc_set_pair_mode(pair)
            lib.b200_tc_set_pdl(pdl)
            Y = torch.zeros(M, ld(N), device=DEV)
            _lib.check(lib.b200_tc_linear_forward(p(Xd), ld(K), p(Wd), ld(K), p(bd), p(Y), ld(N), M, N, K, 1, st))
            _lib.check(lib.b200_tc_linear_forward(p(Xd), ld(K), p(Wd), ld(K), p(bd), p(Y), ld(N), M, N, K, 1, st))   # back to back (PDL)
            torch.cuda.synchronize()
            outs.append(Y)
    finally:
        lib.b200_tc_set_pair_mode(1)
        lib.b200_tc_set_pdl(0)
    for Y in outs[1:]:
        assert torch.equal(Y, outs[0])
    # one / two persistent CTAs per SM (b200_tc_set_ctas_per_sm) and a per-stream SM cap (b200_tc_set_stream_sm_cap) change the
    # schedule of the tiles, not what a tile computes
    M2, N2, K2 = 24576, 256, 512
    X2, W2 = torch.randn(M2, K2, generator=g).to(DEV), (torch.randn(N2, K2, generator=g) / K2 ** 0.5).to(DEV)
    dY2, Yp2 = torch.randn(M2, N2, generator=g).to(DEV), torch.randn(M2, K2, generator=g).to(DEV)
    side = torch.cuda.Stream()
    res = []
    try:
        for cps, cap, epi in ((1, 0, 1), (2, 0, 1), (0, 0, 1), (0, 100, 1), (0, 0, 0), (2, 0, 0)):
            lib.b200_tc_set_ctas_per_sm(cps)
            lib.b200_tc_set_tma_epilogue(epi)
            _lib.check(lib.b200_tc_set_stream_sm_cap(C.c_void_p(side.cuda_stream), cap))
            Y2, dX2 = torch.zeros(M2, N2, device=DEV), torch.zeros(M2, K2, device=DEV)
            torch.cuda.synchronize()
            with torch.cuda.stream(side):
                sp = _lib.stream_ptr()
                _lib.check(lib.b200_tc_linear_forward(p(X2), K2, p(W2), K2, p(bd), p(Y2), N2, M2, N2, K2, 1, sp))
                _lib.check(lib.b200_tc_linear_dgrad(p(dY2), N2, p(W2), K2, p(Yp2), K2, p(dX2), K2, M2, N2, K2, 0, sp))
            torch.cuda.synchronize()
            res.append((Y2, dX2))
    finally:
        lib.b200_tc_set_ctas_per_sm(0)
        lib.b200_tc_set_tma_epilogue(1)
        lib.b200_tc_set_stream_sm_cap(C.c_void_p(side.cuda_stream), 0)
    for Y2, dX2 in res[1:]:
        assert torch.equal(Y2, res[0][0]) and torch.equal(dX2, res[0][1])
    dev = torch.zeros(4, 2, dtype=torch.int64, device=DEV)
    dev[:, 0] = torch.iinfo(torch.int64).max
    meta = np.zeros((4, 4), dtype=np.int64)
    lib.b200_tc_set_trace(C.c_void_p(dev.data_ptr()), meta.ctypes.data_as(C.c_void_p), 4)
    try:
        for _ in range(2):
            _lib.check(lib.b200_tc_linear_forward(p(Xd), ld(K), p(Wd), ld(K), p(bd), p(outs[0]), ld(N), M, N, K, 1, st))
        torch.cuda.synchronize()
    finally:
        lib.b200_tc_set_trace(None, None, 0)
    t = dev.cpu().numpy()
    assert (meta[:2] == np.array([M, N, K, 0 * 1000 + 256 + 500])).all() and (meta[2:] == 0).all()
    assert (t[:2, 1] > t[:2, 0]).all() and t[1, 0] >= t[0, 1] - 2000 and (t[2:, 1] == 0).all()
    assert 5e3 < t[0, 1] - t[0, 0] < 5e5                       # tens of microseconds


def assert_params_close(mine, ref, move, name, max_frac=1e-3):
    d = (mine.cpu() - ref).abs()
    frac = float((d > 0.02 * move).float().mean())
    assert frac <= max_frac and float(d.max()) <= 2.2 * move, (name, frac, float(d.max()), move)


def _build(hid, precise=True):
    from legged_gym_custom_b200.networks import ActorCritic, MlpEstimator
    ac = ActorCritic(52, 29, 736, 3, 132, 12, 10, actor_hidden_dims=hid["actor"], critic_hidden_dims=hid["critic"],
                     priv_encoder_hidden_dims=hid["priv"], scan_encoder_hidden_dims=hid["scan"], latent_encoder_output_dim=20,
                     scan_encoder_output_dim=32, activation='elu', init_noise_std=0.8, device=DEV, precise=precise)
    est = MlpEstimator(52, 10, 3, hidden_dims=hid["est"], activation='elu', use_history=True, device=DEV, precise=precise)
    return ac, est


def _gold_sd(prefix):
    return {k[len(prefix):]: torch.from_numpy(GOLD[k]) for k in GOLD.files if k.startswith(prefix)}


def test_state_dict_roundtrip_and_forward_vs_reference_golden():
    ac, est = _build(HID)
    sd, sd_est = _gold_sd("init/ac/"), _gold_sd("init/est/")
    ac.load_state_dict(sd)
    est.load_state_dict(sd_est)
    back = ac.state_dict()
    assert list(back.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(back[k].cpu(), sd[k]), k
    st = {k[len("storage/"):]: torch.from_numpy(GOLD[k]) for k in GOLD.files if k.startswith("storage/")}
    b = {k: v.to(DEV) for k, v in lu.minibatch(st, torch.arange(32)).items()}
    e = est(b["obs"])
    assert scale_err(e, torch.from_numpy(GOLD["act/est"])) <= 1e-4
    for mode in (False, True):
        mu = ac.act_inference(b["obs"], b["priv"], e, b["scan"], adaptation_mode=mode)
        assert scale_err(mu, torch.from_numpy(GOLD[f"act/mu_{int(mode)}"])) <= 1e-4, mode
        ac.update_distribution(b["obs"], b["priv"], e, b["scan"], adaptation_mode=mode)
        assert scale_err(ac.get_actions_log_prob(b["actions"]), torch.from_numpy(GOLD[f"act/logp_{int(mode)}"])) <= 1e-4
    assert scale_err(ac.evaluate(b["critic_obs"]), torch.from_numpy(GOLD["act/value"])) <= 1e-4


def _ppo(ac, est, N, T, epochs=2, mbs=2, resume=True, schedule='fixed'):
    from legged_gym_custom_b200.learner import PPO
    ppo = PPO(ac, est, num_learning_epochs=epochs, num_mini_batches=mbs, clip_param=0.2, gamma=0.99, lam=0.95, value_loss_coef=1.0,
              entropy_coef=0.01, learning_rate=2e-4, estimator_learning_rate=1e-4, max_grad_norm=1.0, use_clipped_value_loss=True,
              schedule=schedule, desired_kl=0.01, resume=resume, device=DEV, seed=3)
    ppo.init_storage(N, T, [572], [29], [736], [3], [132], [12])
    return ppo


def _fill(ppo, st):
    s = ppo.storage
    d = lambda t: t.to(DEV)
    s.observations.copy_(d(st["obs"])); s.privileged_observations.copy_(d(st["priv"])); s.critic_observations.copy_(d(st["critic_obs"]))
    s.true_estimated_observations.copy_(d(st["true_est"])); s.scan_observations.copy_(d(st["scan"])); s.actions.copy_(d(st["actions"]))
    s.values.copy_(d(st["values"])); s.returns.copy_(d(st["returns"])); s.advantages.copy_(d(st["adv"]))
    s.actions_log_prob.copy_(d(st["old_logp"])); s.mu.copy_(d(st["mu"])); s.sigma.copy_(d(st["sigma"]))


def _tf32_oracle_update(sd, sd_est, st, perm, T, N, epochs, mbs, reg_coef, dagger_first=False):
    """the same update on the oracle in numerics('tf32'): the comparator of the production (precise=False) path"""
    orc = lo.LearnerOracle(sd, sd_est, lr=2e-4, est_lr=1e-4)
    mb, logs, dl = T * N // mbs, [], []
    with lo.numerics("tf32"):
        if dagger_first:
            dl = [orc.dagger_minibatch(lu.minibatch(st, perm[i * mb:(i + 1) * mb])) for _ in range(epochs) for i in range(mbs)]
        for _ in range(epochs):
            for i in range(mbs):
                logs.append(orc.minibatch(lu.minibatch(st, perm[i * mb:(i + 1) * mb]), reg_coef=reg_coef))
    mean = lambda k: sum(l[k] for l in logs) / len(logs)
    return orc, {k: mean(k) for k in ("value", "surrogate", "reg", "estimator")}, (sum(dl) / len(dl) if dl else None)


def test_tcgen05_operands_are_truncated_tf32():
    """pins the assumption behind numerics('tf32'): kind::tf32 reads the fp32 bits and ignores the low 13 mantissa bits.
    One forward GEMM against fp64 products of TRUNCATED operands (fp32-accumulation distance) and of ROUNDED operands (far)."""
    lib = _lib.lib()
    M, N, K = 2048, 256, 512
    g = torch.Generator().manual_seed(9)
    X, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    Xd, Wd, Y = X.to(DEV), W.to(DEV), torch.zeros(M, N, device=DEV)
    _lib.check(lib.b200_tc_linear_forward(Xd.data_ptr(), K, Wd.data_ptr(), K, None, Y.data_ptr(), N, M, N, K, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref_t = (lo.tf32_trunc(X).double() @ lo.tf32_trunc(W).double().t()).float()
    ref_r = (lo.tf32_rna(X).double() @ lo.tf32_rna(W).double().t()).float()
    e_t, e_r = scale_err(Y, ref_t), scale_err(Y, ref_r)
    assert e_t <= 2e-5 and e_r >= 10 * e_t, (e_t, e_r)


def test_mma_sync_heads_round_to_nearest():
    """the other half of numerics('tf32'): layers 5..7 wide run on mma.sync with cvt.rna operands; heads with N <= 4 are
    streaming fp32 dot products (small_n_forward_kernel), exact to fp32 summation order"""
    lib = _lib.lib()
    for (M, N, K) in ((512, 6, 128), (512, 7, 64)):
        g = torch.Generator().manual_seed(3)
        X, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
        Xd, Wd, Y = X.to(DEV), W.to(DEV), torch.zeros(M, 8, device=DEV)
        _lib.check(lib.b200_linear_forward(Xd.data_ptr(), K, Wd.data_ptr(), K, None, Y.data_ptr(), 8, M, N, K, 0, 0, _lib.stream_ptr()))
        torch.cuda.synchronize()
        e_r = scale_err(Y[:, :N], (lo.tf32_rna(X).double() @ lo.tf32_rna(W).double().t()).float())
        e_t = scale_err(Y[:, :N], (lo.tf32_trunc(X).double() @ lo.tf32_trunc(W).double().t()).float())
        assert e_r <= 2e-5 and e_t >= 10 * e_r, (e_r, e_t)
    for (M, N, K, ldx, act) in ((512, 1, 128, 128, 0), (4096, 3, 128, 132, 0), (333, 4, 70, 72, 1), (1, 2, 9, 12, 0)):
        g = torch.Generator().manual_seed(5)
        ldw = (K + 3) // 4 * 4
        X, Wp, b = torch.randn(M, ldx, generator=g), torch.randn(N, ldw, generator=g) / K ** 0.5, torch.randn(N, generator=g)
        W = Wp[:, :K]
        Xd, Wd, bd, Y = X.to(DEV), Wp.to(DEV), b.to(DEV), torch.full((M, 4), 7.0, device=DEV)
        _lib.check(lib.b200_linear_forward(Xd.data_ptr(), ldx, Wd.data_ptr(), ldw, bd.data_ptr(), Y.data_ptr(), 4, M, N, K, act, 0,
                                           _lib.stream_ptr()))
        torch.cuda.synchronize()
        ref = X[:, :K].double() @ W.double().t() + b.double()
        if act:
            ref = torch.where(ref > 0, ref, torch.expm1(ref))
        assert scale_err(Y[:, :N], ref.float()) <= 2e-6, (M, N, K)
        assert (Y[:, N:] == 7.0).all()                      # columns beyond N are not touched
        e_r = scale_err(Y[:, :N], (lo.tf32_rna(X[:, :K]).double() @ lo.tf32_rna(W).double().t() + b.double()).float())
        assert act or e_r >= 1e-5


@pytest.mark.parametrize("dagger_first", [False, True], ids=["update", "dagger-then-update"])
def test_production_path_update_on_reference_golden_storage(dagger_first):
    """PPO.update (and update_dagger -> update on the same object, the runner's order) through the PRODUCTION tcgen05 TF32
    kernels on the reference's golden storage / weights / permutation: against the TF32 oracle at 1e-4 (losses) and the
    Adam-step statistic, and against the reference's own fp32 result at TF32 distance (5e-3)."""
    T, N = 6, 32
    ac, est = _build(HID, precise=False)
    sd, sd_est = _gold_sd("init/ac/"), _gold_sd("init/est/")
    ac.load_state_dict(sd); est.load_state_dict(sd_est)
    ppo = _ppo(ac, est, N, T)
    pre = "seq_storage/" if dagger_first else "storage/"
    st = {k[len(pre):]: torch.from_numpy(GOLD[k]) for k in GOLD.files if k.startswith(pre)}
    _fill(ppo, st)
    perm = torch.from_numpy(GOLD["perm"])
    if dagger_first:
        dloss = ppo.update_dagger_with_indices(perm.to(DEV))
    else:
        ppo.total_updates = 2.0
    v, sur, reg, coef, el = ppo.update_with_indices(perm.to(DEV))
    ref = GOLD["seq/update_returned" if dagger_first else "update/returned"]
    assert coef == ref[3]
    orc, means, dmean = _tf32_oracle_update(sd, sd_est, st, perm, T, N, 2, 2, float(coef), dagger_first)
    for mine, key, r in ((v, "value", ref[0]), (sur, "surrogate", ref[1]), (reg, "reg", ref[2]), (el, "estimator", ref[4])):
        assert abs(mine - means[key]) <= 1e-4 * abs(means[key]), (key, mine, means[key])
        assert abs(mine - r) <= 5e-3 * abs(r), (key, mine, r)
    if dagger_first:
        assert abs(dloss - dmean) <= 1e-4 * abs(dmean)
        assert abs(dloss - float(GOLD["seq/dagger_returned"][0])) <= 5e-3 * abs(dloss)
    sdo = ac.state_dict()
    for k in orc.main_keys + (orc.adapt_keys if dagger_first else []):
        mine = orc.sd[k].detach() if k != "std" else torch.min(orc.sd[k].detach(), torch.tensor(1.0))
        assert_params_close(sdo[k], mine, 2e-4 * 4, k, max_frac=3e-3)      # TF32 on both sides: more near-zero gradients flip sign
    sde = est.state_dict()
    for k in orc.est_keys:
        assert_params_close(sde[k], orc.sd_est[k].detach(), 1e-4 * 4, k, max_frac=3e-3)


def test_dagger_then_update_matches_reference_golden():
    """update_dagger() then update() on the same PPO object against the reference's own run of that sequence (precise
    kernels): the adaptation encoder's stale post-clip gradient norm is part of every PPO minibatch's clip (ppo.py:274)."""
    T, N = 6, 32
    ac, est = _build(HID)
    ac.load_state_dict(_gold_sd("init/ac/")); est.load_state_dict(_gold_sd("init/est/"))
    ppo = _ppo(ac, est, N, T)
    st = {k[len("seq_storage/"):]: torch.from_numpy(GOLD[k]) for k in GOLD.files if k.startswith("seq_storage/")}
    _fill(ppo, st)
    perm = torch.from_numpy(GOLD["perm"]).to(DEV)
    dloss = ppo.update_dagger_with_indices(perm)
    assert abs(dloss - float(GOLD["seq/dagger_returned"][0])) <= 1e-4 * abs(dloss)
    stale = float(ac.main.state[7].item())
    assert 0.0 < stale <= 1.0 + 1e-5                      # post-clip squared norm of the adaptation gradients
    # first PPO minibatch alone: the norm the clip sees must be the reference's (main + stale), then the whole update
    out = ppo.update_with_indices(perm)
    ref = GOLD["seq/update_returned"]
    for mine, r in zip((out[0], out[1], out[2], out[4]), (ref[0], ref[1], ref[2], ref[4])):
        assert abs(mine - r) <= 1e-4 * abs(r), (mine, r)
    after, sd = _gold_sd("seq/ac/"), ac.state_dict()
    for k in after:
        assert_params_close(sd[k], after[k], 2e-4 * 4, k)
    # and the stale term is what makes it match: rescaled by every clip, it is now smaller than it was
    assert 0.0 < float(ac.main.state[7].item()) < stale


@pytest.mark.parametrize("offload_wgrads,graphs", [(False, False), (True, False), (True, True)],
                         ids=["inline-wgrads", "wgrads-on-their-own-stream", "wgrads-on-their-own-stream+cuda-graphs"])
def test_update_matches_reference_golden(offload_wgrads, graphs):
    """PPO.update on the reference's golden storage / weights / permutation (the schedule of the launches -- weight
    gradients inline or on a fifth stream, eager or graph-replayed -- must not change what is computed)."""
    T, N = 6, 32
    ac, est = _build(HID)
    ac.load_state_dict(_gold_sd("init/ac/")); est.load_state_dict(_gold_sd("init/est/"))
    ppo = _ppo(ac, est, N, T)
    ppo.offload_wgrads, ppo.use_graphs = offload_wgrads, graphs
    ppo.total_updates = 2.0
    st = {k[len("storage/"):]: torch.from_numpy(GOLD[k]) for k in GOLD.files if k.startswith("storage/")}
    _fill(ppo, st)
    v, sur, reg, coef, el = ppo.update_with_indices(torch.from_numpy(GOLD["perm"]).to(DEV))
    rv, rsur, rreg, rcoef, rel = GOLD["update/returned"]
    assert coef == rcoef
    for mine, ref in ((v, rv), (sur, rsur), (reg, rreg), (el, rel)):
        assert abs(mine - ref) <= 1e-4 * abs(ref), (mine, ref)
    move = 2e-4 * 4
    after, sd = _gold_sd("update/ac/"), ac.state_dict()
    for k in after:
        assert_params_close(sd[k], after[k], move, k)
    after_est, sde = _gold_sd("update/est/"), est.state_dict()
    for k in after_est:
        assert_params_close(sde[k], after_est[k], 1e-4 * 4, k)


def test_dagger_matches_reference_golden():
    T, N = 6, 32
    ac, est = _build(HID)
    ac.load_state_dict(_gold_sd("init/ac/")); est.load_state_dict(_gold_sd("init/est/"))
    ppo = _ppo(ac, est, N, T)
    st = {k[len("storage/"):]: torch.from_numpy(GOLD[k]) for k in GOLD.files if k.startswith("storage/")}
    _fill(ppo, st)
    loss = ppo.update_dagger_with_indices(torch.from_numpy(GOLD["perm"]).to(DEV))
    ref = float(GOLD["dagger/returned"][0])
    assert abs(loss - ref) <= 1e-4 * abs(ref)
    after, sd = _gold_sd("dagger/ac/"), ac.state_dict()
    for k in after:
        assert_params_close(sd[k], after[k], 2e-4 * 4, k)


@pytest.mark.parametrize("near,graphs", [(False, False), (True, False), (True, True)],
                         ids=["old-policy-far:lr-shrinks", "old-policy-near:lr-grows-then-holds", "near+cuda-graphs"])
def test_adaptive_schedule_matches_oracle(near, graphs):
    """schedule='adaptive' (ppo.py:233-246; oracle pinned to the reference in test_learner_oracle_vs_reference.py): the KL of
    every minibatch and the learning rate it leaves behind, decided on the device.  The rule is discrete (lr = 2e-4 x 1.5^k in
    fp64), so the final learning rate must be IDENTICAL; kl_mean to 1e-3 relative once both policies have taken the same number
    of Adam steps (first minibatch: 1e-4).  3 epochs: with `graphs` the third pass of every minibatch slot is a graph replay."""
    T, N, epochs = 6, 32, 3
    ac, est = _build(HID)
    sd, sd_est = _gold_sd("init/ac/"), _gold_sd("init/est/")
    ac.load_state_dict(sd); est.load_state_dict(sd_est)
    ppo = _ppo(ac, est, N, T, epochs=epochs, schedule='adaptive')
    ppo.use_graphs = graphs
    st = lu.random_storage(T, N, seed=31)
    if near:                                     # stored mu / sigma = the current policy's own -> KL ~ 12 x 1e-5
        f = lambda t: t.flatten(0, 1)
        with torch.no_grad():
            mu = lo.actor_mean(sd, f(st["obs"]), f(st["priv"]), f(st["true_est"]), f(st["scan"]))
        st["mu"], st["sigma"] = mu.view(T, N, -1).clone(), (mu * 0. + sd["std"]).view(T, N, -1).clone()
    _fill(ppo, st)
    perm = torch.randperm(T * N, generator=torch.Generator().manual_seed(8))
    orc = lo.LearnerOracle(sd, sd_est, lr=2e-4, est_lr=1e-4, desired_kl=0.01)
    mb = T * N // 2
    for _ in range(epochs):
        for i in range(2):
            orc.minibatch(lu.minibatch(st, perm[i * mb:(i + 1) * mb]), reg_coef=0.1)
    # the fixture must stay clear of the decision thresholds (0.005, 0.02), or the comparison below would be a coin toss
    assert all(min(abs(kl / 0.005 - 1), abs(kl / 0.02 - 1)) > 0.15 for kl, _ in orc.kl_log), orc.kl_log
    # first minibatch alone (no Adam step taken yet on either side)
    ppo._gather_storage(perm.to(DEV))
    ppo.reg_coef_dev.fill_(0.1)
    ppo._minibatch(0, mb)
    kl0 = float(ppo.kl_acc[1].item())
    assert abs(kl0 - orc.kl_log[0][0]) <= 1e-4 * abs(orc.kl_log[0][0]) and ppo.learning_rate == orc.kl_log[0][1]
    # the whole update from the same initial state
    ac.load_state_dict(sd); est.load_state_dict(sd_est)
    for g in (ac.main, est.group):
        g.exp_avg.zero_(); g.exp_avg_sq.zero_(); g.grads.zero_()
        g.state.copy_(torch.tensor([0.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0], dtype=torch.float64))
    ac.main.set_lr(2e-4); est.group.set_lr(1e-4)
    _fill(ppo, st)
    ppo.update_with_indices(perm.to(DEV))
    assert ppo.learning_rate == orc.lr, (ppo.learning_rate, orc.kl_log)
    kl_last = float(ppo.kl_acc[1].item())
    assert abs(kl_last - orc.kl_log[-1][0]) <= (1e-3 if not near else 5e-2) * abs(orc.kl_log[-1][0]), (kl_last, orc.kl_log[-1])
    assert float(ppo.kl_acc[0].item()) == 0.0
    if near:
        assert orc.lr == 2e-4 * 1.5 ** 4 and [lr for _, lr in orc.kl_log[-2:]] == [orc.lr, orc.lr]     # grew 4x, then held
    else:
        assert orc.lr == pytest.approx(2e-4 / 1.5 ** 6, rel=1e-12)


@pytest.mark.parametrize("dagger", [False, True], ids=["ppo-minibatch:all-five-chains", "dagger-minibatch"])
def test_production_minibatch_of_24576_samples_matches_tf32_oracle(dagger):
    """ONE REAL minibatch (M = 24 576 = 24 steps x 4096 envs / 4, the benchmark's shape) through the production tcgen05 path
    against the oracle in numerics('tf32'): losses <= 1e-4; every gradient tensor <= 1e-3 in relative Frobenius distance
    (rms of the error over rms of the gradient) and <= 1e-2 of its rms in the WORST element.  Every single kernel matches its
    emulation to fp32-accumulation distance (test_tcgen05_operands_are_truncated_tf32, test_mma_sync_heads_round_to_nearest:
    1e-6 .. 1e-5); the chain cannot, because truncation to TF32 is discontinuous: an operand that differs by one fp32 ulp
    between two implementations (accumulation order, __expf) truncates to a different TF32 value with probability 2^-13 -- a
    2^-10 step -- and four layers of forward + backward compound that to ~3e-4 rms with heavy tails over the 321 k elements
    of the widest tensor (measured: 2.6e-4 / 2.2e-3 at M = 24 576, 4.5e-3 worst element at M = 384).  The fp32 oracle is
    10-30x farther away (test_tf32_numerics_mode bounds that distance at 3e-2)."""
    test_gradients_match_oracle_full_size(dagger, False, T=24, N=1024)


@pytest.mark.parametrize("dagger,precise", [(False, True), (True, True), (False, False), (True, False)])
def test_gradients_match_oracle_full_size(dagger, precise, T=4, N=96):
    """one minibatch at the real layer sizes (go2_parkour): flat gradients vs torch autograd on the oracle.
    precise=True: 3xTF32 mma.sync kernels against the fp32 oracle; precise=False: the PRODUCTION tcgen05 TF32 path against
    the oracle in numerics('tf32') (operands rounded exactly as each kernel rounds them, fp32 accumulation).  Both: 1e-3 of
    each tensor's rms for the gradients, 1e-4 for the losses."""
    hid = dict(actor=[512, 256, 128], critic=[512, 256, 128], priv=[64, 20], scan=[128, 64], est=[256, 128])
    ac, est = _build(hid, precise=precise)
    ppo = _ppo(ac, est, N, T, epochs=1, mbs=1)
    st = lu.random_storage(T, N, seed=21)
    _fill(ppo, st)
    sd = {k: v.cpu() for k, v in ac.state_dict().items()}
    sd_est = {k: v.cpu() for k, v in est.state_dict().items()}
    perm = torch.randperm(T * N, generator=torch.Generator().manual_seed(2))
    orc = lo.LearnerOracle(sd, sd_est)
    ppo._adam = lambda group: None                      # keep the raw gradients in the flat buffers
    ppo._gather_storage(perm.to(DEV))
    ppo.reg_coef_dev.fill_(0.07)
    ppo.loss_sums.zero_()
    b = lu.minibatch(st, perm)
    with lo.numerics("fp32" if precise else "tf32"):
        if dagger:
            ppo._dagger_minibatch(0, T * N)
            orc.dagger_minibatch(b)
            groups = [(ac.adapt, ac, orc.adapt_keys)]
        else:
            ppo._minibatch(0, T * N)
            logs = orc.minibatch(b, reg_coef=0.07)
            groups = [(ac.main, ac, orc.main_keys), (est.group, est, orc.est_keys)]
            sums = (ppo.loss_sums / (T * N)).tolist()
            for mine, key in ((sums[0], "surrogate"), (sums[1], "value"), (sums[2], "reg"), (sums[3], "entropy"), (sums[4], "estimator")):
                assert abs(mine - logs[key]) <= 1e-4 * abs(logs[key]), key
    torch.cuda.synchronize()
    for group, owner, keys in groups:
        # read gradients back in checkpoint layout by viewing the grads buffer through state_dict()
        saved = group.params
        group.params = group.grads
        gsd = {k: v.cpu() for k, v in owner.state_dict().items()}
        group.params = saved
        for k in keys:
            ref = orc.last_grads[k]
            if precise:
                assert scale_err(gsd[k], ref) <= 1e-3, (k, scale_err(gsd[k], ref))
            else:       # production: see test_production_minibatch_of_24576_samples_matches_tf32_oracle for why two measures
                assert fro_err(gsd[k], ref) <= 1e-3 and scale_err(gsd[k], ref) <= 1e-2, (k, fro_err(gsd[k], ref), scale_err(gsd[k], ref))


@pytest.mark.parametrize("precise,chain", [(True, False), (False, False), (False, True)],
                         ids=["3xTF32-vs-fp32-oracle", "production-tcgen05-vs-tf32-oracle", "production+one-launch-chains"])
def test_act_and_storage_vs_oracle(precise, chain):
    """PPO.act / process_env_step / compute_returns on the GPU vs the oracle (keyed action noise)."""
    hid = dict(actor=[512, 256, 128], critic=[512, 256, 128], priv=[64, 20], scan=[128, 64], est=[256, 128])
    T, N = 3, 200
    ac, est = _build(hid, precise=precise)
    ac.k.use_chain = est.k.use_chain = chain
    lo._CHAIN[0] = chain
    ppo = _ppo(ac, est, N, T)
    sd = {k: v.cpu() for k, v in ac.state_dict().items()}
    sd_est = {k: v.cpu() for k, v in est.state_dict().items()}
    st = lu.random_storage(T, N, seed=31)
    g = torch.Generator().manual_seed(1)
    rews, dones, tmo = torch.rand(T, N, generator=g), torch.rand(T, N, generator=g) < 0.1, torch.rand(T, N, generator=g) < 0.05
    vals = []
    for t in range(T):
        mode = t == 1
        a = ppo.act(*(st[k][t].to(DEV) for k in ("obs", "priv", "critic_obs", "true_est", "scan")), adaptation_mode=mode)
        with lo.numerics("fp32" if precise else "tf32"):
            ra, rv, rlp, rmu, rsig = lo.ppo_act(sd, sd_est, st["obs"][t], st["priv"][t], st["critic_obs"][t], st["scan"][t], ppo.seed, t, mode)
        tol = (lambda x, y: scale_err(x, y) <= 1e-4) if precise else (lambda x, y: fro_err(x, y) <= 5e-4 and scale_err(x, y) <= 5e-3)
        assert tol(a, ra) and tol(ppo.storage.values[t], rv), (scale_err(a, ra), scale_err(ppo.storage.values[t], rv))
        assert tol(ppo.storage.actions_log_prob[t, :, 0], rlp) and tol(ppo.storage.mu[t], rmu)
        assert torch.equal(ppo.storage.observations[t].cpu(), st["obs"][t]) and torch.equal(ppo.storage.privileged_observations[t].cpu(), st["priv"][t])
        assert torch.equal(ppo.storage.critic_observations[t].cpu(), st["critic_obs"][t]) and torch.equal(ppo.storage.scan_observations[t].cpu(), st["scan"][t])
        ppo.process_env_step(rews[t].to(DEV), dones[t].to(DEV), {"time_outs": tmo[t].to(DEV)})
        vals.append(ppo.storage.values[t].cpu())
        ref_r = lo.bootstrap_rewards(rews[t], vals[-1], tmo[t], 0.99)
        assert torch.equal(ppo.storage.rewards[t, :, 0].cpu(), ref_r)
    ppo.compute_returns(st["critic_obs"][0].to(DEV))
    torch.cuda.synchronize()
    lo._CHAIN[0] = False
    s = ppo.storage
    ret, adv = lo.compute_returns(s.rewards.cpu(), s.dones.cpu(), s.values.cpu(), ppo.last_values.cpu(), 0.99, 0.95)
    assert torch.equal(s.returns.cpu(), ret) and gu.rel_err(s.advantages.cpu().numpy(), adv.numpy()) <= 1e-5


@pytest.mark.parametrize("M", [4096, 1000, 100])
def test_mlp_chain_kernel_matches_per_layer_launches(M):
    """b200_tc_mlp_forward (one persistent launch for a whole Linear / ELU chain, grid-wide barriers between the layers)
    against one b200_tc_linear_forward launch per layer: the same tcgen05 arithmetic in the same k order -> bit-identical for
    every layer the per-layer path runs on tcgen05; a 3-wide head (zero-filled weight rows) against fp64 of truncated operands.
    Launched three times and replayed from a CUDA graph: the barrier epochs advance on the device."""
    lib = _lib.lib()
    g = torch.Generator().manual_seed(M)
    dims = [572, 256, 128, 3]
    X = torch.randn(M, dims[0], generator=g).to(DEV)
    Ws = [(torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5).to(DEV) for i in range(3)]
    bs = [torch.randn(dims[i + 1], generator=g).to(DEV) for i in range(3)]
    ld = lambda n: (n + 3) // 4 * 4
    Ys = [torch.zeros(M, ld(dims[i + 1]), device=DEV) for i in range(3)]
    ref = [torch.zeros(M, ld(dims[i + 1]), device=DEV) for i in range(2)]
    st = _lib.stream_ptr()
    inp, ldi = X, dims[0]
    for i in range(2):
        _lib.check(lib.b200_tc_linear_forward(inp.data_ptr(), ldi, Ws[i].data_ptr(), dims[i], bs[i].data_ptr(), ref[i].data_ptr(), ld(dims[i + 1]),
                                              M, dims[i + 1], dims[i], 1, st))
        inp, ldi = ref[i], ld(dims[i + 1])
    arr = (_lib.MlpLayer * 3)()
    for i in range(3):
        arr[i].W, arr[i].bias, arr[i].Y, arr[i].ldw, arr[i].ldy = Ws[i].data_ptr(), bs[i].data_ptr(), Ys[i].data_ptr(), dims[i], ld(dims[i + 1])
        arr[i].N, arr[i].K, arr[i].act = dims[i + 1], dims[i], int(i < 2)
    sync = torch.zeros(4, dtype=torch.int32, device=DEV)
    run = lambda: _lib.check(lib.b200_tc_mlp_forward(arr, 3, X.data_ptr(), dims[0], M, sync.data_ptr(), 74, _lib.stream_ptr()))
    for _ in range(3):
        for y in Ys:
            y.zero_()
        run()
        torch.cuda.synchronize()
        assert torch.equal(Ys[0], ref[0]) and torch.equal(Ys[1], ref[1])
        head = (lo.tf32_trunc(ref[1][:, :128].cpu()).double() @ lo.tf32_trunc(Ws[2].cpu()).double().t() + bs[2].cpu().double()).float()
        assert scale_err(Ys[2][:, :3], head) <= 2e-5 and float(Ys[2][:, 3:].abs().max()) == 0.0
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=side):
            run()
    for _ in range(3):
        Ys[2].zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert scale_err(Ys[2][:, :3], head) <= 2e-5
    per_launch = 3 * min(74, max((M + 127) // 128 * 4, 1))
    assert int(sync[1].item()) == 6 * per_launch and int(sync[0].item()) == int(sync[1].item()) and int(sync[2].item()) == 0
