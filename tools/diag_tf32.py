"""diagnostic: where the production path and the TF32 oracle differ (per tensor, per column block) + accumulation precision vs K"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import learner_util as lu
import test_learner_gpu as tg
from legged_gym_custom_b200 import _lib
from oracle import learner_oracle as lo
DEV = "cuda:0"
lib = _lib.lib()
for (M, N, K) in [(2048, 256, 64), (2048, 256, 512), (2048, 256, 4096), (24576, 256, 512)]:
    g = torch.Generator().manual_seed(9)
    X, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    Xd, Wd, Y = X.to(DEV), W.to(DEV), torch.zeros(M, N, device=DEV)
    _lib.check(lib.b200_tc_linear_forward(Xd.data_ptr(), K, Wd.data_ptr(), K, None, Y.data_ptr(), N, M, N, K, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref_t = (lo.tf32_trunc(X).double() @ lo.tf32_trunc(W).double().t())
    ref32 = lo.tf32_trunc(X) @ lo.tf32_trunc(W).t()
    print("fwd", M, N, K, "gpu vs fp64-of-trunc", tg.scale_err(Y, ref_t.float()), "cpu-fp32 vs fp64", tg.scale_err(ref32, ref_t.float()))
    # wgrad-like: reduction over M
    dY = torch.randn(M, N, generator=g)
    dYd, dW = dY.to(DEV), torch.zeros(N, K, device=DEV)
    _lib.check(lib.b200_tc_linear_wgrad(dYd.data_ptr(), N, Xd.data_ptr(), K, dW.data_ptr(), K, M, N, K, _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = lo.tf32_trunc(dY).double().t() @ lo.tf32_trunc(X).double()
    print("wgrad", M, N, K, tg.scale_err(dW, ref.float()))

hid = dict(actor=[512, 256, 128], critic=[512, 256, 128], priv=[64, 20], scan=[128, 64], est=[256, 128])
T, N = 24, 1024
ac, est = tg._build(hid, precise=False)
ppo = tg._ppo(ac, est, N, T, epochs=1, mbs=1)
st = lu.random_storage(T, N, seed=21)
tg._fill(ppo, st)
sd = {k: v.cpu() for k, v in ac.state_dict().items()}
sd_est = {k: v.cpu() for k, v in est.state_dict().items()}
perm = torch.randperm(T * N, generator=torch.Generator().manual_seed(2))
orc = lo.LearnerOracle(sd, sd_est)
ppo._adam = lambda group: None
ppo._gather_storage(perm.to(DEV))
ppo.reg_coef_dev.fill_(0.07)
ppo.loss_sums.zero_()
b = lu.minibatch(st, perm)
with lo.numerics("tf32"):
    ppo._minibatch(0, T * N)
    logs = orc.minibatch(b, reg_coef=0.07)
torch.cuda.synchronize()
print("losses", (ppo.loss_sums / (T * N)).tolist(), logs)
for group, owner, keys in [(ac.main, ac, orc.main_keys), (est.group, est, orc.est_keys)]:
    saved = group.params; group.params = group.grads
    gsd = {k: v.cpu() for k, v in owner.state_dict().items()}
    group.params = saved
    for k in keys:
        ref = orc.last_grads[k]
        print(f"{k:50s} err/rms {tg.scale_err(gsd[k], ref):.2e}  rms {float(ref.pow(2).mean().sqrt()):.3e}")
    if "actor.0.weight" in keys:
        a, r = gsd["actor.0.weight"], orc.last_grads["actor.0.weight"]
        for name, lo_, hi_ in (("obs", 0, 572), ("latent", 572, 592), ("scan", 592, 624), ("est", 624, 627)):
            d = (a[:, lo_:hi_] - r[:, lo_:hi_]).abs()
            print("  actor.0 cols", name, "max", float(d.max()), "rms err", float(d.pow(2).mean().sqrt()), "rms ref", float(r[:, lo_:hi_].pow(2).mean().sqrt()))

# --- the mma.sync heads (N < 8): which rounding do they apply?
for (M, N, K) in [(512, 1, 128), (512, 3, 128), (512, 4, 256)]:
    g = torch.Generator().manual_seed(3)
    X, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    Xd, Wd, Y = X.to(DEV), W.to(DEV), torch.zeros(M, 4, device=DEV)
    _lib.check(lib.b200_linear_forward(Xd.data_ptr(), K, Wd.data_ptr(), K, None, Y.data_ptr(), 4, M, N, K, 0, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    for name, r in (("rna", lo.tf32_rna), ("trunc", lo.tf32_trunc), ("fp32", lambda x: x)):
        ref = (r(X).double() @ r(W).double().t()).float()
        print("head fwd", (M, N, K), name, tg.scale_err(Y[:, :N], ref), tg.fro_err(Y[:, :N], ref))
