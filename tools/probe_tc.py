"""Correctness + timing probe of the tcgen05 GEMMs (csrc/mlp_tcgen05.cu) against torch fp32 and the mma.sync baseline."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legged_gym_custom_b200 import _lib  # noqa: E402

lib = _lib.lib()
if os.environ.get("B200_PAIR") is not None:
    lib.b200_tc_set_pair_mode(int(os.environ["B200_PAIR"]))
if os.environ.get("B200_CPS") is not None:
    lib.b200_tc_set_ctas_per_sm(int(os.environ["B200_CPS"]))
if os.environ.get("B200_TMA_EPI") is not None:
    lib.b200_tc_set_tma_epilogue(int(os.environ["B200_TMA_EPI"]))
if os.environ.get("B200_WGRAD_PAIRS") is not None:
    lib.b200_tc_set_wgrad_pairs(int(os.environ["B200_WGRAD_PAIRS"]))
DEV = "cuda:0"
ld = lambda k: (k + 3) // 4 * 4
p = lambda t: t.data_ptr()


def err(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().pow(2).mean().sqrt())


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


modes = sys.argv[1:] or ["fwd", "dgrad", "wgrad"]
shapes = [(4096, 512, 627), (24576, 512, 627), (24576, 256, 512), (24576, 128, 256), (24576, 512, 736), (24576, 256, 572), (333, 128, 132),
          (1000, 64, 128), (500, 32, 64), (24576, 12, 128), (24576, 20, 64), (24576, 128, 132), (24576, 64, 128), (24500, 256, 512),
          (65536, 512, 627)]
torch.backends.cuda.matmul.allow_tf32 = False
for M, N, K in shapes:
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    X = torch.zeros(M, ld(K), device=DEV); X[:, :K] = torch.randn(M, K, device=DEV, generator=g)
    W = torch.zeros(N, ld(K), device=DEV); W[:, :K] = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    b = torch.randn(N, device=DEV, generator=g)
    dY = torch.zeros(M, ld(N), device=DEV); dY[:, :N] = torch.randn(M, N, device=DEV, generator=g)
    st = _lib.stream_ptr()
    line = f"M={M:6d} N={N:4d} K={K:4d} |"
    if "fwd" in modes:
        Y = torch.zeros(M, ld(N), device=DEV)
        rc = lib.b200_tc_linear_forward(p(X), ld(K), p(W), ld(K), p(b), p(Y), ld(N), M, N, K, 1, st)
        torch.cuda.synchronize()
        assert rc == 0, lib.b200_last_error()
        ref = torch.nn.functional.elu(X[:, :K] @ W[:, :K].t() + b)
        e = err(Y[:, :N], ref)
        t_tc = timeit(lambda: lib.b200_tc_linear_forward(p(X), ld(K), p(W), ld(K), p(b), p(Y), ld(N), M, N, K, 1, st))
        t_mma = timeit(lambda: lib.b200_linear_forward(p(X), ld(K), p(W), ld(K), p(b), p(Y), ld(N), M, N, K, 1, 0, st))
        line += f" fwd err {e:.1e} tc {t_tc:7.1f}us ({2 * M * N * K / t_tc / 1e6:6.1f} TF) mma {t_mma:7.1f}us |"
    if "dgrad" in modes:
        Yp = torch.randn(M, ld(K), device=DEV, generator=g)
        dX = torch.ones(M, ld(K), device=DEV)
        rc = lib.b200_tc_linear_dgrad(p(dY), ld(N), p(W), ld(K), p(Yp), ld(K), p(dX), ld(K), M, N, K, 1, st)
        torch.cuda.synchronize()
        assert rc == 0, lib.b200_last_error()
        ref = 1.0 + (dY[:, :N] @ W[:, :K]) * torch.where(Yp[:, :K] > 0, torch.ones((), device=DEV), Yp[:, :K] + 1.0)
        e = err(dX[:, :K], ref)
        t_tc = timeit(lambda: lib.b200_tc_linear_dgrad(p(dY), ld(N), p(W), ld(K), p(Yp), ld(K), p(dX), ld(K), M, N, K, 0, st))
        t_mma = timeit(lambda: lib.b200_linear_dgrad(p(dY), ld(N), p(W), ld(K), p(Yp), ld(K), p(dX), ld(K), M, N, K, 0, 0, st))
        line += f" dgrad err {e:.1e} tc {t_tc:7.1f}us ({2 * M * N * K / t_tc / 1e6:6.1f} TF) mma {t_mma:7.1f}us |"
        db = torch.zeros(ld(K), device=DEV)
        rc = lib.b200_tc_linear_dgrad_bias(p(dY), ld(N), p(W), ld(K), p(Yp), ld(K), p(dX), ld(K), M, N, K, 0, p(db), st)
        torch.cuda.synchronize()
        assert rc == 0, lib.b200_last_error()
        eb = err(db[:K], (ref - 1.0).sum(0))
        t_b = timeit(lambda: lib.b200_tc_linear_dgrad_bias(p(dY), ld(N), p(W), ld(K), p(Yp), ld(K), p(dX), ld(K), M, N, K, 0, p(db), st))
        line += f" +dbias err {eb:.1e} {t_b:7.1f}us |"
    if "wgrad" in modes:
        dW = torch.full((N, ld(K)), 0.5, device=DEV)
        rc = lib.b200_tc_linear_wgrad(p(dY), ld(N), p(X), ld(K), p(dW), ld(K), M, N, K, st)
        torch.cuda.synchronize()
        assert rc == 0, lib.b200_last_error()
        ref = 0.5 + dY[:, :N].t() @ X[:, :K]
        e = err(dW[:, :K], ref)
        t_tc = timeit(lambda: lib.b200_tc_linear_wgrad(p(dY), ld(N), p(X), ld(K), p(dW), ld(K), M, N, K, st))
        db = torch.zeros(N, device=DEV)
        t_mma = timeit(lambda: lib.b200_linear_wgrad(p(dY), ld(N), p(X), ld(K), p(dW), ld(K), p(db), M, N, K, 0, st))
        line += f" wgrad err {e:.1e} tc {t_tc:7.1f}us ({2 * M * N * K / t_tc / 1e6:6.1f} TF) mma {t_mma:7.1f}us"
    print(line, flush=True)
