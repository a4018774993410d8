import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from legged_gym_custom_b200 import _lib
from oracle import learner_oracle as lo
lib = _lib.lib(); DEV = "cuda:0"
for (M, N, K, ldy) in ((512, 3, 128, 4), (512, 6, 128, 8), (512, 7, 64, 8), (512, 8, 128, 8), (512, 5, 128, 8)):
    g = torch.Generator().manual_seed(3)
    X, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    Xd, Wd, Y = X.to(DEV), W.to(DEV), torch.zeros(M, ldy, device=DEV)
    _lib.check(lib.b200_linear_forward(Xd.data_ptr(), K, Wd.data_ptr(), K, None, Y.data_ptr(), ldy, M, N, K, 0, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    Y = Y.cpu()[:, :N].double()
    rr = lo.tf32_rna(X).double() @ lo.tf32_rna(W).double().t()
    rt = lo.tf32_trunc(X).double() @ lo.tf32_trunc(W).double().t()
    rf = X.double() @ W.double().t()
    print(M, N, K, "per-col max err vs rna", [(float((Y[:, c] - rr[:, c]).abs().max())) for c in range(N)])
    print("   vs trunc", float((Y - rt).abs().max()), "vs fp32", float((Y - rf).abs().max()), "vs rna", float((Y - rr).abs().max()))
