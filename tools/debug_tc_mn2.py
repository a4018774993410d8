"""Which dimension breaks the MN-major paths? exact small-integer operands, mismatch map per 32x32 block."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legged_gym_custom_b200 import _lib
lib = _lib.lib(); DEV = "cuda:0"; p = lambda t: t.data_ptr(); st = _lib.stream_ptr()
g = torch.Generator(device=DEV).manual_seed(0)
ri = lambda *s: torch.randint(-2, 3, s, device=DEV, generator=g).float()

def blockmap(a, b):
    bad = (a != b)
    R, C = bad.shape
    rows = []
    for r in range(0, R, 32):
        rows.append("".join("X" if bad[r:r + 32, c:c + 32].any() else "." for c in range(0, C, 32)))
    return " ".join(rows[:8]) + (" ..." if len(rows) > 8 else "")

for (M, N, K) in [(128, 32, 64), (128, 64, 32), (128, 64, 64), (256, 128, 128), (128, 32, 256), (128, 256, 32)]:
    dY, W = ri(M, N), ri(N, K)
    dX = torch.zeros(M, K, device=DEV)
    rc = lib.b200_tc_linear_dgrad(p(dY), N, p(W), K, None, 0, p(dX), K, M, N, K, 0, st); torch.cuda.synchronize()
    print(f"dgrad M={M} N(red)={N} K(out)={K}: rc={rc} bad blocks [rows of 32 x cols of 32]: {blockmap(dX, dY @ W)}")
for (M, N, K) in [(32, 32, 64), (32, 64, 32), (64, 32, 32), (64, 64, 64), (32, 128, 32), (32, 256, 32), (256, 128, 128)]:
    dY, X = ri(M, N), ri(M, K)
    dW = torch.zeros(N, K, device=DEV)
    rc = lib.b200_tc_linear_wgrad(p(dY), N, p(X), K, p(dW), K, M, N, K, st); torch.cuda.synchronize()
    print(f"wgrad M(red)={M} N(rows)={N} K(cols)={K}: rc={rc} bad: {blockmap(dW, dY.t() @ X)}")
