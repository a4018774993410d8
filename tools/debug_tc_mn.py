"""Decode how the MN-major tcgen05 descriptors read shared memory, using one-hot / index-coded operands."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legged_gym_custom_b200 import _lib
lib = _lib.lib(); DEV = "cuda:0"; p = lambda t: t.data_ptr(); st = _lib.stream_ptr()
torch.set_printoptions(linewidth=250, sci_mode=False)
# dgrad: dX[M,K] = dY[M,N] . W[N,K];  M=128, N=32 (one k-block), K=32 (one MN chunk)
M, N, K = 128, 32, 32
dY = torch.zeros(M, N, device=DEV)
dY[torch.arange(M), torch.arange(M) % N] = 1.0            # row i selects reduction index i % 32
W = (torch.arange(N, device=DEV).float()[:, None] * 100 + torch.arange(K, device=DEV).float()[None, :]).contiguous()
dX = torch.zeros(M, K, device=DEV)
rc = lib.b200_tc_linear_dgrad(p(dY), N, p(W), K, None, 0, p(dX), K, M, N, K, 0, st); torch.cuda.synchronize()
print("dgrad rc", rc, "expected row i = W[i%32] = 100*(i%32) + k")
print(dX[:12].int())
print("rows 32..35", dX[32:36].int())
# wgrad: dW[N,K] = dY^T X ; M=32 (one k block), N=128?? use N=32,K=32
M, N, K = 32, 32, 32
dY = torch.zeros(M, N, device=DEV); dY[torch.arange(M), torch.arange(M)] = 1.0     # identity -> dW = X
X = (torch.arange(M, device=DEV).float()[:, None] * 100 + torch.arange(K, device=DEV).float()[None, :]).contiguous()
dW = torch.zeros(N, K, device=DEV)
rc = lib.b200_tc_linear_wgrad(p(dY), N, p(X), K, p(dW), K, M, N, K, st); torch.cuda.synchronize()
print("wgrad rc", rc, "expected dW = X (row n = 100 n + k)")
print(dW[:12].int())
# wgrad with X = ones -> dW[n,k] = sum_m dY[m,n]: tests A alone
dY = (torch.arange(M, device=DEV).float()[:, None] * 100 + torch.arange(N, device=DEV).float()[None, :]).contiguous()
Xi = torch.zeros(M, K, device=DEV); Xi[torch.arange(M), torch.arange(M)] = 1.0     # dW[n,k] = dY[k,n] = 100 k + n
dW = torch.zeros(N, K, device=DEV)
rc = lib.b200_tc_linear_wgrad(p(dY), N, p(Xi), K, p(dW), K, M, N, K, st); torch.cuda.synchronize()
print("wgrad(A coded) expected dW[n,k] = 100 k + n")
print(dW[:12].int())
