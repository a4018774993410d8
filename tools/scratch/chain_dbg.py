import sys; sys.path.insert(0,"tests"); sys.path.insert(0,".")
import torch
from legged_gym_custom_b200 import _lib
from oracle import learner_oracle as lo
lib=_lib.lib(); DEV="cuda:0"
for M in (100, 128, 1000, 1024):
    g=torch.Generator().manual_seed(M)
    dims=[572,256,128,3]
    X=torch.randn(M,dims[0],generator=g).to(DEV)
    Ws=[(torch.randn(dims[i+1],dims[i],generator=g)/dims[i]**0.5).to(DEV) for i in range(3)]
    bs=[torch.randn(dims[i+1],generator=g).to(DEV) for i in range(3)]
    ld=lambda n:(n+3)//4*4
    Ys=[torch.zeros(M,ld(dims[i+1]),device=DEV) for i in range(3)]
    arr=(_lib.MlpLayer*3)()
    for i in range(3):
        arr[i].W,arr[i].bias,arr[i].Y,arr[i].ldw,arr[i].ldy=Ws[i].data_ptr(),bs[i].data_ptr(),Ys[i].data_ptr(),dims[i],ld(dims[i+1]); arr[i].N,arr[i].K,arr[i].act=dims[i+1],dims[i],int(i<2)
    sync=torch.zeros(4,dtype=torch.int32,device=DEV)
    _lib.check(lib.b200_tc_mlp_forward(arr,3,X.data_ptr(),dims[0],M,sync.data_ptr(),74,_lib.stream_ptr())); torch.cuda.synchronize()
    head=(lo.tf32_trunc(Ys[1][:,:128].cpu()).double()@lo.tf32_trunc(Ws[2].cpu()).double().t()+bs[2].cpu().double()).float()
    d=(Ys[2][:,:3].cpu()-head).abs()
    badrows=(d.max(1).values>1e-3).nonzero()[:,0]
    print("M",M,"sync",sync.tolist(),"bad rows",len(badrows), badrows[:8].tolist(), badrows[-4:].tolist() if len(badrows) else "", "pad col max", float(Ys[2][:,3:].abs().max()))
    if len(badrows): print("  got",Ys[2][badrows[0]].tolist(),"want",head[badrows[0]].tolist())
