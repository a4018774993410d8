"""One-launch MLP chain (b200_tc_mlp_forward) against one launch per layer, at rollout size, both replayed from CUDA graphs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legged_gym_custom_b200 import _lib  # noqa: E402

lib = _lib.lib()
DEV = "cuda:0"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ld = lambda n: (n + 3) // 4 * 4
chains = {"estimator": [572, 256, 128, 3], "actor": [628, 512, 256, 128, 12], "critic": [736, 512, 256, 128, 1], "scan": [132, 128, 64, 32],
          "priv": [32, 64, 20, 20]}


def graph_time(fn, reps=20, replays=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * replays)


for name, dims in chains.items():
    n = len(dims) - 1
    X = torch.randn(M, dims[0], device=DEV)
    Ws = [torch.randn(dims[i + 1], ld(dims[i]), device=DEV) / dims[i] ** 0.5 for i in range(n)]
    bs = [torch.randn(dims[i + 1], device=DEV) for i in range(n)]
    Ys = [torch.zeros(M, ld(dims[i + 1]), device=DEV) for i in range(n)]
    arr = (_lib.MlpLayer * n)()
    for i in range(n):
        arr[i].W, arr[i].bias, arr[i].Y, arr[i].ldw, arr[i].ldy = Ws[i].data_ptr(), bs[i].data_ptr(), Ys[i].data_ptr(), ld(dims[i]), ld(dims[i + 1])
        arr[i].N, arr[i].K, arr[i].act = dims[i + 1], dims[i], int(i < n - 1)
    sync = torch.zeros(4, dtype=torch.int32, device=DEV)

    def per_layer():
        inp, ldi = X, dims[0]
        for i in range(n):
            if dims[i + 1] >= 8:
                lib.b200_tc_linear_forward(inp.data_ptr(), ldi, Ws[i].data_ptr(), ld(dims[i]), bs[i].data_ptr(), Ys[i].data_ptr(), ld(dims[i + 1]), M,
                                           dims[i + 1], dims[i], int(i < n - 1), _lib.stream_ptr())
            else:
                lib.b200_linear_forward(inp.data_ptr(), ldi, Ws[i].data_ptr(), ld(dims[i]), bs[i].data_ptr(), Ys[i].data_ptr(), ld(dims[i + 1]), M,
                                        dims[i + 1], dims[i], int(i < n - 1), 0, _lib.stream_ptr())
            inp, ldi = Ys[i], ld(dims[i + 1])
    t_layers = graph_time(per_layer)
    line = f"{name:10s} {dims}: per-layer launches {t_layers:6.1f} us |"
    for cap in (37, 74, 148, 296):
        t = graph_time(lambda: lib.b200_tc_mlp_forward(arr, n, X.data_ptr(), dims[0], M, sync.data_ptr(), cap, _lib.stream_ptr()))
        line += f" chain@{cap}: {t:6.1f}"
    print(line, flush=True)
