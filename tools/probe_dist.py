"""Timing of the fused data-parallel optimiser step alone (csrc/dist_adam.cu) against NCCL all-reduce + b200_clip_adam:
torchrun --nproc-per-node N tools/probe_dist.py"""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legged_gym_custom_b200 import _lib  # noqa: E402
from legged_gym_custom_b200.dist import FusedDistAdam  # noqa: E402
from legged_gym_custom_b200.networks import FlatGroup  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device(f"cuda:{rank}")
dist.init_process_group("nccl", device_id=dev)
lib = _lib.lib()
for n in (1061976, 181132, 5200):          # main, estimator, adaptation groups (floats)
    n = (n + 7) // 8 * 8
    for mc in (1, 0):
        os.environ["B200GYM_DIST_MULTICAST"] = str(mc)
        g = FlatGroup(); g.add("w.weight", n // 8, 8); g.finalize(dev, 2e-4)
        g.params.normal_()
        f = FusedDistAdam(g, dist.group.WORLD, 1.0)
        ref = FlatGroup(); ref.add("w.weight", n // 8, 8); ref.finalize(dev, 2e-4)

        def fused():
            f.step()

        def nccl():
            dist.all_reduce(ref.grads)
            lib.b200_clip_adam(ref.params.data_ptr(), ref.grads.data_ptr(), ref.exp_avg.data_ptr(), ref.exp_avg_sq.data_ptr(), ref.n,
                               C.c_void_p(ref.state.data_ptr()), 1.0 / world, 1.0, 0.9, 0.999, 1e-8, _lib.stream_ptr())
        out = []
        for fn in (fused, nccl):
            for _ in range(5):
                fn()
            torch.cuda.synchronize(); dist.barrier()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(20):
                    fn()
            gr.replay(); torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                gr.replay()
            e1.record(); torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1) * 1e3 / 200)
        if rank == 0:
            print(f"n={n:8d} world={world} multicast={f.multicast}: fused {out[0]:6.1f} us/step   nccl all-reduce + sumsq + clip_adam {out[1]:6.1f} us/step", flush=True)
        dist.barrier()
torch.cuda.synchronize()
os._exit(0)
