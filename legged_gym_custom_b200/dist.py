"""Multi-GPU plumbing: one process per GPU, envs sharded per rank, and ONE exchange step -- the sum all-reduce of each
optimiser's flat gradient buffer over NCCL (NVLink / NVSwitch) before the fused clip + Adam kernel, which folds the
1/world_size into its gradient scale.  Every rank then applies the same reduced gradient, so the
replicated parameters stay bit-identical without any parameter broadcast after start-up.

The reference has no distributed path at all (single process, single device); this is new, and deliberately tiny.
"""
import os

import torch
import torch.distributed as dist


def env_rank_info():
    """(rank, world_size, local_rank) from the torchrun / torch.distributed.run environment."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None, device=None):
    """Initialise the default group from MASTER_ADDR / MASTER_PORT (use 127.0.0.1 on a single box)."""
    rank, world, _ = env_rank_info()
    if world == 1:
        return None
    if not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return dist.group.WORLD


def shard_seed(seed, rank):
    """each rank owns its envs, its RNG keys and its minibatch permutation"""
    return int(seed) + 7919 * int(rank)


def allreduce_flat_grads(groups, process_group):
    """sum all-reduce of the flat gradient buffer of each FlatGroup (in place)."""
    if process_group is None:
        return
    for g in groups:
        dist.all_reduce(g.grads, op=dist.ReduceOp.SUM, group=process_group)


def broadcast_parameters(groups, process_group, src=0):
    """start-up only: make every rank start from rank `src`'s parameters and optimiser state."""
    if process_group is None:
        return
    for g in groups:
        for t in (g.params, g.exp_avg, g.exp_avg_sq, g.state):
            dist.broadcast(t, src=src, group=process_group)


def shard_bounds(n, world, rank):
    """[lo, hi) in floats of rank's shard of an n-float flat buffer (n % 4 == 0): float4-granular, as b200_dist_adam cuts it"""
    n4 = n // 4
    per = (n4 + world - 1) // world
    lo = min(per * rank, n4)
    return 4 * lo, 4 * min(lo + per, n4)


class FusedDistAdam:
    """One FlatGroup under the fused exchange + clip + Adam kernel (csrc/dist_adam.cu, b200_dist_adam): moves the group's
    params / grads / moments into SYMMETRIC memory (torch.distributed._symmetric_memory: one allocation per rank, all of them
    mapped into every rank, plus the NVSwitch multicast address when the fabric offers one) and keeps the argument block of
    the kernel.  `step()` is one launch; there is no NCCL call on the path."""

    def __init__(self, group, process_group, max_grad_norm, betas=(0.9, 0.999), eps=1e-8):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm

        from . import _lib
        self.lib, self.group, self.pg = _lib.lib(), group, process_group
        W, r = dist.get_world_size(process_group), dist.get_rank(process_group)
        self.world, self.rank = W, r
        n, dev = group.n, group.params.device
        self._handles = {}

        def alloc(name, numel, dtype, init=None):
            t = symm.empty(numel, dtype=dtype, device=dev)
            t.zero_() if init is None else t.copy_(init)
            self._handles[name] = symm.rendezvous(t, process_group)
            return t
        group.params = alloc("params", n, torch.float32, group.params)
        group.grads = alloc("grads", n, torch.float32, group.grads)
        group.exp_avg = alloc("exp_avg", n, torch.float32, group.exp_avg)          # symmetric only so that a checkpoint can
        group.exp_avg_sq = alloc("exp_avg_sq", n, torch.float32, group.exp_avg_sq)  # collect the shards (full_moments)
        self.sync = alloc("sync", 4 * W, torch.int64)
        lo, hi = shard_bounds(n, W, r)
        self.gsum = torch.zeros(max(4, (n // 4 + W - 1) // W * 4), device=dev)
        self.local = torch.zeros(8, dtype=torch.int32, device=dev)
        hg, hp, hs = self._handles["grads"], self._handles["params"], self._handles["sync"]
        mc_g, mc_p = int(getattr(hg, "multicast_ptr", 0) or 0), int(getattr(hp, "multicast_ptr", 0) or 0)
        self.multicast = bool(mc_g and mc_p) and os.environ.get("B200GYM_DIST_MULTICAST", "1") != "0"
        ptrs = lambda h: (C.c_void_p * W)(*[int(h.buffer_ptrs[p]) for p in range(W)])
        self._peer_arrays = (ptrs(hg), ptrs(hp), ptrs(hs))            # kept alive: the struct only points at them
        a = self.args = _lib.DistAdamArgs()
        a.grads, a.params = group.grads.data_ptr(), group.params.data_ptr()
        a.grads_peer, a.params_peer, a.sync_peer = self._peer_arrays
        a.grads_mc, a.params_mc = (mc_g, mc_p) if self.multicast else (None, None)
        a.exp_avg, a.exp_avg_sq, a.gsum = group.exp_avg.data_ptr(), group.exp_avg_sq.data_ptr(), self.gsum.data_ptr()
        a.state, a.local = group.state.data_ptr(), self.local.data_ptr()
        a.n, a.world, a.rank = n, W, r
        a.max_norm, a.beta1, a.beta2, a.eps = max_grad_norm, betas[0], betas[1], eps
        torch.cuda.synchronize(dev)
        dist.barrier(group=process_group)                              # every rank's buffers are initialised before anyone steps

    def step(self):
        import ctypes as C

        from . import _lib
        self.args.state = self.group.state.data_ptr()
        _lib.check(self.lib.b200_dist_adam(C.byref(self.args), _lib.stream_ptr()))

    def full_moments(self):
        """(exp_avg, exp_avg_sq) with every rank's shard in place, read over the peer mappings -- for checkpoints; call it
        at an iteration boundary, behind a barrier"""
        n, W = self.group.n, self.world
        out = []
        for name in ("exp_avg", "exp_avg_sq"):
            full = torch.empty(n, device=self.group.params.device)
            for p in range(W):
                lo, hi = shard_bounds(n, W, p)
                full[lo:hi] = self._handles[name].get_buffer(p, (n,), torch.float32)[lo:hi]
            out.append(full)
        return out


def enable_fused_dist_adam(groups, process_group, max_grad_norm):
    """-> description of the exchange that will run.  Symmetric memory needs peer access between all GPUs of the group
    (one NVLink / NVSwitch box); where PyTorch cannot set it up the NCCL all-reduce + local clip + Adam path stays."""
    if process_group is None:
        return "single process"
    if os.environ.get("B200GYM_FUSED_DIST", "1") == "0":
        return "NCCL all-reduce + local clip + Adam (fused kernel disabled by B200GYM_FUSED_DIST=0)"
    try:
        fused = [FusedDistAdam(g, process_group, max_grad_norm) for g in groups]
    except Exception as e:                                              # collective: fails (or not) on every rank alike
        return f"NCCL all-reduce + local clip + Adam (symmetric memory unavailable: {type(e).__name__}: {e})"
    for g, f in zip(groups, fused):
        g.dist = f
    how = "multimem.ld_reduce / multimem.st through the NVSwitch" if fused[0].multicast else "peer loads / stores over NVLink"
    return f"fused reduce-scatter + clip + Adam + parameter all-gather kernel ({how}), no NCCL call per step"
