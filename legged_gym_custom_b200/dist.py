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
