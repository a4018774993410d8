"""The N>1 host logic on CPU: world_size-2 gloo run of the gradient exchange (dist.allreduce_flat_grads) followed by the
clip + Adam arithmetic (oracle) with grad_scale = 1/world -- replicas must end bit-identical and equal to a single-process
run on the averaged gradient."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from legged_gym_custom_b200 import dist as bdist
from legged_gym_custom_b200.networks import FlatGroup
from oracle import learner_oracle as lo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_group(seed):
    g = FlatGroup()
    g.add("a.weight", 7, 13)
    g.add("a.bias", 1, 7)
    g.add("std", 1, 12)
    g.finalize("cpu", 2e-4)
    g.params.copy_(torch.randn(g.n, generator=torch.Generator().manual_seed(seed)))
    return g


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    pg = bdist.init_process_group("gloo")
    assert bdist.env_rank_info() == (rank, world, rank) and bdist.shard_seed(1234, rank) == 1234 + 7919 * rank
    g = _make_group(seed=100 + rank)                     # different start on purpose
    bdist.broadcast_parameters([g], pg)                  # -> rank 0's parameters everywhere
    for step in range(3):
        g.grads.copy_(torch.randn(g.n, generator=torch.Generator().manual_seed(1000 * step + rank)))
        bdist.allreduce_flat_grads([g], pg)
        p = [g.params]
        state = dict(step=step, m=[g.exp_avg], v=[g.exp_avg_sq])
        lo.clip_and_adam(p, [g.grads / world], state, 2e-4, 1.0)
    gathered = [torch.zeros_like(g.params) for _ in range(world)]
    dist.all_gather(gathered, g.params, group=pg)
    if rank == 0:
        out.put([t.clone() for t in gathered])
    dist.barrier()
    dist.destroy_process_group()


def _kl_worker(rank, world, port, out):
    """schedule='adaptive' across ranks (learner.PPO._adaptive_lr): every rank adds its minibatch's KL SUM, the sums are
    all-reduced, and the rule is applied to sum / (M * world) -- so all replicas move their learning rate together"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    pg = bdist.init_process_group("gloo")
    lr, log = 2e-4, []
    for step in range(4):
        mu, old_mu, sigma, old_sigma = _kl_batch(step, rank)
        kl = torch.sum(torch.log(sigma / old_sigma + 1.e-5) + (old_sigma ** 2 + (old_mu - mu) ** 2) / (2.0 * sigma ** 2) - 0.5, dim=-1)
        acc = kl.double().sum().reshape(1)                          # kl_sum_kernel's accumulator
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=pg)
        kl_mean = (acc / (mu.shape[0] * world)).float()             # adaptive_lr_kernel
        if kl_mean > 0.01 * 2.0:
            lr = max(1e-5, lr / 1.5)
        elif kl_mean < 0.01 / 2.0 and kl_mean > 0.0:
            lr = min(1e-2, lr * 1.5)
        log.append(lr)
    out.put((rank, log))
    dist.barrier()
    dist.destroy_process_group()


def _kl_batch(step, rank, M=64, A=12):
    g = torch.Generator().manual_seed(100 * step + rank)
    spread = [0.0, 0.02, 0.5, 0.5][step] * (1 + rank)               # rank 1 always sees the larger policy change
    mu = torch.randn(M, A, generator=g)
    return mu, mu + spread * torch.randn(M, A, generator=g), torch.ones(M, A), torch.ones(M, A)


def test_two_rank_adaptive_kl_decision_is_global():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_kl_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    logs = dict(out.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert logs[0] == logs[1]
    # single-process oracle on the concatenated minibatches (oracle.learner_oracle.adaptive_lr = ppo.py:233-246)
    lr, want = 2e-4, []
    for step in range(4):
        b = [_kl_batch(step, r) for r in range(world)]
        mu, old_mu, sigma, old_sigma = (torch.cat([x[i] for x in b]) for i in range(4))
        lr, _ = lo.adaptive_lr(lr, mu, sigma, old_mu, old_sigma, 0.01)
        want.append(lr)
    assert logs[0] == want and len(set(want)) >= 3                  # grows, holds / shrinks: more than one branch taken


def test_two_rank_gradient_exchange_keeps_replicas_identical():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(gathered[0], gathered[1])
    # single-process reference on the averaged gradients
    g = _make_group(seed=100)
    for step in range(3):
        grads = sum(torch.randn(g.n, generator=torch.Generator().manual_seed(1000 * step + r)) for r in range(world))
        lo.clip_and_adam([g.params], [grads / world], dict(step=step, m=[g.exp_avg], v=[g.exp_avg_sq]), 2e-4, 1.0)
    assert torch.equal(g.params, gathered[0])
