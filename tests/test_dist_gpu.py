"""The fused data-parallel optimiser step (csrc/dist_adam.cu, b200_dist_adam) on 2 GPUs of one box: reduce-scatter of the
gradients through peer / multicast memory + global-norm clip + Adam on the owner's shard + parameter all-gather, against the
single-GPU kernels (sum of the two gradients -> b200_clip_adam with grad_scale 1/2).  Needs >= 2 GPUs (skipped otherwise:
the driver's single-GPU run does not see it; run under `gpurun --gpus 2`)."""
import ctypes as C
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, steps, multicast, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      B200GYM_DIST_MULTICAST=str(int(multicast)))
    import torch.distributed as dist
    from legged_gym_custom_b200 import _lib
    from legged_gym_custom_b200.dist import FusedDistAdam, shard_bounds
    from legged_gym_custom_b200.networks import FlatGroup
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", device_id=dev)
    try:
        lib = _lib.lib()
        g = FlatGroup()
        g.add("w.weight", n // 8, 8)
        g.finalize(dev, 2e-4)
        gen = torch.Generator().manual_seed(5)
        p0 = torch.randn(g.n, generator=gen)
        g.params.copy_(p0)
        fused = FusedDistAdam(g, dist.group.WORLD, max_grad_norm=1.0)
        assert fused.multicast == bool(multicast) or not multicast
        # reference: one GPU, gradients summed on the host, the existing local kernels
        ref = FlatGroup()
        ref.add("w.weight", n // 8, 8)
        ref.finalize(dev, 2e-4)
        ref.params.copy_(p0)
        if rank == 0:
            ref.state[7] = 0.37                      # a stale squared norm in the clip (ppo.py:274 quirk): both paths must carry it
        g.state[7] = 0.37 if rank == 0 else 0.0
        dist.broadcast(g.state, src=0)
        dist.broadcast(ref.state, src=0)
        for s in range(steps):
            grads = [torch.randn(g.n, generator=torch.Generator().manual_seed(100 * s + r)) * (0.02 if s % 2 else 3.0) for r in range(world)]
            g.grads.copy_(grads[rank])
            fused.step()
            ref.grads.copy_(sum(grads))
            _lib.check(lib.b200_clip_adam(ref.params.data_ptr(), ref.grads.data_ptr(), ref.exp_avg.data_ptr(), ref.exp_avg_sq.data_ptr(), ref.n,
                                          C.c_void_p(ref.state.data_ptr()), 1.0 / world, 1.0, 0.9, 0.999, 1e-8, _lib.stream_ptr()))
            torch.cuda.synchronize()
            dist.barrier()
            assert float(g.grads.abs().max()) == 0.0                                   # every shard zeroed on every rank
            rel = float((g.params - ref.params).abs().max() / ref.params.abs().max())
            assert rel <= 2e-6, (s, rel)
            st, rs = g.state.cpu(), ref.state.cpu()
            assert st[1] == rs[1] == s + 1 and st[2] == rs[2] and st[3] == rs[3] and st[4] == rs[4]
            assert abs(float(st[6] - rs[6])) <= 1e-6 * float(rs[6]) and abs(float(st[7] - rs[7])) <= 1e-6 * max(float(rs[7]), 1e-30)
        # replicas are bit-identical; moments are sharded and collectable
        mine = g.params.clone()
        other = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(other, mine)
        assert all(torch.equal(o, other[0]) for o in other)
        m, v = fused.full_moments()
        assert float((m - ref.exp_avg).abs().max()) <= 1e-6 * float(ref.exp_avg.abs().max())
        assert float((v - ref.exp_avg_sq).abs().max()) <= 1e-6 * float(ref.exp_avg_sq.abs().max())
        lo, hi = shard_bounds(g.n, world, rank)
        assert torch.equal(g.exp_avg[lo:hi], m[lo:hi])
        # graph replay: the barrier epoch and the optimiser state advance on the device
        g.grads.copy_(grads[rank])
        torch.cuda.synchronize(); dist.barrier()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            fused.step()                                  # warm (not captured)
            torch.cuda.synchronize(); dist.barrier()
            with torch.cuda.graph(graph, stream=side):
                fused.step()
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize(); dist.barrier()
        assert int(g.state[1].item()) == steps + 1 + 3
        out[rank] = "ok multicast" if fused.multicast else "ok peer"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("multicast", [1, 0], ids=["nvswitch-multicast-if-available", "peer-loads-and-stores"])
def test_fused_dist_adam_two_gpus(multicast):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), 1 << 18, 4, multicast, out), nprocs=2, join=True)
    assert len(out) == 2 and all(v.startswith("ok") for v in out.values()), dict(out)
    print(dict(out))


def _runner_worker(rank, world, port, schedule, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from legged_gym_custom_b200 import configs
    from legged_gym_custom_b200.env import Go2Env
    from legged_gym_custom_b200.runner import OnPolicyRunner, class_to_dict
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", device_id=dev)
    try:
        env_cfg, train_cfg = configs.TASKS["go2_parkour"]

        class Cfg(env_cfg):
            class env(env_cfg.env):
                num_envs = 256
        env = Go2Env(Cfg, sim_device=str(dev), seed=1234 + rank)          # every rank simulates its own shard of the envs
        tc = class_to_dict(train_cfg)
        tc["runner"]["resume"] = False
        tc["algorithm"]["schedule"] = schedule
        runner = OnPolicyRunner(env, tc, log_dir=None, device=dev, process_group=dist.group.WORLD)
        runner.enable_graphs()
        runner.capture_graphs()
        for it in range(3):                                               # DAgger iteration, then two PPO iterations (graph replays)
            runner.iteration(it)
        torch.cuda.synchronize()
        ac, est = runner.alg.actor_critic, runner.alg.estimator
        flat = torch.cat([ac.main.params, ac.adapt.params, est.group.params,
                          ac.main.state[:5].float(), est.group.state[:5].float(), ac.adapt.state[:5].float()])
        assert torch.isfinite(flat).all()
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        assert all(torch.equal(b, both[0]) for b in both), "replicas diverged"
        # the shards are different data: the rollouts must NOT be identical
        obs = runner.alg.storage.observations[0, :8].clone()
        obs_all = [torch.empty_like(obs) for _ in range(world)]
        dist.all_gather(obs_all, obs)
        assert not torch.equal(obs_all[0], obs_all[1])
        out[rank] = (runner.alg.dist_mode, float(ac.main.state[4].item()), int(ac.main.state[1].item()))
        runner.release_graphs()                 # graphs with captured NCCL work must go before the communicator does
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("schedule", ["fixed", "adaptive"])
def test_two_gpu_runner_keeps_replicas_bit_identical(schedule):
    """the whole learning iteration on 2 ranks (envs sharded, captured graphs, fused optimiser step; with schedule='adaptive'
    the NCCL all-reduce of the KL sums inside every minibatch graph): parameters, Adam step counts and learning rates stay
    bit-identical across the ranks while their rollouts differ"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_runner_worker, args=(2, _free_port(), schedule, out), nprocs=2, join=True)
    assert len(out) == 2 and out[0] == out[1], dict(out)
    mode, lr, steps = out[0]
    assert steps == 2 * 20                                   # two PPO updates of 5 epochs x 4 minibatches
    if schedule == "adaptive":
        assert lr != 2e-4                                    # the KL rule moved the learning rate (and identically on both ranks)
    print(dict(out))
