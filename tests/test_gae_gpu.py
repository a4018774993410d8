"""GAE + advantage normalisation + time-out bootstrap kernels vs the CPU oracle
(rollout_storage.py:110-124, ppo.py:160-166) through the C ABI."""
import ctypes as C

import numpy as np
import pytest
import torch

import golden_util as gu
from legged_gym_custom_b200 import _lib
from oracle import learner_oracle as lo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run_gae(rewards, dones, values, last_values, gamma, lam):
    lib = _lib.lib()
    T, N = rewards.shape[:2]
    d = lambda t: t.to(DEV).contiguous()
    r, dn, v, lv = d(rewards), d(dones), d(values), d(last_values)
    ret, adv = torch.empty_like(r), torch.empty_like(r)
    scratch = torch.empty(int(lib.b200_gae_scratch_bytes(T, N)), dtype=torch.uint8, device=DEV)
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(lib.b200_compute_returns(p(r), p(dn), p(v), p(lv), p(ret), p(adv), T, N, gamma, lam, p(scratch), _lib.stream_ptr()))
    torch.cuda.synchronize()
    return ret.cpu(), adv.cpu()


@pytest.mark.parametrize("T,N", [(24, 4096), (24, 1), (24, 333), (7, 65536), (40, 100)])
def test_gae_matches_oracle(T, N):
    g = torch.Generator().manual_seed(T * 1000 + N)
    rewards = torch.rand(T, N, 1, generator=g) * 0.05
    values = torch.randn(T, N, 1, generator=g)
    dones = (torch.rand(T, N, 1, generator=g) < 0.02).byte()
    last_values = torch.randn(N, 1, generator=g)
    ret_ref, adv_ref = lo.compute_returns(rewards, dones, values, last_values, 0.99, 0.95)
    ret, adv = _run_gae(rewards, dones, values, last_values, 0.99, 0.95)
    assert torch.equal(ret, ret_ref), f"returns differ: {gu.rel_err(ret.numpy(), ret_ref.numpy()):.2e}"   # same op order, no FMA
    assert gu.rel_err(adv.numpy(), adv_ref.numpy()) <= 1e-5


def test_gae_all_done_and_constant():
    """edge cases: every step terminal (no bootstrapping), and zero variance handled like torch (inf/nan-free check skipped)."""
    T, N = 24, 256
    rewards, values = torch.ones(T, N, 1), torch.zeros(T, N, 1)
    dones = torch.ones(T, N, 1, dtype=torch.uint8)
    ret, _ = _run_gae(rewards, dones, values + torch.randn(T, N, 1), torch.randn(N, 1), 0.99, 0.95)
    ret_ref, _ = lo.compute_returns(rewards, dones, values, torch.zeros(N, 1), 0.99, 0.95)
    assert torch.equal(ret_ref, torch.ones(T, N, 1))
    assert torch.allclose(ret, torch.ones(T, N, 1), atol=1e-6)   # delta = r - v, return = delta + v


@pytest.mark.parametrize("with_timeouts", [True, False])
def test_store_step_scalars(with_timeouts):
    lib = _lib.lib()
    N = 4097
    g = torch.Generator().manual_seed(1)
    rew, values = torch.rand(N, generator=g), torch.randn(N, 1, generator=g)
    reset = torch.rand(N, generator=g) < 0.1
    tmo = (torch.rand(N, generator=g) < 0.05) & reset
    ref = lo.bootstrap_rewards(rew, values, tmo if with_timeouts else None, 0.99)
    d = lambda t: t.to(DEV).contiguous()
    rew_d, val_d, reset_d, tmo_d = d(rew), d(values), d(reset), d(tmo)
    out_r, out_d = torch.empty(N, device=DEV), torch.empty(N, dtype=torch.uint8, device=DEV)
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(lib.b200_store_step_scalars(p(rew_d), p(reset_d), p(tmo_d) if with_timeouts else None, p(val_d), 0.99,
                                           p(out_r), p(out_d), N, _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(out_r.cpu(), ref)
    assert torch.equal(out_d.cpu().bool(), reset)
