"""Deploy wire format: the TorchScript files written by legged_gym_custom_b200.export against the
reference's own export (golden fixture tests/golden/deploy_small.npz, made by oracle/make_golden_deploy.py), and the
batched GPU network stage of the deploy controller against the same fixture."""
import os
import types

import numpy as np
import pytest
import torch

from legged_gym_custom_b200 import export

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "deploy_small.npz")
SHIPPED = "/root/reference/deploy/networks/go2"


def _gold():
    z = np.load(GOLD)
    ac = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("ac/")}
    est = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("est/")}
    return z, ac, est


def _holder(sd, **attrs):
    return types.SimpleNamespace(state_dict=lambda: sd, **attrs)


def _chain(mods, obs, scan, clip_obs, clip_act):
    """deploy_base.py:241-266 on loaded TorchScript modules"""
    with torch.no_grad():
        obs_c = torch.clip(obs, -clip_obs, clip_obs)
        latent = mods["adaptation"](obs_c[:, :520].reshape(-1, 10, 52))
        est = mods["estimator"](obs_c)
        scan_latent = mods["scan_encoder"](scan)
        raw = mods["policy"](torch.cat((obs_c, latent, scan_latent, est), dim=-1))
    return dict(latent=latent, est=est, scan_latent=scan_latent, raw_actions=raw, actions=torch.clip(raw, -clip_act, clip_act))


def test_export_files_match_reference_export(tmp_path):
    z, ac, est = _gold()
    paths = export.export_policy_as_jit(_holder(ac), _holder(est, num_proprio=52, use_history=True), str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == sorted(export.FILES.values())
    mods = {name: torch.jit.load(p) for name, p in paths.items()}
    # same module tree: identical state_dict keys in identical order, as the reference's files have them
    for name, m in mods.items():
        assert list(m.state_dict().keys()) == [str(k) for k in z["keys/" + name]], name
    assert mods["policy"].original_name == "Sequential" and mods["adaptation"].original_name == "AdaptationEncoder"
    assert mods["estimator"].original_name == "MlpEstimator" and mods["scan_encoder"].original_name == "ScanEncoder"
    clip_obs, clip_act = (float(v) for v in z["meta/clip"])
    got = _chain(mods, torch.from_numpy(z["in/obs"]), torch.from_numpy(z["in/scan"]), clip_obs, clip_act)
    for k, v in got.items():
        # same ATen CPU ops on the same weights: 1e-6 absolute leaves room for a different BLAS blocking only
        np.testing.assert_allclose(v.numpy(), z["out/" + k], rtol=0, atol=2e-6, err_msg=k)


def test_load_deploy_networks_round_trip(tmp_path):
    _, ac, est = _gold()
    export.export_policy_as_jit(_holder(ac), _holder(est), str(tmp_path))
    ac2, est2 = export.load_deploy_networks(str(tmp_path))
    wanted = [k for k in ac if k.startswith(("actor.", "adaptation_encoder_.", "scan_encoder."))]
    assert list(ac2.keys()) == wanted                       # checkpoint key names and order; critic / priv encoder / std stay behind
    for k in wanted:
        assert torch.equal(ac2[k], ac[k]), k
    assert list(est2.keys()) == list(est.keys()) and all(torch.equal(est2[k], est[k]) for k in est)


def test_estimator_without_history_reads_last_proprio(tmp_path):
    g = torch.Generator().manual_seed(3)
    est = {"estimator.0.weight": torch.randn(8, 52, generator=g), "estimator.0.bias": torch.randn(8, generator=g),
           "estimator.2.weight": torch.randn(3, 8, generator=g), "estimator.2.bias": torch.randn(3, generator=g)}
    _, ac, _ = _gold()
    paths = export.export_policy_as_jit(_holder(ac), _holder(est, num_proprio=52, use_history=False), str(tmp_path))
    m = torch.jit.load(paths["estimator"])
    obs = torch.randn(5, 572, generator=g)
    ref = torch.nn.functional.elu(obs[:, -52:] @ est["estimator.0.weight"].t() + est["estimator.0.bias"]) @ est["estimator.2.weight"].t() \
        + est["estimator.2.bias"]
    assert torch.allclose(m(obs), ref, atol=1e-6)


@pytest.mark.skipif(not os.path.isdir(SHIPPED), reason="reference tree absent (GPU box)")
@pytest.mark.parametrize("model", ["parkour_v12_ft_i", "parkour_v12_ft_iii"])
def test_shipped_deploy_networks_re_export_bit_identical(tmp_path, model):
    """the trained networks the reference ships: read -> checkpoint keys -> written again -> same outputs, bit for bit"""
    src = os.path.join(SHIPPED, model)
    ac, est = export.load_deploy_networks(src)
    assert ac["actor.0.weight"].shape == (512, 627) and est["estimator.0.weight"].shape[1] == 572
    paths = export.export_policy_as_jit(_holder(ac), _holder(est), str(tmp_path))
    g = torch.Generator().manual_seed(1)
    obs, scan = torch.randn(16, 572, generator=g), torch.randn(16, 132, generator=g) * 0.3
    a = _chain({n: torch.jit.load(os.path.join(src, f)) for n, f in export.FILES.items()}, obs, scan, 100.0, 3.14)
    b = _chain({n: torch.jit.load(p) for n, p in paths.items()}, obs, scan, 100.0, 3.14)
    for k in a:
        assert torch.equal(a[k], b[k]), k


@pytest.mark.skipif(not os.path.isdir(SHIPPED), reason="reference tree absent (GPU box)")
def test_pre_parkour_model_directory_has_no_scan_encoder():
    ac, est = export.load_deploy_networks(os.path.join(SHIPPED, "cheetah_v8"))
    assert not any(k.startswith("scan_encoder.") for k in ac) and ac["actor.0.weight"].shape == (512, 595)


@pytest.mark.gpu
@pytest.mark.parametrize("precise", [True, False])
def test_deploy_policy_gpu_matches_reference_chain(tmp_path, precise):
    """DeployPolicy (sm_100a kernels) on the golden export: 3xTF32 mode to 1e-4 of the output scale, TF32 mode (the
    reference's GPU matmul precision, scripts/train.py:39) to 5e-2 -- the fixture's observations reach the +-100 clip, and
    truncating those to 10 mantissa bits alone costs 1.8e-2 of the output scale (emulated on the CPU); the action clip is
    exact where the reference clipped"""
    z, ac, est = _gold()
    export.export_policy_as_jit(_holder(ac), _holder(est), str(tmp_path))
    clip_obs, clip_act = (float(v) for v in z["meta/clip"])
    pol = export.DeployPolicy(str(tmp_path), device="cuda:0", clip_obs=clip_obs, clip_actions=clip_act, precise=precise)
    obs, scan = torch.from_numpy(z["in/obs"]).cuda(), torch.from_numpy(z["in/scan"]).cuda()
    actions = pol(obs, scan).cpu().numpy()
    raw_ref, ref = z["out/raw_actions"], z["out/actions"]
    scale = float(np.sqrt((raw_ref ** 2).mean()))
    tol = (1e-4 if precise else 5e-2) * scale
    assert np.abs(actions - ref).max() <= tol, (np.abs(actions - ref).max(), tol)
    far = np.abs(raw_ref) > clip_act + 10 * tol              # clipped by a margin no rounding can cross
    assert far.any() and np.array_equal(actions[far], ref[far])
    # intermediate stages through the public methods
    lat = pol.actor_critic.adaptation_encoder(torch.clamp(obs, -clip_obs, clip_obs)).cpu().numpy()
    assert np.abs(lat - z["out/latent"]).max() <= 1e-4 * max(1.0, float(np.abs(z["out/latent"]).max()))
    e = pol.estimator(torch.clamp(obs, -clip_obs, clip_obs)).cpu().numpy()
    es = float(np.sqrt((z["out/est"] ** 2).mean()))
    assert np.abs(e - z["out/est"]).max() <= (1e-4 if precise else 5e-2) * es
