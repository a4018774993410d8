"""Generate tests/golden/deploy_small.npz: the reference's own TorchScript export + deploy network chain (CPU).

TEST INFRASTRUCTURE; authoring container only.  `python -m oracle.make_golden_deploy`

The reference's ActorCritic / MlpEstimator (small hidden sizes, torch-seeded init) are exported with the reference's
`export_policy_as_jit` (legged_gym/utils/helpers.py:180-214); the four saved files are loaded back with
`torch.jit.load` and run in the order of the deploy controller (deploy/base/deploy_base.py:241-266) on seeded inputs,
some of them beyond the +-100 observation clip.  Stored: the state dicts, the inputs, every intermediate and the
clipped actions.
"""
import os
import tempfile

import numpy as np
import torch

from oracle import ref_runner

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "deploy_small.npz")
HID = dict(actor=[64, 32, 16], critic=[16], priv=[16], scan=[32, 16], est=[32, 16])
B, CLIP_OBS, CLIP_ACT = 96, 100.0, 3.14


def main():
    ref_runner._setup_path()
    import isaacgym  # noqa: F401  (the stub; helpers.py imports it)
    from legged_gym.envs import task_registry  # noqa: F401  (import order of the reference's scripts; avoids its circular import)
    from legged_gym.utils.helpers import export_policy_as_jit
    from rsl_rl.modules import ActorCritic
    from rsl_rl.modules.support_networks import MlpEstimator
    torch.manual_seed(23)
    ac = ActorCritic(52, 29, 736, 3, 132, 12, 10, actor_hidden_dims=HID["actor"], critic_hidden_dims=HID["critic"],
                     priv_encoder_hidden_dims=HID["priv"], scan_encoder_hidden_dims=HID["scan"], latent_encoder_output_dim=20,
                     scan_encoder_output_dim=32, activation='elu', init_noise_std=1.0)
    est = MlpEstimator(52, 10, 3, hidden_dims=HID["est"], activation='elu', use_history=True)
    # trained networks emit actions of order 1; scale the head so that the +-3.14 action clip is exercised
    with torch.no_grad():
        ac.actor[-1].weight.mul_(6.0)
    g = torch.Generator().manual_seed(5)
    obs = torch.randn(B, 572, generator=g) * 1.5
    obs[::7, ::13] *= 90.0                                  # beyond clip_observations
    scan = torch.clamp(torch.randn(B, 132, generator=g) * 0.4, -1, 1)
    with tempfile.TemporaryDirectory() as d:
        export_policy_as_jit(ac, est, d)
        policy, adaptation = torch.jit.load(os.path.join(d, "policy.pt")), torch.jit.load(os.path.join(d, "adaptation_module.pt"))
        estimator, scan_encoder = torch.jit.load(os.path.join(d, "estimator.pt")), torch.jit.load(os.path.join(d, "scan_encoder.pt"))
        keys = {f: list(m.state_dict().keys()) for f, m in (("policy", policy), ("adaptation", adaptation), ("estimator", estimator),
                                                            ("scan_encoder", scan_encoder))}
        with torch.no_grad():
            obs_c = torch.clip(obs, -CLIP_OBS, CLIP_OBS)
            latent = adaptation(obs_c[:, :520].reshape(B, 10, 52))
            est_out = estimator(obs_c)
            scan_latent = scan_encoder(scan)
            raw = policy(torch.cat((obs_c, latent, scan_latent, est_out), dim=-1))
            actions = torch.clip(raw, -CLIP_ACT, CLIP_ACT)
    out = {"in/obs": obs.numpy(), "in/scan": scan.numpy(), "out/latent": latent.numpy(), "out/est": est_out.numpy(),
           "out/scan_latent": scan_latent.numpy(), "out/raw_actions": raw.numpy(), "out/actions": actions.numpy(),
           "meta/clip": np.array([CLIP_OBS, CLIP_ACT], np.float32)}
    for k, v in ac.state_dict().items():
        out["ac/" + k] = v.detach().numpy()
    for k, v in est.state_dict().items():
        out["est/" + k] = v.detach().numpy()
    for f, ks in keys.items():
        out["keys/" + f] = np.array(ks)
    np.savez_compressed(OUT, **out)
    frac = float((raw.abs() > CLIP_ACT).float().mean())
    print("wrote", OUT, os.path.getsize(OUT), "bytes; actions clipped: %.1f %%" % (100 * frac))


if __name__ == "__main__":
    main()
