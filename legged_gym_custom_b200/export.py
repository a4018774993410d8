"""Deploy-side wire format of the trained networks.

The reference hands its policy to the robot as four TorchScript files written by
`export_policy_as_jit` (legged_gym/utils/helpers.py:180-214) and read back by the deploy controller
(deploy/base/deploy_base.py:33-36): `policy.pt` (the actor `nn.Sequential`, keys `0.weight` ...),
`adaptation_module.pt` (AdaptationEncoder, support_networks.py:128-175), `estimator.pt` (MlpEstimator,
support_networks.py:44-93) and `scan_encoder.pt` (ScanEncoder, support_networks.py:9-41).  This module writes
exactly those files from the kernel-backed `ActorCritic` / `MlpEstimator` (flat device buffers -> reference
`state_dict` -> TorchScript), reads them back (`load_deploy_networks`; also how BASELINE config 4 can start from
the trained `deploy/networks/go2/parkour_v12_ft_*` weights), and runs the controller's network stage batched on
the GPU (`DeployPolicy`, deploy_base.py:241-266) through the same C-ABI kernels as `ActorCritic.act_inference`.

The `torch.nn` containers below exist only as the serialisation schema TorchScript needs (module tree, key names,
forward signatures); nothing in the product path evaluates them.
"""
import os
from collections import OrderedDict

import torch
from torch import nn

FILES = OrderedDict(policy="policy.pt", adaptation="adaptation_module.pt", estimator="estimator.pt", scan_encoder="scan_encoder.pt")
_ADAPT_PREFIX, _SCAN_PREFIX = "adaptation_encoder_.", "scan_encoder."


# ---- serialisation schema ----------------------------------------------------------------------------------
def _elu_mlp(prefix, sd):
    """nn.Sequential(Linear, ELU, Linear, ELU, ..., Linear) with the layer shapes found under `prefix` (keys
    `<prefix>0.weight`, `<prefix>2.weight`, ...: the reference numbers Linear layers 0, 2, 4, ...)."""
    layers, i = [], 0
    while f"{prefix}{i}.weight" in sd:
        out_f, in_f = sd[f"{prefix}{i}.weight"].shape
        if layers:
            layers.append(nn.ELU())
        layers.append(nn.Linear(in_f, out_f))
        i += 2
    if not layers:
        raise KeyError(f"no '{prefix}<n>.weight' entries in state_dict")
    seq = nn.Sequential(*layers)
    seq.load_state_dict({k[len(prefix):]: torch.as_tensor(v).detach().cpu().float() for k, v in sd.items()
                         if k.startswith(prefix) and k[len(prefix):].split(".")[0].isdigit()})
    return seq


class AdaptationEncoder(nn.Module):
    """history [B, H, P] -> Linear(P, 30)+ELU per step -> Conv1d(30, 20, 4, 2)+ELU -> Conv1d(20, 10, 2, 1)+ELU ->
    Flatten -> Linear(30, out)+ELU; module names follow support_networks.py:147-166 so the keys match."""

    def __init__(self, num_proprio, output_dim):
        super().__init__()
        act = nn.ELU()
        self.fc_encoder = nn.Sequential(nn.Linear(num_proprio, 30), act)
        self.conv_layers = nn.Sequential(nn.Conv1d(30, 20, kernel_size=4, stride=2), act,
                                         nn.Conv1d(20, 10, kernel_size=2, stride=1), act, nn.Flatten())
        self.fc_final = nn.Sequential(nn.Linear(30, output_dim), act)

    def forward(self, unflattened_obs_history):
        steps = self.fc_encoder(unflattened_obs_history)
        return self.fc_final(self.conv_layers(steps.permute(0, 2, 1)))


class ScanEncoder(nn.Module):
    def __init__(self, seq):
        super().__init__()
        self.scan_encoder = seq

    def forward(self, scan_obs):
        return self.scan_encoder(scan_obs)


class MlpEstimator(nn.Module):
    """use_history=False estimators read only the last `num_proprio` columns (support_networks.py:86-93)."""

    def __init__(self, seq, num_proprio: int, use_history: bool):
        super().__init__()
        self.estimator = seq
        self.num_proprio = num_proprio
        self.use_history = use_history

    def forward(self, obs_with_history):
        if self.use_history:
            return self.estimator(obs_with_history)
        return self.estimator(obs_with_history[:, -self.num_proprio:])


def build_export_modules(actor_critic_sd, estimator_sd, num_proprio=52, use_history=True):
    """reference-keyed state_dicts -> {'policy', 'adaptation', 'estimator', 'scan_encoder'} CPU modules."""
    ad = {k[len(_ADAPT_PREFIX):]: torch.as_tensor(v).detach().cpu().float() for k, v in actor_critic_sd.items() if k.startswith(_ADAPT_PREFIX)}
    adaptation = AdaptationEncoder(ad["fc_encoder.0.weight"].shape[1], ad["fc_final.0.weight"].shape[0])
    adaptation.load_state_dict(ad)
    return OrderedDict(policy=_elu_mlp("actor.", actor_critic_sd), adaptation=adaptation,
                       estimator=MlpEstimator(_elu_mlp("estimator.", estimator_sd), num_proprio, use_history),
                       scan_encoder=ScanEncoder(_elu_mlp(_SCAN_PREFIX + "scan_encoder.", actor_critic_sd)))


def export_policy_as_jit(actor_critic, estimator, path):
    """helpers.py:180-214 for the kernel-backed networks: writes policy.pt, adaptation_module.pt, estimator.pt and
    scan_encoder.pt under `path` and returns their paths.  `actor_critic` / `estimator` are anything with a
    reference-keyed `state_dict()`."""
    os.makedirs(path, exist_ok=True)
    mods = build_export_modules(actor_critic.state_dict(), estimator.state_dict(),
                                num_proprio=getattr(estimator, "num_proprio", 52), use_history=getattr(estimator, "use_history", True))
    out = OrderedDict()
    for name, mod in mods.items():
        out[name] = os.path.join(path, FILES[name])
        torch.jit.script(mod.eval()).save(out[name])
    return out


# ---- reading the deploy files back ---------------------------------------------------------------------------
def load_deploy_networks(path):
    """The four TorchScript files of a deploy model directory -> (actor_critic_state_dict, estimator_state_dict) in the
    reference's checkpoint key names.  The actor-critic part is partial (actor, adaptation encoder, scan encoder: what
    the robot needs); load it with `ActorCritic.load_state_dict(sd, strict=False)`.  Older model directories
    (deploy/networks/go2/cheetah_v8*) have no scan_encoder.pt and are returned without those keys."""
    ac, est = OrderedDict(), OrderedDict()
    for name, prefix, dst in (("policy", "actor.", ac), ("adaptation", _ADAPT_PREFIX, ac), ("scan_encoder", _SCAN_PREFIX, ac),
                              ("estimator", "", est)):
        f = os.path.join(path, FILES[name])
        if not os.path.exists(f):
            if name == "scan_encoder":
                continue
            raise FileNotFoundError(f)
        for k, v in torch.jit.load(f, map_location="cpu").state_dict().items():
            dst[prefix + k] = v.detach().clone()
    return ac, est


def _hidden_dims(sd, prefix):
    dims, i = [], 0
    while f"{prefix}{i}.weight" in sd:
        dims.append(int(sd[f"{prefix}{i}.weight"].shape[0]))
        i += 2
    return dims[:-1], dims[-1]


class DeployPolicy:
    """The network stage of the deploy controller's step (deploy_base.py:241-266), batched on the GPU:
    clip(obs) -> adaptation latent (history part), estimator, scan encoder -> actor([obs | latent | scan latent | est])
    -> clip(actions).  Built from a deploy model directory; every product of it runs in the sm_100a kernels."""

    def __init__(self, path, device="cuda:0", clip_obs=100.0, clip_actions=3.14, precise=False):
        from .networks import ActorCritic, MlpEstimator as KernelEstimator
        ac_sd, est_sd = load_deploy_networks(path)
        if _SCAN_PREFIX + "scan_encoder.0.weight" not in ac_sd:
            raise NotImplementedError("model directory without scan_encoder.pt (pre-parkour network layout)")
        num_proprio = int(ac_sd[_ADAPT_PREFIX + "fc_encoder.0.weight"].shape[1])
        actor_hidden, num_actions = _hidden_dims(ac_sd, "actor.")
        scan_hidden, scan_out = _hidden_dims(ac_sd, _SCAN_PREFIX + "scan_encoder.")
        est_hidden, num_est = _hidden_dims(est_sd, "estimator.")
        latent = int(ac_sd[_ADAPT_PREFIX + "fc_final.0.weight"].shape[0])
        num_scan = int(ac_sd[_SCAN_PREFIX + "scan_encoder.0.weight"].shape[1])
        actor_in = int(ac_sd["actor.0.weight"].shape[1])
        num_obs = actor_in - latent - scan_out - num_est
        history, rem = divmod(num_obs, num_proprio)
        assert rem == 0, (num_obs, num_proprio)
        est_in = int(est_sd["estimator.0.weight"].shape[1])
        self.clip_obs, self.clip_actions, self.num_obs, self.num_scan_obs = float(clip_obs), float(clip_actions), num_obs, num_scan
        # the privileged encoder and the critic never run on the robot: smallest legal shapes
        self.actor_critic = ActorCritic(num_proprio=num_proprio, num_privileged_obs=4, num_critic_obs=4, num_estimated_obs=num_est,
                                        num_scan_obs=num_scan, num_actions=num_actions, history_buffer_length=history - 1,
                                        actor_hidden_dims=actor_hidden, critic_hidden_dims=[4], priv_encoder_hidden_dims=[4],
                                        scan_encoder_hidden_dims=scan_hidden, latent_encoder_output_dim=latent,
                                        scan_encoder_output_dim=scan_out, device=device, precise=precise)
        self.actor_critic.load_state_dict(ac_sd, strict=False)
        self.estimator = KernelEstimator(num_proprio=num_proprio, history_buffer_length=history - 1, output_dim=num_est,
                                         hidden_dims=est_hidden, use_history=(est_in == num_obs), device=device, precise=precise)
        self.estimator.load_state_dict(est_sd)

    def __call__(self, obs, scan_obs):
        obs = torch.clamp(obs, -self.clip_obs, self.clip_obs)
        est = self.estimator(obs)
        actions = self.actor_critic.act_inference(obs, None, est, scan_obs, adaptation_mode=True)
        return torch.clamp(actions, -self.clip_actions, self.clip_actions)
