"""Synthetic rollout storage + minibatch dicts shared by the learner tests."""
import torch

DIMS = dict(obs=572, priv=29, critic=736, est=3, scan=132, act=12)


def random_storage(T, N, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, scale=1.0: torch.randn(*s, generator=g) * scale
    obs = r(T, N, DIMS["obs"], scale=0.5)
    priv, est, scan = r(T, N, DIMS["priv"], scale=0.3), r(T, N, DIMS["est"]), torch.clamp(r(T, N, DIMS["scan"], scale=0.5), -1, 1)
    return dict(obs=obs, priv=priv, true_est=est, scan=scan, critic_obs=torch.cat([obs, priv, est, scan], -1),
                actions=r(T, N, DIMS["act"]), values=r(T, N, 1), returns=r(T, N, 1), adv=r(T, N, 1),
                old_logp=r(T, N, 1, scale=2.0) - 17.0, mu=r(T, N, DIMS["act"]), sigma=torch.ones(T, N, DIMS["act"]))


def minibatch(st, idx):
    return {k: v.flatten(0, 1)[idx] for k, v in st.items()}
