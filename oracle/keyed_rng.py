"""Draw-site context that reroutes the reference's RNG calls to the keyed Philox function.

TEST INFRASTRUCTURE. Used by oracle/ref_runner.py while the unmodified reference runs:
the runner wraps the reference methods that draw random numbers (SURVEY.md §8(c) list)
in `site(...)` blocks that publish (site id, global step, env ids); `draw()` then serves
`torch_rand_float`, `torch.rand`, `torch.rand_like` and `torch.randint_like` from
oracle.philox.keyed_uniform with consecutive lanes, instead of torch's global stream.
"""
import contextlib

import numpy as np
import torch

from . import philox

_ctx = None          # dict(site, step, env_ids (np int64), lane, seed)
SEED = 1234


class _Patch:
    """Swap torch.rand / rand_like / randint_like for keyed versions while a site is active."""

    def __enter__(self):
        self.saved = (torch.rand, torch.rand_like, torch.randint_like)

        def rand(*shape, **kw):
            if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
                shape = tuple(shape[0])
            u = draw(shape, kw.get("device", "cpu"))
            return u if u is not None else self.saved[0](*shape, **kw)

        def rand_like(t, **kw):
            u = draw(tuple(t.shape), t.device)
            return u if u is not None else self.saved[1](t, **kw)

        def randint_like(t, *a, **kw):
            high = a[-1] if a else kw["high"]
            r = draw_u32(tuple(t.shape))
            if r is None:
                return self.saved[2](t, *a, **kw)
            return torch.from_numpy((r % np.uint32(int(high))).astype(np.int64)).to(t.dtype)

        torch.rand, torch.rand_like, torch.randint_like = rand, rand_like, randint_like
        return self

    def __exit__(self, *exc):
        torch.rand, torch.rand_like, torch.randint_like = self.saved


@contextlib.contextmanager
def site(site_id, step, env_ids):
    global _ctx
    prev = _ctx
    _ctx = dict(site=site_id, step=int(step), env_ids=np.asarray(env_ids, dtype=np.int64).reshape(-1), lane=0)
    try:
        with _Patch():
            yield
    finally:
        _ctx = prev


def _lanes(shape):
    """Map a draw of `shape` onto (env rows, consecutive lanes) of the active site."""
    c = _ctx
    n = len(c["env_ids"])
    shape = tuple(int(s) for s in shape)
    if len(shape) == 1:
        assert shape[0] == n, (shape, n)
        width = 1
    else:
        assert shape[0] == n and len(shape) == 2, (shape, n)
        width = shape[1]
    lanes = np.arange(c["lane"], c["lane"] + width)
    if c["site"] == philox.SITE_OBS_NOISE:
        lanes = philox.noise_lane(lanes)
    c["lane"] += width
    return lanes, shape


def draw(shape, device="cpu"):
    if _ctx is None:
        return None
    lanes, shape = _lanes(shape)
    u = philox.keyed_uniform(SEED, _ctx["site"], _ctx["step"], _ctx["env_ids"], lanes)
    return torch.from_numpy(u.reshape(shape).copy())


def draw_u32(shape):
    if _ctx is None:
        return None
    lanes, shape = _lanes(shape)
    return philox.keyed_u32(SEED, _ctx["site"], _ctx["step"], _ctx["env_ids"], lanes).reshape(shape)
