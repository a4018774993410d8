"""Keyed counter-based uniforms shared by the oracle and the CUDA kernels.

TEST INFRASTRUCTURE (oracle/): imported only by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs.

The reference draws its randomness from torch's global Philox stream at nine
data-dependent sites (e.g. legged_robot.py:491, :520, :526,
:539, :572, go2.py:428-456, :519).  A fused kernel cannot reproduce torch's
stream order, so BOTH sides use the same stateless function instead:

    u32 = Philox4x32-10(counter=(env, step, site, lane >> 2), key=(seed_lo, seed_hi))[lane & 3]
    u   = (u32 >> 8) * 2**-24            in [0, 1), exactly representable in fp32

The CUDA twin lives in legged_gym_custom_b200/csrc/philox.cuh.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

# draw sites (must match csrc/philox.cuh)
SITE_CMD_PERIODIC = 0   # go2.py:393-396 -> _resample_commands; lanes vx, vy, heading|yaw, zero-mask
SITE_PUSH = 1           # legged_robot.py:539; lanes x, y
SITE_CURRICULUM = 2     # legged_robot.py:572 randint_like; lane 0 (raw u32 % max_level)
SITE_RESET_DOFS = 3     # legged_robot.py:491; lanes 0..num_dof-1
SITE_RESET_ROOT = 4     # legged_robot.py:520 (lanes 0,1) and :526 (lanes 2..7)
SITE_CMD_RESET = 5      # go2.py:230 -> _resample_commands; lanes as SITE_CMD_PERIODIC
SITE_OBS_NOISE = 6      # go2.py:519 rand_like; lanes 0..num_proprio-1
SITE_ACTION_NOISE = 7   # actor_critic.py:204 Normal.sample; lanes 2a, 2a+1 (Box-Muller pair of action a)


def noise_lane(i):
    """Observation-noise element i -> keyed lane.  Elements i and i+32 (the two a GPU lane owns) share one Philox
    block: block = i & 15, word = i >> 4  (csrc/env_core.cuh stage 3)."""
    i = np.asarray(i)
    return ((i & 15) << 2) | (i >> 4)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11). All args broadcastable uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & MASK, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def keyed_u32(seed, site, step, env_ids, lanes):
    """raw uint32 draws, shape [len(env_ids), len(lanes)]."""
    env = np.asarray(env_ids, dtype=np.uint64).reshape(-1, 1)
    lanes = np.asarray(lanes, dtype=np.uint64).reshape(1, -1)
    out = philox4x32_10(env, np.uint64(int(step) & 0xFFFFFFFF), np.uint64(site), lanes >> np.uint64(2),
                        int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    sel = (lanes & np.uint64(3)).astype(np.int64)
    sel = np.broadcast_to(sel, out[0].shape)
    stacked = np.stack(out, axis=-1)
    return np.take_along_axis(stacked, sel[..., None], axis=-1)[..., 0]


def keyed_uniform(seed, site, step, env_ids, lanes):
    """fp32 uniforms in [0,1), shape [len(env_ids), len(lanes)]."""
    r = keyed_u32(seed, site, step, env_ids, lanes)
    return ((r >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
