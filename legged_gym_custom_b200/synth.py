"""Synthetic PhysX frames (BASELINE.json: "Benchmarks replay synthetic state tensors of
each task's shape"; recipe = SURVEY.md §8(d) state generator).

PhysX is out of scope and is treated as an opaque producer of the four Isaac Gym state
tensors (reference: legged_robot.py:632-646, go2.py:136-138).  One *frame set* per env
step holds what PhysX would have written:

    dof      [decimation, N*12, 2]   interleaved (pos, vel), one per decimation substep
    root     [N, 13]                 pos3, quat xyzw, linvel3, angvel3
    contact  [N*19, 3]               net contact force per rigid body
    rigid    [N*19, 13]              rigid body states (only the 4 foot z are consumed)

Host-side numpy, seeded; used by tests, bench.py and the golden-vector generator so that
the reference, the oracle and the CUDA kernels all see identical tensors.
"""
import numpy as np

NUM_DOF = 12
NUM_BODIES = 19
FEET = (6, 10, 14, 18)          # FL, FR, RL, RR foot body indices
DEFAULT_DOF_POS = np.array([0.1, 0.8, -1.5, -0.1, 0.8, -1.5, 0.1, 1.0, -1.5, -0.1, 1.0, -1.5], dtype=np.float32)
DOF_LOWER = np.array([-1.0472, -1.5708, -2.7227] * 2 + [-1.0472, -0.5236, -2.7227] * 2, dtype=np.float32)
DOF_UPPER = np.array([1.0472, 3.4907, -0.83776] * 2 + [1.0472, 4.5379, -0.83776] * 2, dtype=np.float32)


def _quat_mul(a, b):
    ax, ay, az, aw = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    bx, by, bz, bw = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], axis=1)


def make_frames(num_envs, env_origins, rng, decimation=4, hole_prob=0.002, flip_prob=0.002, body_hit_prob=0.005):
    """One env step worth of PhysX output. `env_origins` [N,3] fp32; `rng` np.random.Generator."""
    n = num_envs
    f32 = np.float32
    root = np.zeros((n, 13), dtype=f32)
    root[:, 0] = env_origins[:, 0] + 2.0 + rng.uniform(0.0, 24.0, n)
    root[:, 1] = env_origins[:, 1] + rng.uniform(-1.0, 1.0, n)
    root[:, 2] = env_origins[:, 2] + rng.uniform(0.30, 0.45, n)
    root[rng.random(n) < hole_prob, 2] = -1.5
    tilt = np.concatenate([rng.normal(0, 0.05, (n, 2)), np.zeros((n, 1)), np.ones((n, 1))], axis=1)
    yaw = rng.uniform(-0.3, 0.3, n)
    qyaw = np.stack([np.zeros(n), np.zeros(n), np.sin(yaw / 2), np.cos(yaw / 2)], axis=1)
    q = _quat_mul(qyaw, tilt)
    flip = rng.random(n) < flip_prob
    q[flip] = _quat_mul(q[flip], np.tile(np.array([[1.0, 0, 0, 0.05]]), (int(flip.sum()), 1)))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    root[:, 3:7] = q
    root[:, 7:10] = rng.normal(0, 0.5, (n, 3))
    root[:, 7] += 0.8
    root[:, 10:13] = rng.normal(0, 0.5, (n, 3))

    dof = np.zeros((decimation, n * NUM_DOF, 2), dtype=f32)
    base_pos = DEFAULT_DOF_POS[None] + rng.normal(0, 0.2, (n, NUM_DOF))
    for k in range(decimation):
        pos = np.clip(base_pos + rng.normal(0, 0.02, (n, NUM_DOF)), DOF_LOWER, DOF_UPPER)
        vel = rng.normal(0, 1.5, (n, NUM_DOF))
        dof[k, :, 0] = pos.reshape(-1)
        dof[k, :, 1] = vel.reshape(-1)

    contact = np.zeros((n, NUM_BODIES, 3), dtype=f32)
    for b in FEET:
        on = rng.random(n) < 0.5
        contact[:, b, 2] = np.where(on, 40.0 + rng.normal(0, 5, n), 0.0)
        contact[:, b, 0:2] = np.where(on[:, None], rng.normal(0, 3, (n, 2)), 0.0)
        stumble = rng.random(n) < 0.01
        contact[stumble, b, 0] = 400.0
    others = [b for b in range(NUM_BODIES) if b not in FEET]
    hit = rng.random((n, len(others))) < body_hit_prob
    contact[:, others, :] = np.where(hit[..., None], rng.normal(0, 20, (n, len(others), 3)), 0.0)

    rigid = rng.normal(0, 0.3, (n, NUM_BODIES, 13)).astype(f32)
    for b in FEET:
        rigid[:, b, 2] = env_origins[:, 2] + rng.uniform(0.02, 0.12, n)

    return dict(dof=dof, root=root, contact=contact.reshape(n * NUM_BODIES, 3).astype(f32),
                rigid=rigid.reshape(n * NUM_BODIES, 13))
