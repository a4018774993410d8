"""gymapi subset: a fake Gym whose state tensors are plain torch CPU tensors and whose
`simulate` / `refresh_*` calls invoke replay hooks (PhysX is an opaque producer of state
tensors; BASELINE.json north_star)."""
import numpy as np
import torch

SIM_PHYSX = 1
SIM_FLEX = 0
KEY_ESCAPE = 0
KEY_V = 1

# go2.urdf with collapse_fixed_joints (Head_*/foot joints keep dont_collapse="true")
GO2_BODIES = ["base", "Head_upper", "Head_lower"] + [
    f"{leg}_{part}" for leg in ("FL", "FR", "RL", "RR") for part in ("hip", "thigh", "calf", "foot")]
GO2_DOFS = [f"{leg}_{j}_joint" for leg in ("FL", "FR", "RL", "RR") for j in ("hip", "thigh", "calf")]
GO2_LOWER = [-1.0472, -1.5708, -2.7227] * 2 + [-1.0472, -0.5236, -2.7227] * 2
GO2_UPPER = [1.0472, 3.4907, -0.83776] * 2 + [1.0472, 4.5379, -0.83776] * 2
GO2_VEL = [30.1, 30.1, 20.07] * 4
GO2_EFFORT = [23.7, 23.7, 35.55] * 4


class Vec3:
    def __init__(self, x=0., y=0., z=0.):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def __add__(self, o):
        return Vec3(self.x + o.x, self.y + o.y, self.z + o.z)


class Transform:
    def __init__(self, p=None, r=None):
        self.p = p if p is not None else Vec3()
        self.r = r


class _Bag:
    pass


class _PhysX(_Bag):
    use_gpu = False
    num_subscenes = 0
    num_threads = 0


class SimParams(_Bag):
    def __init__(self):
        self.physx = _PhysX()
        self.use_gpu_pipeline = False
        self.dt = 1. / 60.
        self.substeps = 2


class AssetOptions(_Bag):
    pass


class PlaneParams(_Bag):
    pass


class CameraProperties(_Bag):
    pass


class TriangleMeshParams(_Bag):
    def __init__(self):
        self.transform = Transform()


class HeightFieldParams(_Bag):
    def __init__(self):
        self.transform = Transform()


class _BodyProps:
    def __init__(self):
        self.mass = 6.921
        self.com = Vec3(0.021112, 0., -0.005366)


class FakeGym:
    """Holds the four PhysX state tensors; replay hooks write them in place."""

    def __init__(self):
        self.num_envs = 0
        self.root = self.dof = self.contact = self.rigid = None
        self.on_simulate = None        # hook(gym): called per decimation substep
        self.on_refresh_root = None    # hook(gym): called at the top of post_physics_step

    def _alloc(self):
        if self.root is None:
            n = self.num_envs
            self.root = torch.zeros(n, 13)
            self.root[:, 6] = 1.0
            self.dof = torch.zeros(n * len(GO2_DOFS), 2)
            self.contact = torch.zeros(n * len(GO2_BODIES), 3)
            self.rigid = torch.zeros(n * len(GO2_BODIES), 13)

    # --- sim / assets -------------------------------------------------------------
    def create_sim(self, *a):
        return self

    def prepare_sim(self, sim):
        pass

    def add_ground(self, *a):
        pass

    def add_heightfield(self, *a):
        pass

    def add_triangle_mesh(self, *a):
        pass

    def load_asset(self, *a):
        return "go2"

    def get_asset_dof_count(self, a):
        return len(GO2_DOFS)

    def get_asset_rigid_body_count(self, a):
        return len(GO2_BODIES)

    def get_asset_rigid_body_names(self, a):
        return list(GO2_BODIES)

    def get_asset_dof_names(self, a):
        return list(GO2_DOFS)

    def get_asset_dof_properties(self, a):
        props = np.zeros(len(GO2_DOFS), dtype=[("lower", "f4"), ("upper", "f4"), ("velocity", "f4"), ("effort", "f4")])
        props["lower"], props["upper"], props["velocity"], props["effort"] = GO2_LOWER, GO2_UPPER, GO2_VEL, GO2_EFFORT
        return props

    def get_asset_rigid_shape_properties(self, a):
        return [_Bag() for _ in range(4)]

    def set_asset_rigid_shape_properties(self, *a):
        pass

    def create_env(self, *a):
        self.num_envs += 1
        return self.num_envs - 1

    def create_actor(self, *a):
        return 0

    def set_actor_dof_properties(self, *a):
        pass

    def get_actor_rigid_body_properties(self, *a):
        return [_BodyProps() for _ in GO2_BODIES]

    def set_actor_rigid_body_properties(self, *a, **k):
        pass

    def find_actor_rigid_body_handle(self, env, actor, name):
        return GO2_BODIES.index(name)

    # --- state tensors ------------------------------------------------------------
    def acquire_actor_root_state_tensor(self, sim):
        self._alloc()
        return self.root

    def acquire_dof_state_tensor(self, sim):
        self._alloc()
        return self.dof

    def acquire_net_contact_force_tensor(self, sim):
        self._alloc()
        return self.contact

    def acquire_rigid_body_state_tensor(self, sim):
        self._alloc()
        return self.rigid

    def refresh_dof_state_tensor(self, sim):
        pass

    def refresh_actor_root_state_tensor(self, sim):
        if self.on_refresh_root is not None:
            self.on_refresh_root(self)

    def refresh_net_contact_force_tensor(self, sim):
        pass

    def refresh_rigid_body_state_tensor(self, sim):
        pass

    def set_dof_actuation_force_tensor(self, sim, t):
        self.last_torques = t

    def simulate(self, sim):
        if self.on_simulate is not None:
            self.on_simulate(self)

    def fetch_results(self, *a):
        pass

    def set_dof_state_tensor_indexed(self, *a):
        pass

    def set_actor_root_state_tensor_indexed(self, *a):
        pass

    def set_actor_root_state_tensor(self, *a):
        pass


_GYM = None


def acquire_gym():
    global _GYM
    _GYM = FakeGym()
    return _GYM
