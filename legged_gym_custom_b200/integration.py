"""Reference-side bindings: the two pieces a maintainer of the reference adds to put libb200gym.so under
its own scripts -- both exercised by tests/test_dropin_gpu.py.

* `bind_reference_runner(module)`: the reference's `rsl_rl.runners.on_policy_runner` builds `ActorCritic`, `MlpEstimator`
  and `PPO` by name (on_policy_runner.py:33-78); rebinding those three names makes its UNMODIFIED `OnPolicyRunner`
  (constructor, `learn`, `log`, `save`, `load`) drive the kernel-backed classes and a `Go2Env`.
* `IsaacGymPhysX`: the PhysX provider over Isaac Gym's tensor API (legged_robot.py:79-88, :504-506, :530-540, go2.py:352-353):
  wraps the simulator's four state tensors once (`bind`, zero copies) and forwards simulate / refresh / push-back.
"""
import torch


def bind_reference_runner(on_policy_runner_module):
    """Rebind the three class names the reference's runner module instantiates; returns its OnPolicyRunner class."""
    from .learner import PPO
    from .networks import ActorCritic, MlpEstimator
    m = on_policy_runner_module
    m.ActorCritic, m.MlpEstimator, m.PPO = ActorCritic, MlpEstimator, PPO
    return m.OnPolicyRunner


class IsaacGymPhysX:
    """`gym`, `sim`: the Isaac Gym handles; `gymtorch`: the isaacgym.gymtorch module (wrap_tensor / unwrap_tensor)."""

    def __init__(self, gym, sim, gymtorch):
        self.gym, self.sim, self.gt = gym, sim, gymtorch

    def bind(self, env):
        """once, from Go2Env.__init__: the env's PhysX-owned slots become the simulator's own tensors"""
        g, s, wrap = self.gym, self.sim, self.gt.wrap_tensor
        b = env.bufs
        b.rebind("root_states", wrap(g.acquire_actor_root_state_tensor(s)))
        b.rebind("dof_state", wrap(g.acquire_dof_state_tensor(s)))
        b.rebind("contact_forces", wrap(g.acquire_net_contact_force_tensor(s)))
        b.rebind("rigid_body_states", wrap(g.acquire_rigid_body_state_tensor(s)))

    def begin_step(self, env):
        pass

    def simulate(self, env, substep):                              # legged_robot.py:81-85
        self.gym.set_dof_actuation_force_tensor(self.sim, self.gt.unwrap_tensor(env.torques))
        self.gym.simulate(self.sim)
        if env.device.type == "cpu":
            self.gym.fetch_results(self.sim, True)
        self.gym.refresh_dof_state_tensor(self.sim)

    def refresh(self, env):                                        # go2.py:352-353, :272
        self.gym.refresh_actor_root_state_tensor(self.sim)
        self.gym.refresh_net_contact_force_tensor(self.sim)
        self.gym.refresh_rigid_body_state_tensor(self.sim)

    def push_state(self, env):                                     # legged_robot.py:504-506, :530-532, :540
        ids = env.reset_buf.nonzero().flatten().to(torch.int32)    # the one host-visible sync, and only because PhysX wants a count
        if len(ids):
            u = self.gt.unwrap_tensor
            self.gym.set_dof_state_tensor_indexed(self.sim, u(env.dof_state), u(ids), len(ids))
            self.gym.set_actor_root_state_tensor_indexed(self.sim, u(env.root_states), u(ids), len(ids))
