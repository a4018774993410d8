"""CPU-side guard: the host classes expose every method the runner / bench call (no GPU needed)."""
import inspect

from legged_gym_custom_b200 import env, learner, networks, runner


def test_methods_exist():
    for cls, names in ((learner.PPO, ["act", "process_env_step", "compute_returns", "update", "update_dagger", "update_with_indices",
                                      "update_dagger_with_indices", "set_device_counter", "_run_captured", "_minibatch", "_dagger_minibatch",
                                      "_fork", "_fork_onto", "_join", "_on", "_capped", "_offload", "_adaptive_lr", "_adam", "_gather_storage", "init_storage", "enforce_max_std"]),
                       (runner.OnPolicyRunner, ["learn", "iteration", "rollout", "_rollout_eager", "enable_graphs", "save", "load",
                                                "get_inference_policy", "log"]),
                       (env.Go2Env, ["step", "step5", "reset", "reset_idx", "set_device_counter", "get_observations",
                                     "get_privileged_observations", "get_critic_observations", "get_estimated_observations",
                                     "get_scan_observations", "get_heights"]),
                       (networks.ActorCritic, ["act", "act_inference", "evaluate", "get_actions_log_prob", "update_distribution",
                                               "privileged_encoder", "adaptation_encoder", "state_dict", "load_state_dict", "fwd_adapt",
                                               "bwd_adapt", "_fwd_adapt_gemms"]),
                       (networks.MlpEstimator, ["forward", "state_dict", "load_state_dict", "fwd"])):
        for n in names:
            assert callable(getattr(cls, n, None)), f"{cls.__name__}.{n} is missing"


def test_self_attribute_calls_resolve():
    """every `self.<name>(` call inside PPO / OnPolicyRunner / Go2Env refers to an attribute defined somewhere in the class source"""
    import re
    for cls in (learner.PPO, runner.OnPolicyRunner, env.Go2Env, networks.ActorCritic):
        src = inspect.getsource(cls)
        called = set(re.findall(r"self\.(_?[a-zA-Z_][a-zA-Z0-9_]*)\(", src))
        defined = set(re.findall(r"def (_?[a-zA-Z_][a-zA-Z0-9_]*)\(", src)) | set(re.findall(r"self\.(_?[a-zA-Z_][a-zA-Z0-9_]*)\s*=", src))
        defined |= set(dir(cls))
        missing = called - defined
        assert not missing, (cls.__name__, missing)
