"""B200-native hot path of JustinMLu/legged_gym_custom (env torch path + rsl_rl learner).

See DESIGN.md. The CUDA library (csrc/ -> libb200gym.so) is loaded lazily by `_lib`;
every product entry point fails loudly if it is missing -- there is no CPU fallback.
"""
__version__ = "0.1.0"
