"""Minimal stand-in for NVIDIA Isaac Gym Preview 4, so that the UNMODIFIED reference
(/root/reference/legged_gym, rsl_rl) imports and runs on CPU with PhysX replaced by
replayed state tensors.  TEST INFRASTRUCTURE: used only by oracle/make_golden.py and
oracle/ref_runner.py inside the authoring container; never imported by the product.

Surface = exactly what the reference touches.
"""
from . import gymapi, gymutil, gymtorch, torch_utils  # noqa: F401
