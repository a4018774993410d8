"""Empty matplotlib stand-in (legged_gym/utils/logger.py imports pyplot at module scope)."""
