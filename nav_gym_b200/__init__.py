"""nav_gym_b200 — B200-native batched simulator for nav-gym's NavGym-v0 per-step hot path."""
from . import gym_shim  # noqa: F401

__all__ = ['gym_shim']
