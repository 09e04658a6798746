"""nav_gym_b200 — B200-native batched simulator for nav-gym's NavGym-v0 per-step hot path.

Importing the package registers ``NavGym-v0`` (reference nav_gym_env/__init__.py:4-40) with
gym — the real package when importable, otherwise the bundled registry/spaces shim — so
``gym.make('NavGym-v0')`` returns the drop-in :class:`nav_gym_b200.env.NavGymEnv`.  The
vectorised entry point is :class:`nav_gym_b200.batched_env.BatchedNavGym`.
"""
from . import gym_shim  # noqa: F401
from .env import register_env

register_env()

__all__ = ['gym_shim', 'register_env']
