"""Minimal stand-in for the slice of OpenAI gym 0.10.5 that NavGym-v0 touches.

The reference registers its env with ``gym.envs.registration.register`` and derives from
``gym.Env`` / ``gym.utils.EzPickle`` (reference nav_gym_env/__init__.py:1-40,
env.py:30,56-78,133-142).  gym is not installable in the build image (no network), so the
drop-in surface ships this registry + spaces shim.  When a real ``gym`` (or ``gymnasium``)
is importable, :func:`install` leaves it alone and only registers ``NavGym-v0`` with it.
"""
import importlib
import sys
import types

import numpy as np


class Space(object):
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)

    def sample(self):
        raise NotImplementedError

    def contains(self, x):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)


class Box(Space):
    """gym.spaces.Box: either (low, high arrays) or (scalar low, scalar high, shape)."""

    def __init__(self, low=None, high=None, shape=None, dtype=np.float32):
        if shape is None:
            low = np.asarray(low)
            high = np.asarray(high)
            assert low.shape == high.shape
            shape = low.shape
        else:
            low = np.full(shape, low, dtype=np.float64)
            high = np.full(shape, high, dtype=np.float64)
        super(Box, self).__init__(shape, dtype)
        self.low = low.astype(self.dtype)
        self.high = high.astype(self.dtype)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1e6)
        hi = np.where(np.isfinite(self.high), self.high, 1e6)
        return np.random.uniform(low=lo, high=hi, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    def __repr__(self):
        return "Box" + str(self.shape)


class Dict(Space):
    def __init__(self, spaces):
        super(Dict, self).__init__(None, None)
        self.spaces = dict(spaces)

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}

    def contains(self, x):
        return isinstance(x, dict) and all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

    def __getitem__(self, k):
        return self.spaces[k]

    def __repr__(self):
        return "Dict(" + ", ".join("%s:%r" % kv for kv in self.spaces.items()) + ")"


class Env(object):
    metadata = {'render.modes': []}
    reward_range = (-float('inf'), float('inf'))
    spec = None
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode='human'):
        raise NotImplementedError

    def close(self):
        return

    def seed(self, seed=None):
        return []

    @property
    def unwrapped(self):
        return self


class EzPickle(object):
    """Pickle an object by its constructor arguments (gym.utils.EzPickle)."""

    def __init__(self, *args, **kwargs):
        self._ezpickle_args = args
        self._ezpickle_kwargs = kwargs

    def __getstate__(self):
        return {"_ezpickle_args": self._ezpickle_args, "_ezpickle_kwargs": self._ezpickle_kwargs}

    def __setstate__(self, d):
        out = type(self)(*d["_ezpickle_args"], **d["_ezpickle_kwargs"])
        self.__dict__.update(out.__dict__)


class EnvSpec(object):
    def __init__(self, id, entry_point=None, kwargs=None, **_ignored):
        self.id = id
        self.entry_point = entry_point
        self.kwargs = {} if kwargs is None else kwargs

    def make(self, **overrides):
        kw = dict(self.kwargs)
        kw.update(overrides)
        ep = self.entry_point
        if callable(ep):
            cls = ep
        else:
            mod, name = ep.split(':')
            cls = getattr(importlib.import_module(mod), name)
        env = cls(**kw)
        env.spec = self
        return env


registry = {}


def register(id, **kwargs):
    registry[id] = EnvSpec(id, **kwargs)


def make(id, **kwargs):
    if id not in registry:
        raise KeyError("No registered env with id: %s" % id)
    return registry[id].make(**kwargs)


def as_module():
    """Package the shim as a module tree shaped like ``gym`` (gym, gym.spaces, gym.utils,
    gym.envs.registration) so that ``import gym; gym.make('NavGym-v0')`` works."""
    gym = types.ModuleType('gym')
    spaces = types.ModuleType('gym.spaces')
    utils = types.ModuleType('gym.utils')
    envs = types.ModuleType('gym.envs')
    registration = types.ModuleType('gym.envs.registration')
    spaces.Space, spaces.Box, spaces.Dict = Space, Box, Dict
    utils.EzPickle = EzPickle
    registration.register, registration.make = register, make
    registration.registry, registration.EnvSpec = registry, EnvSpec
    envs.registration = registration
    envs.register = register
    gym.Env, gym.Space = Env, Space
    gym.spaces, gym.utils, gym.envs = spaces, utils, envs
    gym.make, gym.register = make, register
    gym.__version__ = '0.10.5+navgym_b200.shim'
    gym.__navgym_shim__ = True
    return {'gym': gym, 'gym.spaces': spaces, 'gym.utils': utils, 'gym.envs': envs,
            'gym.envs.registration': registration}


def install(force=False):
    """Return a gym-like module.  Uses the real gym if importable, else installs the shim
    into ``sys.modules``."""
    if not force:
        if 'gym' in sys.modules:
            return sys.modules['gym']
        try:
            return importlib.import_module('gym')
        except ImportError:
            pass
    mods = as_module()
    sys.modules.update(mods)
    return mods['gym']
