"""Minimal stand-in for the `gym` symbols the reference touches (gym is not installed in this image).

Used ONLY by tests/golden/make_golden.py to import the reference's chess_v2.py unmodified
(chess_v2.py:9-11,111,156-157,168,177 and gym_chess/__init__.py:3).
"""
import sys
import types

import numpy as np


def install():
    if "gym" in sys.modules:
        return
    gym = types.ModuleType("gym")

    class Env:
        metadata = {}

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == tuple(self.shape) and (x >= self.low).all() and (x <= self.high).all()

    class Discrete:
        def __init__(self, n):
            self.n = n

        def contains(self, x):
            try:
                return 0 <= int(x) < self.n and int(x) == x
            except (TypeError, ValueError):
                return False

        def sample(self):
            return int(np.random.randint(self.n))

    class Error(Exception):
        pass

    spaces = types.ModuleType("gym.spaces")
    spaces.Box, spaces.Discrete = Box, Discrete
    error = types.ModuleType("gym.error")
    error.Error = Error
    utils = types.ModuleType("gym.utils")
    # gym.utils.colorize as published by OpenAI gym (gym/utils/colorize.py; the reference pins gym>=0,<1, setup.py:39,
    # which is not under /root/reference): ANSI SGR code = colour number (+10 as a background), ";1" when bold
    color2num = dict(gray=30, red=31, green=32, yellow=33, blue=34, magenta=35, cyan=36, white=37, crimson=38)

    def colorize(string, color, bold=False, highlight=False):
        num = color2num[color] + (10 if highlight else 0)
        attrs = ";".join([str(num)] + (["1"] if bold else []))
        return "\x1b[%sm%s\x1b[0m" % (attrs, string)

    utils.colorize = colorize
    seeding = types.ModuleType("gym.utils.seeding")

    def np_random(seed=None):
        rng = np.random.RandomState(seed)
        return rng, seed

    seeding.np_random = np_random
    utils.seeding = seeding
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.registry = {}
    registration.register = lambda id, **kw: registration.registry.__setitem__(id, kw)
    envs.registration = registration
    gym.Env, gym.spaces, gym.error, gym.utils, gym.envs = Env, spaces, error, utils, envs
    for name, mod in [("gym", gym), ("gym.spaces", spaces), ("gym.error", error), ("gym.utils", utils),
                      ("gym.utils.seeding", seeding), ("gym.envs", envs), ("gym.envs.registration", registration)]:
        sys.modules[name] = mod
