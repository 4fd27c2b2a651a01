"""Minimal clean-room stand-in for the classic `gym` 0.1x API surface that
mila-iqia/babyai touches (gym.Env, gym.Wrapper, gym.spaces, gym.make,
gym.envs.registration.register, gym.utils.seeding).

TEST INFRASTRUCTURE ONLY. The real `gym` is a third-party dependency of the
reference (setup.py:10 `gym>=0.9.6`) that is not installed in this image.
This package exists so the reference's own, unmodified babyai.levels.*,
babyai.rl.utils.penv and babyai.evaluate can be imported and executed as the
parity anchor.
"""
from . import error, spaces, utils
from .core import Env, Wrapper, ObservationWrapper
from . import envs
from .envs.registration import make, register, spec

__version__ = "0.17.shim"
