"""``FluidWrappers`` for batched environments (reference: ``fluidgym/wrappers/*.py``).

Same classes, constructor arguments and error messages as the reference; the only difference is the leading
environment dimension ``B`` every tensor carries, so "flatten" keeps ``B`` (and the agent dimension in MARL
mode) and noise is drawn per environment from one device generator.
"""
from __future__ import annotations

import torch

from . import spaces
from .spaces import flatten_dict_space

DEFAULT_KEYS = ["temperature", "velocity"]     # wrappers/flatten_obs.py:10


class FluidWrapper:
    """wrappers/fluid_wrapper.py:15-263: forwards everything to the wrapped environment."""

    def __init__(self, env):
        self._env = env

    def __getattr__(self, name):
        if name == "_env":
            raise AttributeError(name)
        return getattr(self._env, name)

    @property
    def unwrapped(self):
        env = self._env
        while isinstance(env, FluidWrapper):
            env = env._env
        return env

    @property
    def observation_space(self):
        return self._env.observation_space

    @property
    def action_space(self):
        return self._env.action_space

    def reset(self, seed=None, randomize=None):
        return self._env.reset(seed=seed, randomize=randomize)

    def step(self, action):
        return self._env.step(action)

    def _device(self):
        return getattr(self._env, "device", None) or getattr(self._env, "cuda_device")


class ActionNoise(FluidWrapper):
    """wrappers/action_noise.py:9-67: ``action + N(0, sigma)`` before stepping."""

    def __init__(self, env, sigma: float, seed: int):
        super().__init__(env)
        self._sigma = sigma
        self._rng = torch.Generator(device=self._device()).manual_seed(seed)

    def step(self, action):
        action = torch.as_tensor(action, dtype=torch.float32, device=self._device())
        noise = torch.randn(action.shape, generator=self._rng, device=action.device, dtype=action.dtype)
        return self._env.step(action + noise * self._sigma)


class SensorNoise(FluidWrapper):
    """wrappers/sensor_noise.py:9-100: Gaussian noise on every observation tensor (reset and step)."""

    def __init__(self, env, sigma: float, seed: int):
        super().__init__(env)
        self._sigma = sigma
        self._rng = torch.Generator(device=self._device()).manual_seed(seed)

    def _add_noise(self, obs):
        return {k: v + torch.randn(v.shape, generator=self._rng, device=v.device, dtype=v.dtype) * self._sigma for k, v in obs.items()}

    def reset(self, seed=None, randomize=None):
        obs, info = self._env.reset(seed=seed, randomize=randomize)
        return self._add_noise(obs), info

    def step(self, action):
        obs, *rest = self._env.step(action)
        return (self._add_noise(obs), *rest)


class ObsExtraction(FluidWrapper):
    """wrappers/obs_extraction.py:10-107: keep only the listed observation keys."""

    def __init__(self, env, keys):
        super().__init__(env)
        if len(keys) == 0:
            raise ValueError("Keys list must be non-empty or None.")
        if not isinstance(self._env.observation_space, spaces.Dict):
            raise ValueError("ObsExtraction wrapper only supports Dict observation spaces.")
        for k in keys:
            if k not in self._env.observation_space.spaces:
                raise ValueError(f"Key '{k}' not found in observation space.")
        self._keys = list(keys)
        self._space = spaces.Dict({k: self._env.observation_space.spaces[k] for k in keys})

    @property
    def observation_space(self):
        return self._space

    def _filter(self, obs):
        return {k: obs[k] for k in self._keys}

    def reset(self, seed=None, randomize=None):
        obs, info = self._env.reset(seed=seed, randomize=randomize)
        return self._filter(obs), info

    def step(self, action):
        obs, *rest = self._env.step(action)
        return (self._filter(obs), *rest)


class FlattenObservation(FluidWrapper):
    """wrappers/flatten_obs.py:13-102: concatenate the (``temperature``, ``velocity``) entries into one vector
    per environment (per agent in MARL mode)."""

    def __init__(self, env):
        super().__init__(env)
        if not isinstance(self._env.observation_space, spaces.Dict):
            raise ValueError("FlattenObservation wrapper only supports Dict observation spaces.")
        self._keys = [k for k in DEFAULT_KEYS if k in self._env.observation_space.spaces]
        self._space = flatten_dict_space(self._env.observation_space, self._keys)
        # reference: start_dim = 1 if use_marl else 0; one more here for the environment dimension
        self._start = 2 if getattr(env, "use_marl", False) else 1

    @property
    def observation_space(self):
        return self._space

    def _flatten(self, obs):
        return torch.cat([obs[k].flatten(start_dim=self._start) for k in self._keys], dim=self._start)

    def reset(self, seed=None, randomize=None):
        obs, info = self._env.reset(seed=seed, randomize=randomize)
        return self._flatten(obs), info

    def step(self, action):
        obs, *rest = self._env.step(action)
        return (self._flatten(obs), *rest)
