"""FluidWrappers on a device-free stand-in environment (reference behaviour: wrappers/*.py and the error strings
they raise); the same wrappers run on the CUDA environments in tests/test_gpu_parity.py."""
import numpy as np
import pytest
import torch

from fluidgym_b200 import spaces, wrappers


class FakeEnv:
    device = torch.device("cpu")
    use_marl = False
    n_envs = 3

    def __init__(self):
        self.last_action = None
        self.observation_space = spaces.Dict({"temperature": spaces.Box(0.0, 1.75, shape=(8, 4)),
                                              "velocity": spaces.Box(-np.inf, np.inf, shape=(2, 8, 4)),
                                              "pressure": spaces.Box(-np.inf, np.inf, shape=(8, 4))})
        self.action_space = spaces.Box(-1.0, 1.0, shape=(2, 1))

    def _obs(self):
        B = self.n_envs
        return {"temperature": torch.ones(B, 8, 4), "velocity": torch.zeros(B, 2, 8, 4), "pressure": torch.full((B, 8, 4), 2.0)}

    def reset(self, seed=None, randomize=None):
        return self._obs(), {}

    def step(self, action):
        self.last_action = action
        return self._obs(), torch.zeros(self.n_envs), False, False, {}


def test_flatten_keeps_the_environment_dimension():
    env = wrappers.FlattenObservation(FakeEnv())
    obs, _ = env.reset(seed=0)
    assert obs.shape == (3, 32 + 64)                       # temperature then velocity (DEFAULT_KEYS order)
    assert env.observation_space.shape == (96,)
    assert torch.all(obs[:, :32] == 1) and torch.all(obs[:, 32:] == 0)
    obs, r, term, trunc, info = env.step(torch.zeros(3, 2, 1))
    assert obs.shape == (3, 96)


def test_obs_extraction_and_its_errors():
    env = wrappers.ObsExtraction(FakeEnv(), ["pressure"])
    obs, _ = env.reset(seed=0)
    assert list(obs) == ["pressure"] and list(env.observation_space.spaces) == ["pressure"]
    with pytest.raises(ValueError, match="Key 'vorticity' not found in observation space."):
        wrappers.ObsExtraction(FakeEnv(), ["vorticity"])
    with pytest.raises(ValueError, match="non-empty"):
        wrappers.ObsExtraction(FakeEnv(), [])


def test_noise_wrappers_are_seeded_and_scaled():
    base = FakeEnv()
    a = torch.zeros(3, 2, 1)
    e1, e2 = wrappers.ActionNoise(base, sigma=0.5, seed=1), wrappers.ActionNoise(FakeEnv(), sigma=0.5, seed=1)
    e1.step(a); e2.step(a)
    assert torch.equal(e1.unwrapped.last_action, e2.unwrapped.last_action)
    assert 0.05 < float(e1.unwrapped.last_action.std()) < 2.0
    s = wrappers.SensorNoise(FakeEnv(), sigma=0.1, seed=3)
    obs, _ = s.reset(seed=0)
    assert 0.05 < float((obs["temperature"] - 1).std()) < 0.2
    obs2, *_ = s.step(a)
    assert not torch.equal(obs["temperature"], obs2["temperature"])
    # wrappers stack and forward unknown attributes
    st = wrappers.FlattenObservation(wrappers.SensorNoise(FakeEnv(), 0.0, 0))
    assert st.n_envs == 3 and isinstance(st.unwrapped, FakeEnv)
