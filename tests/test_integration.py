"""RL-library adapters (fluidgym_b200/integration.py) on a stand-in batched environment (CPU): slot layout, auto-reset with
terminated_observation, the reference's constructor checks (integration/gymnasium.py:22-37, sb3/vec_env.py:96-152,
pettingzoo.py:27-49)."""
import numpy as np
import pytest
import torch

from fluidgym_b200 import spaces
from fluidgym_b200.integration import GymFluidEnv, PettingZooFluidEnv, VecFluidEnv


class FakeEnv:
    """Batched environment with the call protocol of fluidgym_b200.envs.*: obs dict of [B, (A,) ...] tensors."""

    def __init__(self, n_envs, n_agents=1, use_marl=False, episode_length=3):
        self.n_envs, self.n_agents_, self.use_marl, self.episode_length = n_envs, n_agents, use_marl, episode_length
        self.device = torch.device("cpu")
        self.t = 0
        self.resets = 0
        self.last_action = None

    @property
    def n_agents(self):
        return self.n_agents_ if self.use_marl else 1

    @property
    def observation_space(self):
        return spaces.Dict({"velocity": spaces.Box(-1, 1, shape=(2, 4)), "pressure": spaces.Box(-1, 1, shape=(4,))})

    @property
    def action_space(self):
        return spaces.Box(-1.0, 1.0, shape=(1,))

    def _obs(self):
        lead = (self.n_envs, self.n_agents) if self.use_marl else (self.n_envs,)
        base = torch.arange(int(np.prod(lead)), dtype=torch.float32).reshape(lead) + 100 * self.t
        return {"velocity": base[..., None, None].expand(*lead, 2, 4).clone(), "pressure": base[..., None].expand(*lead, 4).clone()}

    def seed(self, seed):
        self._seed = seed

    def reset(self, seed=None, randomize=None):
        self.t = 0
        self.resets += 1
        return self._obs(), {}

    def step(self, action):
        self.last_action = action
        self.t += 1
        r = action.sum(dim=-1) if self.use_marl else action.sum(dim=-1)
        trunc = self.t >= self.episode_length
        return self._obs(), r, False, trunc, {"drag": torch.full((self.n_envs,), float(self.t))}


def test_vec_env_single_agent_slots_and_auto_reset():
    env = FakeEnv(n_envs=4)
    v = VecFluidEnv(env)
    assert v.num_envs == 4
    obs = v.reset(seed=1)
    assert obs["velocity"].shape == (4, 2, 4) and obs["pressure"].shape == (4, 4)
    for k in range(3):
        obs, rew, dones, infos = v.step(np.full((4, 1), 0.5, dtype=np.float32))
        assert rew.shape == (4,) and rew.dtype == np.float32 and np.allclose(rew, 0.5)
        assert env.last_action.shape == (4, 1)
        assert infos[2]["drag"] == float(k + 1)
    assert dones.all() and env.resets == 2                    # truncated after 3 steps -> auto reset
    assert infos[1]["TimeLimit.truncated"] is True
    assert infos[1]["terminated_observation"]["pressure"][0] == 301.0     # obs of the finished episode, slot 1, t = 3
    assert obs["pressure"][1, 0] == 1.0                       # obs after the reset
    assert v.get_attr("episode_length") == [3] * 4 and v.env_is_wrapped(object) == [False] * 4
    v.set_attr("episode_length", 5)
    assert env.episode_length == 5


def test_vec_env_marl_is_environment_major():
    env = FakeEnv(n_envs=2, n_agents=3, use_marl=True)
    v = VecFluidEnv(env, auto_reset=False)
    assert v.num_envs == 6
    obs = v.reset()
    assert obs["velocity"].shape == (6, 2, 4)
    assert obs["pressure"][:, 0].tolist() == [0, 1, 2, 3, 4, 5]           # slot = env * n_agents + agent
    acts = np.arange(6, dtype=np.float32).reshape(6, 1)
    obs, rew, dones, infos = v.step(acts)
    assert env.last_action.shape == (2, 3, 1) and env.last_action[1, 0, 0] == 3.0
    assert rew.tolist() == [0, 1, 2, 3, 4, 5] and not dones.any()
    assert infos[4]["drag"] == 1.0 and "terminated_observation" not in infos[0]


def test_gym_and_pettingzoo_views_and_their_checks():
    with pytest.raises(ValueError, match="does not support multi-agent environments"):
        GymFluidEnv(FakeEnv(1, 3, use_marl=True))
    with pytest.raises(ValueError, match="Unsupported render mode"):
        GymFluidEnv(FakeEnv(1), render_mode="human")
    with pytest.raises(ValueError, match="n_envs=1"):
        GymFluidEnv(FakeEnv(2))
    g = GymFluidEnv(FakeEnv(1))
    obs, info = g.reset(seed=0)
    assert obs["velocity"].shape == (2, 4)
    obs, r, term, trunc, info = g.step(np.array([0.25], dtype=np.float32))
    assert isinstance(r, float) and r == 0.25 and term is False and info["drag"] == 1.0

    with pytest.raises(ValueError, match="can only be used with MARL"):
        PettingZooFluidEnv(FakeEnv(1))
    pz = PettingZooFluidEnv(FakeEnv(1, 3, use_marl=True, episode_length=1))
    obs, infos = pz.reset(seed=0)
    assert list(obs) == ["agent_0", "agent_1", "agent_2"] and obs["agent_2"]["pressure"][0] == 2.0
    obs, rew, terms, truncs, infos = pz.step({a: np.array([i], dtype=np.float32) for i, a in enumerate(pz.possible_agents)})
    assert rew == {"agent_0": 0.0, "agent_1": 1.0, "agent_2": 2.0} and all(truncs.values()) and pz.agents == []


@pytest.mark.parametrize("marl", [False, True])
def test_torchrl_adapter_protocol(marl):
    """TorchRLFluidEnv (reference: integration/torchrl.py:87-278): batch size (n_envs,) / (n_envs, n_agents), spec shapes with the
    trailing unit dimension of reward / done flags, _step / _reset / _set_seed."""
    from fluidgym_b200.integration import TorchRLFluidEnv
    env = FakeEnv(3, n_agents=4, use_marl=marl, episode_length=2)
    tenv = TorchRLFluidEnv(env)
    lead = (3, 4) if marl else (3,)
    assert tuple(tenv.batch_size) == lead
    assert tuple(tenv.action_spec.shape) == (*lead, 1)
    assert tuple(tenv.reward_spec.shape) == (*lead, 1)
    assert tuple(tenv.observation_spec["velocity"].shape) == (*lead, 2, 4)
    assert tuple(tenv.done_spec["truncated"].shape) == (*lead, 1)
    tenv.set_seed(7)
    assert env._seed == 7
    td = tenv.reset()
    assert tuple(td["pressure"].shape) == (*lead, 4)
    a = torch.ones(*lead, 1)
    out = tenv.step({"action": a})["next"]
    assert tuple(out["reward"].shape) == (*lead, 1) and out["reward"].dtype == torch.float32
    assert not bool(out["done"].any()) and out["done"].dtype == torch.bool and tuple(out["done"].shape) == (*lead, 1)
    out = tenv.step({"action": a})["next"]
    assert bool(out["truncated"].all()) and bool(out["done"].all()) and not bool(out["terminated"].any())
    with pytest.raises(NotImplementedError):
        TorchRLFluidEnv(env, from_pixels=True)
