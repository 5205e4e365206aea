"""``ParallelFluidEnv`` with the reference's constructor semantics (``envs/parallel_env.py:30-444``):
``len(cuda_ids)`` environments, environment i on device ``cuda_ids[i]``; ``reset``/``step`` take and return
tensors stacked over environments on the CPU.  The mechanism is different: instead of one OS process + one
CUDA context per environment talking over pipes, all environments that share a device form ONE batched
environment advanced by one launch sequence; devices are driven concurrently by one host thread each (the
C ABI calls release the GIL) and never communicate."""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import torch


class ParallelFluidEnv:
    def __init__(self, env_id: str, cuda_ids: list[int], **env_kwargs):
        import fluidgym_b200
        if env_kwargs.get("differentiable", False):
            raise ValueError("ParallelFluidEnv does not support differentiable environments")   # parallel_env.py:54-57
        self.cuda_ids = list(cuda_ids)
        self.n_envs = len(self.cuda_ids)
        self.devices = sorted(set(self.cuda_ids))
        self._index = {d: [i for i, c in enumerate(self.cuda_ids) if c == d] for d in self.devices}
        self.envs = {d: fluidgym_b200.make(env_id, n_envs=len(self._index[d]), device=f"cuda:{d}", **env_kwargs)
                     for d in self.devices}
        self._pool = ThreadPoolExecutor(max_workers=len(self.devices))
        first = self.envs[self.devices[0]]
        self.n_agents = first.n_agents
        self.episode_length = first.episode_length

    def _scatter(self, fn):
        futs = {d: self._pool.submit(fn, d) for d in self.devices}
        return {d: f.result() for d, f in futs.items()}

    def _gather(self, per_dev):
        """per_dev[d] = tensor / dict of tensors with leading dim len(index[d]) -> stacked in env order on CPU."""
        sample = per_dev[self.devices[0]]
        if isinstance(sample, dict):
            return {k: self._gather({d: per_dev[d][k] for d in self.devices}) for k in sample}
        out = [None] * self.n_envs
        for d in self.devices:
            t = per_dev[d].detach().cpu()
            for j, i in enumerate(self._index[d]):
                out[i] = t[j]
        return torch.stack(out)

    def seed(self, seed: int):
        for k, d in enumerate(self.devices):
            self.envs[d].seed(seed + k)

    def reset(self, seed: int | None = None, randomize: bool | None = None):
        res = self._scatter(lambda d: self.envs[d].reset(None if seed is None else seed + self.devices.index(d), randomize))
        return self._gather({d: r[0] for d, r in res.items()}), {}

    def step(self, action: torch.Tensor):
        if action.shape[0] != self.n_envs:
            raise ValueError(f"Action batch {action.shape[0]} does not match the number of environments {self.n_envs}.")

        def run(d):
            a = action[self._index[d]].to(f"cuda:{d}", non_blocking=True)
            return self.envs[d].step(a)

        res = self._scatter(run)
        obs = self._gather({d: r[0] for d, r in res.items()})
        reward = self._gather({d: r[1] for d, r in res.items()})
        info = self._gather({d: r[4] for d, r in res.items()})
        terminated = any(r[2] for r in res.values())
        truncated = any(r[3] for r in res.values())
        return obs, reward, terminated, truncated, info

    def sample_action(self):
        return self._gather(self._scatter(lambda d: self.envs[d].sample_action()))

    def close(self):
        self._pool.shutdown(wait=True)

    def render(self, *a, **k):            # parallel_env.py: render/get_state/set_state/detach are not supported
        raise NotImplementedError

    get_state = set_state = detach = render
