"""Minimal ``Box`` / ``Dict`` spaces with the attributes the reference's environments and wrappers read from
``gymnasium.spaces`` (``.shape``, ``.low``, ``.high``, ``.dtype``, ``.spaces``; envs/fluid_env.py:300-330,
wrappers/util.py:7-81).  If ``gymnasium`` is importable its classes are used instead, so adapters written
against gymnasium keep working."""
from __future__ import annotations

import numpy as np

try:                                       # pragma: no cover - gymnasium is not in this image
    from gymnasium.spaces import Box, Dict  # type: ignore
except Exception:
    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(shape) if shape is not None else tuple(np.shape(low))
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Dict(dict):
        def __init__(self, spaces=None, **kw):
            super().__init__(spaces or {}, **kw)
            self.spaces = self


def flatten_dict_space(space: "Dict", keys=None) -> "Box":
    """wrappers/util.py:25-81"""
    if not isinstance(space, Dict):
        raise TypeError(f"Expected spaces.Dict, got {type(space)}")
    if keys is not None:
        for k in keys:
            if k not in space.spaces:
                raise KeyError(f"Key '{k}' not found in the Dict space.")
    items = [(k, space.spaces[k]) for k in (keys if keys is not None else list(space.spaces))]
    if not items:
        raise ValueError("Dict space contains no Box subspaces to flatten.")
    low = np.concatenate([np.asarray(s.low).reshape(-1) for _, s in items]).astype(np.float32)
    high = np.concatenate([np.asarray(s.high).reshape(-1) for _, s in items]).astype(np.float32)
    return Box(low=low, high=high, dtype=np.float32)
