"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (fluidgym_b200/).

Makes the UNMODIFIED reference package installed under ``baseline/_ref`` importable in this image:

* the image lacks ``gymnasium``, ``seaborn``, ``matplotlib``, ``imageio`` and ``skimage``
   -> permissive stub modules are fabricated for them, with a
  minimal real ``gymnasium.spaces.Box/Dict`` because the envs read ``.shape`` from them;
* ``fluidgym/envs/util/__init__.py:3-7`` imports a symbol (``state_vjp``) that
  ``diff_tools.py`` does not define -> that one package ``__init__`` is bypassed by pre-seeding
  ``sys.modules`` with an empty package shell that still resolves its sub-modules from disk.

Nothing here changes reference behaviour on the solver path; it only lets ``import fluidgym``
succeed.  Used by ``oracle/ref_harness.py`` (golden-vector generation on the GPU box) and by
``bench.py --impl reference``.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = os.path.join(REPO, "baseline", "_ref")

_STUB_ROOTS = ("gymnasium", "seaborn", "matplotlib", "mpl_toolkits", "imageio", "skimage")


class _Anything:
    """Callable/attribute sink used for plotting APIs that are never exercised."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __getitem__(self, i):
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (_Anything,), {})
        setattr(self, name, obj)
        return obj


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _StubModule(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, module):
        if module.__name__ == "gymnasium.spaces":
            module.Box = Box
            module.Dict = Dict
            module.Space = Space
        if module.__name__ == "gymnasium":
            import importlib

            module.spaces = importlib.import_module("gymnasium.spaces")
        if module.__name__ == "matplotlib.path":
            module.Path = PolygonPath
        if module.__name__ == "seaborn":
            module.color_palette = lambda *a, **k: [(0.0, 0.0, 0.0)] * 10


class PolygonPath:
    """Stand-in for ``matplotlib.path.Path`` as used by ``airfoil_env_base.py:174-208`` (point-in-polygon mask of
    the render grid, which decides which sensor pixels are dropped): even-odd crossing rule."""

    def __init__(self, vertices, *a, **k):
        self.vertices = np.asarray(vertices, dtype=np.float64)

    def contains_points(self, points, *a, **k):
        pts = np.asarray(points, dtype=np.float64)
        x, y = pts[:, 0], pts[:, 1]
        v = self.vertices
        inside = np.zeros(len(pts), dtype=bool)
        n = len(v)
        for i in range(n):
            x0, y0 = v[i]
            x1, y1 = v[(i + 1) % n]
            if y0 == y1:
                continue
            cond = (y0 > y) != (y1 > y)
            xi = x0 + (y - y0) * (x1 - x0) / (y1 - y0)
            inside ^= cond & (x < xi)
        return inside


class Space:
    pass


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.dtype = low, high, dtype
        self.shape = tuple(shape) if shape is not None else np.shape(low)

    def sample(self):
        lo = np.broadcast_to(np.asarray(self.low, dtype=np.float64), self.shape)
        hi = np.broadcast_to(np.asarray(self.high, dtype=np.float64), self.shape)
        return np.random.uniform(np.nan_to_num(lo, neginf=-1), np.nan_to_num(hi, posinf=1)).astype(self.dtype)


class Dict(Space, dict):
    def __init__(self, spaces=None, **kw):
        dict.__init__(self, spaces or {}, **kw)
        self.spaces = self

    def sample(self):
        return {k: v.sample() for k, v in self.items()}


def install(ref_root: str = REF_ROOT) -> str:
    """Put the reference on ``sys.path`` with the shims active; returns the root used."""
    if not os.path.isdir(os.path.join(ref_root, "fluidgym")):
        raise FileNotFoundError(
            f"reference install not found under {ref_root}; run oracle/build_ref.sh where /root/reference exists"
        )
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        missing = []
        for root in _STUB_ROOTS:
            try:
                __import__(root)
            except Exception:
                missing.append(root)
        if missing:
            globals()["_STUB_ROOTS"] = tuple(missing)
            sys.meta_path.append(_StubFinder())
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    name = "fluidgym.envs.util"
    if name not in sys.modules:
        shell = types.ModuleType(name)
        shell.__path__ = [os.path.join(ref_root, "fluidgym", "envs", "util")]
        shell.__package__ = name
        sys.modules[name] = shell
    return ref_root


def load_pisotorch(ref_root: str = REF_ROOT):
    """Import only the compiled extension (no python package side effects)."""
    install(ref_root)
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    import importlib

    return importlib.import_module("fluidgym.simulation.extensions.PISOtorch")
