"""fluidgym_b200 -- B200-native (sm_100a) batched implementation of FluidGym's PISO solver step.

Drop-in scope: the fluid-solver step behind ``fluidgym.make(...).step()`` --
assembly, BiCGStab predictor, pressure CG, velocity correction, outflow / jet boundary conditions,
drag-lift reward and sensor sampling -- for whole batches of environments per launch.
"""
from __future__ import annotations

__version__ = "0.1.0"

_REGISTRY = {}


def register(env_id: str, entry_point, **defaults):
    _REGISTRY[env_id] = (entry_point, defaults)


def make(env_id: str, n_envs: int = 1, **kwargs):
    """``fluidgym.make`` (registry.py:101-116) with an extra ``n_envs``: the returned environment steps
    ``n_envs`` independent copies at once (the reference needs one OS process per copy,
    envs/parallel_env.py:162-175)."""
    if env_id not in _REGISTRY:
        _register_defaults()
    if env_id not in _REGISTRY:
        raise KeyError(f"unknown environment id {env_id!r}; available: {sorted(_REGISTRY)}")
    entry, defaults = _REGISTRY[env_id]
    cfg = dict(defaults)
    # reference-only constructor flags that have no meaning here are accepted and ignored
    import inspect
    params = inspect.signature(entry.__init__).parameters
    # load_domain_statistics (fluid_env.py:234-238): read after construction; False unless asked for (the files come from the
    # reference's HuggingFace dataset, envs/common.py::DomainStatistics)
    load_stats = bool(kwargs.pop("load_domain_statistics", False)) if "load_domain_statistics" not in params else False
    for k in ("load_initial_domain", "dtype"):
        if k not in params:
            kwargs.pop(k, None)
    if "differentiable" in kwargs:
        if "differentiable" not in params:
            if kwargs.pop("differentiable"):
                raise NotImplementedError(f"{env_id}: differentiable=True is not available for this environment family yet")
    cfg.update(kwargs)
    env = entry(n_envs=n_envs, **cfg)
    if load_stats:
        env.load_domain_statistics()
    return env


def _register_defaults():
    from .envs.cylinder import CYLINDER_JET_2D_DEFAULT_CONFIG, CylinderJet2DEnv
    # fluidgym/__init__.py:28-60: easy = Re 100 / res 24, medium = Re 250 / res 32, hard = Re 500 / res 32
    register("CylinderJet2D-easy-v0", CylinderJet2DEnv, **CYLINDER_JET_2D_DEFAULT_CONFIG)
    register("CylinderJet2D-medium-v0", CylinderJet2DEnv, **{**CYLINDER_JET_2D_DEFAULT_CONFIG, "reynolds_number": 250.0, "resolution": 32})
    from .envs.cylinder import CYLINDER_ROT_2D_DEFAULT_CONFIG, CylinderRot2DEnv
    # fluidgym/__init__.py:52-74
    register("CylinderRot2D-easy-v0", CylinderRot2DEnv, **CYLINDER_ROT_2D_DEFAULT_CONFIG)
    register("CylinderRot2D-medium-v0", CylinderRot2DEnv, **{**CYLINDER_ROT_2D_DEFAULT_CONFIG, "reynolds_number": 250.0, "resolution": 32})
    register("CylinderRot2D-hard-v0", CylinderRot2DEnv, **{**CYLINDER_ROT_2D_DEFAULT_CONFIG, "reynolds_number": 500.0, "resolution": 32})
    from .envs.airfoil import AIRFOIL_2D_DEFAULT_CONFIG, Airfoil2DEnv
    # fluidgym/__init__.py:306-328
    register("Airfoil2D-easy-v0", Airfoil2DEnv, **{**AIRFOIL_2D_DEFAULT_CONFIG, "reynolds_number": 1e3})
    register("Airfoil2D-medium-v0", Airfoil2DEnv, **{**AIRFOIL_2D_DEFAULT_CONFIG, "reynolds_number": 3e3})
    register("Airfoil2D-hard-v0", Airfoil2DEnv, **{**AIRFOIL_2D_DEFAULT_CONFIG, "reynolds_number": 5e3})
    from .envs.tcf import LARGE_TCF_3D_DEFAULT_CONFIG, SMALL_TCF_3D_DEFAULT_CONFIG, TCF3DBothEnv, TCF3DBottomEnv
    # fluidgym/__init__.py:216-300
    for size, cfg in (("Small", SMALL_TCF_3D_DEFAULT_CONFIG), ("Large", LARGE_TCF_3D_DEFAULT_CONFIG)):
        for walls, entry in (("bottom", TCF3DBottomEnv), ("both", TCF3DBothEnv)):
            for level, re_w in (("easy", 180), ("medium", 330), ("hard", 550)):
                register(f"TCF{size}3D-{walls}-{level}-v0", entry, **{**cfg, "reynolds_number_wall": re_w})
    from .envs.rbc import RBC_2D_DEFAULT_CONFIG, RBC2DEnv
    # fluidgym/__init__.py:106-157
    register("RBC2D-easy-v0", RBC2DEnv, **{**RBC_2D_DEFAULT_CONFIG, "rayleigh_number": 8e4, "adaptive_cfl": 0.8})
    register("RBC2D-medium-v0", RBC2DEnv, **{**RBC_2D_DEFAULT_CONFIG, "rayleigh_number": 4e5, "adaptive_cfl": 0.5})
    register("RBC2D-hard-v0", RBC2DEnv, **{**RBC_2D_DEFAULT_CONFIG, "rayleigh_number": 8e5, "adaptive_cfl": 0.5})
    register("RBC2D-wide-easy-v0", RBC2DEnv, **{**RBC_2D_DEFAULT_CONFIG, "aspect_ratio": 2, "n_heaters": 24, "rayleigh_number": 8e4})
    register("RBC2D-wide-medium-v0", RBC2DEnv, **{**RBC_2D_DEFAULT_CONFIG, "aspect_ratio": 2, "n_heaters": 24, "rayleigh_number": 4e5,
                                                   "adaptive_cfl": 0.5})
    register("RBC2D-wide-hard-v0", RBC2DEnv, **{**RBC_2D_DEFAULT_CONFIG, "aspect_ratio": 2, "n_heaters": 24, "rayleigh_number": 8e5,
                                                 "adaptive_cfl": 0.5})
    from .envs.rbc3d import RBC_3D_DEFAULT_CONFIG, RBC3DEnv
    # fluidgym/__init__.py:159-214
    for level, ra in (("easy", 6e3), ("medium", 8e3), ("hard", 1e4)):
        register(f"RBC3D-{level}-v0", RBC3DEnv, **{**RBC_3D_DEFAULT_CONFIG, "rayleigh_number": ra, "adaptive_cfl": 0.5})
        register(f"RBC3D-wide-{level}-v0", RBC3DEnv, **{**RBC_3D_DEFAULT_CONFIG, "aspect_ratio": 2, "n_heaters": 16, "rayleigh_number": ra,
                                                        "adaptive_cfl": 0.5})
    from .envs.cylinder3d import CYLINDER_JET_3D_DEFAULT_CONFIG, CylinderJet3DEnv
    # fluidgym/__init__.py:79-101.  Host side verified on the CPU against the reference's env.step, the CUDA path on a B200 against the
    # reference's substep, reset, env.step and gradients (tests/test_gpu_extruded.py); res 8 is part of tests/test_gpu_all_envs.py.
    register("CylinderJet3D-easy-v0", CylinderJet3DEnv, **{**CYLINDER_JET_3D_DEFAULT_CONFIG, "reynolds_number": 100.0, "resolution": 24})
    register("CylinderJet3D-medium-v0", CylinderJet3DEnv, **{**CYLINDER_JET_3D_DEFAULT_CONFIG, "reynolds_number": 250.0, "resolution": 32})
    register("CylinderJet3D-hard-v0", CylinderJet3DEnv, **{**CYLINDER_JET_3D_DEFAULT_CONFIG, "reynolds_number": 500.0, "resolution": 48})
    from .envs.airfoil3d import AIRFOIL_3D_DEFAULT_CONFIG, Airfoil3DEnv
    # fluidgym/__init__.py:333-352.  env.step and gradients pinned on a B200 to the reference's at res_z = 8 (tests/test_gpu_extruded.py,
    # tests/golden/airfoil3d_{env,grad}.npz); 4.5 M cells per environment at the default 96 planes, so not in tests/test_gpu_all_envs.py.
    for level, re in (("easy", 1e3), ("medium", 3e3), ("hard", 5e3)):
        register(f"Airfoil3D-{level}-v0", Airfoil3DEnv, **{**AIRFOIL_3D_DEFAULT_CONFIG, "reynolds_number": re})
    register("CylinderJet2D-hard-v0", CylinderJet2DEnv, **{**CYLINDER_JET_2D_DEFAULT_CONFIG, "reynolds_number": 500.0, "resolution": 32})
