"""On-disk domain format: files written by the reference's own save_domain() (tests/golden/{cyl24,rbc}_domain.*) load into
the same tables the product compiles from scratch, and our writer round-trips."""
import os

import numpy as np

from conftest import GOLDEN


def test_reference_written_cylinder_domain_loads_into_identical_tables(cyl24_own, tmp_path):
    from fluidgym_b200.domain_io import load_domain, save_domain
    spec_own, cd_own = cyl24_own
    spec, state = load_domain(os.path.join(GOLDEN, "cyl24_domain"))
    assert [b.vertex.shape for b in spec.blocks] == [b.vertex.shape for b in spec_own.blocks]
    for a, b in zip(spec.blocks, spec_own.blocks):
        assert np.array_equal(a.vertex, b.vertex)
        assert [x.type for x in a.bounds] == [x.type for x in b.bounds]
        assert [(x.other, tuple(x.axes)) for x in a.bounds if x.type == 2] == [(x.other, tuple(x.axes)) for x in b.bounds if x.type == 2]
    cd = spec.prepare()
    for name in ("nbr", "fl_comp", "Cd", "Wp", "no_idx", "no_wv", "b_cell", "b_face", "rev"):
        assert np.array_equal(getattr(cd, name), getattr(cd_own, name)), name
    assert state["u"].shape == (2, cd.N) and state["p"].shape == (cd.N,) and state["bvel"].shape == (2, cd.NB)
    assert np.abs(state["u"]).max() > 0.5 and state["T"] is None          # a developed flow after one env.step
    # inflow profile is what the product builds itself; the outflow faces carry the advected values of the saved state
    inflow = slice(0, 24)
    assert np.allclose(state["bvel"][:, inflow], cd_own.bvel0[:, inflow], atol=1e-6)
    # round trip through our writer
    save_domain(spec, state, str(tmp_path / "rt"))
    spec2, state2 = load_domain(str(tmp_path / "rt"))
    for k in ("u", "p", "bvel"):
        assert np.array_equal(state[k], state2[k])
    assert np.array_equal(spec2.prepare().nbr, cd.nbr)


def test_reference_written_rbc_domain_with_scalar(tmp_path):
    from fluidgym_b200.domain_io import load_domain, save_domain
    from fluidgym_b200.envs.rbc_domain import make_rbc_domain
    spec, state = load_domain(os.path.join(GOLDEN, "rbc_domain"))
    own, info = make_rbc_domain(8e4, 0.7, 12, 8, 1.0, False)
    assert np.array_equal(spec.blocks[0].vertex, own.blocks[0].vertex)
    assert abs(spec.viscosity - own.viscosity) < 1e-9 and abs(spec.scalar_viscosity - own.scalar_viscosity) < 1e-9
    cd, cd_own = spec.prepare(), own.prepare()
    for name in ("nbr", "Cd", "Cd_s", "sb_neumann"):
        assert np.array_equal(getattr(cd, name), getattr(cd_own, name)), name
    assert state["T"].shape == (cd.N,) and state["sbval"].shape == (cd.NB,)
    assert state["sbval"][:96].min() >= 0.25 and state["sbval"][96:].max() == 0.0      # heated bottom, cold top
    save_domain(spec, state, str(tmp_path / "rt"))
    _, state2 = load_domain(str(tmp_path / "rt"))
    for k in ("u", "p", "T", "sbval", "bvel"):
        assert np.array_equal(state[k], state2[k])


# ---- D = 3 single-block boxes (RBC3D, TCF) ---------------------------------------------------------------------------
def test_reference_written_3d_domain_file_loads(golden):
    """tests/golden/rbc3d_domain.{json,npz} was written by the UNMODIFIED reference's save_domain() (oracle/ref_harness.py
    --env RBC3D-easy-v0 --save-domain-only, one env.step after reset); rbc3d_domain_state.npz is the reference's own view of
    the same state."""
    import os
    from conftest import GOLDEN
    from fluidgym_b200.domain_io import load_box_domain
    d = load_box_domain(os.path.join(GOLDEN, "rbc3d_domain"))
    st, ref = d["state"], golden("rbc3d_domain_state.npz")
    assert d["closed"] == (False, True, False) and d["vertex"].shape == (3, 17, 11, 17)
    assert np.array_equal(d["vertex"], golden("rbc3d_geometry.npz")["vertex"])
    assert np.array_equal(st["u"], ref["u"]) and np.array_equal(st["p"], ref["p"]) and np.array_equal(st["T"], ref["T"])
    assert st["bvel"].shape == (3, 512) and not st["bvel"].any()
    assert np.array_equal(st["sbval"][:256], ref["sb2"]) and np.all(st["sbval"][256:] == ref["sb3"][0])
    assert abs(d["scalar_viscosity"] - (6e3 * 0.7) ** -0.5) < 1e-7


def test_3d_domain_file_round_trip(tmp_path):
    from fluidgym_b200.domain_io import load_box_domain, save_box_domain
    from fluidgym_b200.envs.rbc3d import rbc3d_vertex_grid
    v = rbc3d_vertex_grid(8, 5, 3.14159, 1.0, 1.02)
    N, NB = 8 * 5 * 8, 2 * 8 * 8
    rng = np.random.default_rng(0)
    st = dict(u=rng.standard_normal((3, N)).astype(np.float32), p=rng.standard_normal(N).astype(np.float32),
              bvel=rng.standard_normal((3, NB)).astype(np.float32), T=rng.random(N).astype(np.float32), sbval=rng.random(NB).astype(np.float32))
    path = str(tmp_path / "box")
    save_box_domain(v, (False, True, False), 0.01, st, path, scalar_viscosity=0.02)
    d = load_box_domain(path)
    assert np.array_equal(d["vertex"], v) and abs(d["viscosity"] - 0.01) < 1e-9
    for k, a in st.items():
        assert np.array_equal(d["state"][k], a), k
    # without a scalar (channel flow)
    save_box_domain(v, (False, True, False), 0.01, dict(u=st["u"], p=st["p"], bvel=st["bvel"]), path)
    d = load_box_domain(path)
    assert d["state"]["T"] is None and d["state"]["sbval"] is None and np.array_equal(d["state"]["bvel"], st["bvel"])


def test_2d_initial_domain_pool_glue_without_a_gpu(tmp_path):
    """InitialDomains (envs/common.py) for the 2-D multi-block families on a stand-in solver: the file written by the reference
    for CylinderJet2D-easy (tests/golden/cyl24_domain.*) is served from the directory layout of the published splits."""
    import os
    import shutil
    import torch
    from conftest import GOLDEN
    from fluidgym_b200.envs.cylinder import CylinderJet2DEnv
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
    e = object.__new__(CylinderJet2DEnv)
    e.spec = make_cylinder_domain(24)
    cd = e.spec.prepare()
    e.reynolds_number, e.resolution = 100.0, 24

    class S:
        has_scalar = False
        u, p, bvel = torch.zeros(2, 2, cd.N), torch.zeros(2, cd.N), torch.zeros(2, 2, cd.NB)
    e.solver, e.n_envs, e.device, e._dstate = S, 2, torch.device("cpu"), None
    e.initial_domains_path = str(tmp_path)
    e._np_rng = np.random.default_rng(0)
    d = tmp_path / e.initial_domain_id / "0"
    d.mkdir(parents=True)
    for ext in ("json", "npz"):
        shutil.copy(os.path.join(GOLDEN, f"cyl24_domain.{ext}"), d / f"train.{ext}")
    assert e._load_initial_domains_on_reset(False) == [0, 0]
    assert float(S.u.abs().max()) > 0.5 and torch.equal(S.u[0], S.u[1]) and float(S.bvel.abs().max()) > 0.5
    path = e.save_initial_domain(7, env_index=1)
    from fluidgym_b200.domain_io import load_domain
    spec, st = load_domain(path)
    assert np.array_equal(st["u"], S.u[1].numpy()) and len(spec.blocks) == 5


def test_domain_statistics_set_the_reward_normalisers(tmp_path):
    """load_domain_statistics (fluid_env.py:1205-1221): domain_statistics.json next to the initial domains -> Stats of velocity
    magnitude / pressure / every metric, and the family's reward normaliser (cd_ref = drag mean; nu_ref = Nusselt median in 2-D,
    mean in 3-D; lift / drag ratio for the airfoil; bottom or total wall stress for the channel)."""
    import json
    import pytest
    from fluidgym_b200.envs.airfoil import Airfoil2DEnv
    from fluidgym_b200.envs.common import STATISTICS_FILENAME, Stats
    from fluidgym_b200.envs.cylinder import CylinderJet2DEnv
    from fluidgym_b200.envs.cylinder3d import CylinderJet3DEnv
    from fluidgym_b200.envs.rbc import RBC2DEnv
    from fluidgym_b200.envs.rbc3d import RBC3DEnv
    from fluidgym_b200.envs.tcf import TCF3DBothEnv, TCF3DBottomEnv

    def st(mean, p50=None):
        return dict(mean=mean, min=0.0, max=9.0, p5=0.1, p25=0.2, p50=mean if p50 is None else p50, p75=0.8, p95=0.9)

    cases = [(CylinderJet2DEnv, "cd_ref", {"drag": st(3.2), "lift": st(0.1)}, 3.2),
             (CylinderJet3DEnv, "cd_ref", {"drag": st(1.3), "lift": st(0.0)}, 1.3),
             (Airfoil2DEnv, "cl_cd_ref", {"drag": st(0.5), "lift": st(1.5)}, 3.0),
             (RBC2DEnv, "nu_ref", {"nusselt": st(4.0, p50=3.5)}, 3.5),
             (RBC3DEnv, "nu_ref", {"nusselt": st(4.0, p50=3.5)}, 4.0),
             (TCF3DBottomEnv, "tau_ref", {"wall_stress": st(2.0), "wall_stress_bottom": st(1.1), "wall_stress_top": st(0.9)}, 1.1),
             (TCF3DBothEnv, "tau_ref", {"wall_stress": st(2.0), "wall_stress_bottom": st(1.1), "wall_stress_top": st(0.9)}, 2.0)]
    for k, (cls, attr, metrics, want) in enumerate(cases):
        env = object.__new__(cls)
        env.initial_domains_path = str(tmp_path)
        env_id = f"case{k}"
        d = tmp_path / env_id
        d.mkdir()
        stats = {"velocity_magnitude": st(0.7), "pressure": st(0.2), **metrics}
        (d / STATISTICS_FILENAME).write_text(json.dumps(stats))
        env.__class__ = type("Stub", (cls,), {"initial_domain_id": env_id})
        out = env.load_domain_statistics()
        assert out == stats and isinstance(env.velocity_stats, Stats) and env.velocity_stats.mean == 0.7 and env.pressure_stats.p95 == 0.9
        assert set(env.metrics_stats) == set(cls.metrics)
        assert abs(getattr(env, attr) - want) < 1e-12, (cls.__name__, attr)
    # missing file: the reference's open() error surfaces
    env = object.__new__(type("Stub", (CylinderJet2DEnv,), {"initial_domain_id": "nowhere"}))
    env.initial_domains_path = str(tmp_path)
    with pytest.raises(FileNotFoundError):
        env.load_domain_statistics()
    # round trip through save_domain_statistics
    path = env.save_domain_statistics({"velocity_magnitude": st(1.0), "pressure": st(0.0), "drag": st(2.5), "lift": st(0.0)})
    assert path.endswith(STATISTICS_FILENAME)
    env.load_domain_statistics()
    assert env.cd_ref == 2.5
