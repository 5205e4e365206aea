"""Host side of the RBC3D environment (fluidgym_b200/envs/rbc3d.py) against golden data of the unmodified reference run on
RBC3D-easy-v0 with n_heaters=4, resolution=4 (tests/golden/rbc3d_*.npz, made by oracle/ref_harness.py +
tests/golden/extract_rbc3d_fixtures.py): grid, metrics, rendered-voxel sensor map incl. the 6-of-8-corner splat of the
reference's 3-D kernel, moving windows, heater profile, Nusselt rewards.  No GPU needed: the solver-free parts are driven
through a stand-in object."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from fluidgym_b200.box3d import Box3DDomain
from fluidgym_b200.envs.rbc3d import RBC3DEnv, extract_moving_window_3d, rbc3d_vertex_grid
from fluidgym_b200.sensors import sensor_tables_3d


@pytest.fixture(scope="module")
def stub(golden):
    meta = json.load(open(os.path.join(GOLDEN, "rbc3d_meta.json")))
    e = object.__new__(RBC3DEnv)
    e.n_envs, e.n_heaters, e.heater_width = 1, 4, 4
    e.nx = e.nz = 16
    e.aspect = torch.pi
    e.ny = round(2.0 * 16 / e.aspect)
    e.L = 1.0 * e.aspect
    e.device = torch.device("cpu")
    e.Ra, e.Pr, e.nu_ref = 6e3, 0.7, 0.0
    e.local_obs_window, e.local_reward_weight, e.use_marl = 3, 0.0015, True
    vertex = rbc3d_vertex_grid(e.nx, e.ny, e.L, 1.0, 1.02)
    e.dom = Box3DDomain(vertex, closed=(False, True, False), viscosity=meta["viscosity"])
    e.cell_size = torch.from_numpy(np.ascontiguousarray(e.dom.det))
    return e, meta


def test_grid_and_metrics_match_reference(stub, golden):
    e, meta = stub
    g = golden("rbc3d_geometry.npz")
    assert e.ny == 10 and np.array_equal(e.dom.vertex, g["vertex"])
    T = g["Tdiag"]
    for d in range(3):
        assert (np.abs(e.dom.h[d] - T[..., d]) / T[..., d]).max() < 1e-6
        assert (np.abs(e.dom.minv[d] - T[..., 3 + d]) / T[..., 3 + d]).max() < 2e-6
    assert (np.abs(e.dom.det - T[..., 6]) / T[..., 6]).max() < 2e-6
    assert abs(e.dom.visc - meta["viscosity"]) < 1e-9
    kappa = float(torch.tensor([(6e3 * 0.7) ** -0.5], dtype=torch.float32)[0])
    assert abs(kappa - meta["thermal_diffusivity"]) < 1e-9


def _obs_from_state(e, T, u, p):
    RBC3DEnv._setup_sensors.__wrapped__(e) if hasattr(RBC3DEnv._setup_sensors, "__wrapped__") else RBC3DEnv._setup_sensors(e)
    idx, w = e.sens_idx.numpy(), e.sens_w.numpy()

    def sample(f):                                   # what fgb_sample_sensors_n computes
        f = np.asarray(f, dtype=np.float32).reshape(-1, e.dom.N)
        out = (w[None] * f[:, idx]).sum(axis=1)      # [C, ns]
        nsx, nsy = e.n_sensors_x, e.n_sensors_y
        return torch.from_numpy(out.reshape(1, -1, nsx, nsy, nsx)).permute(0, 1, 4, 3, 2).contiguous()
    wkw = dict(n_agents=4, agent_width=4, n_agents_per_window=3)
    gT, gu, gp = sample(T)[:, 0], sample(u), sample(p)[:, 0]
    return (extract_moving_window_3d(gT, **wkw)[0], torch.stack([extract_moving_window_3d(gu[:, c], **wkw) for c in range(3)], dim=2)[0],
            extract_moving_window_3d(gp, **wkw)[0])


def test_sensor_map_and_windows_match_reference_observations(stub, golden):
    """reset state of the reference -> rendered-voxel sensors -> 16 agents' moving windows == the reference's observation"""
    e, _ = stub
    st = golden("rbc3d_steps.npz")
    oT, ou, op = _obs_from_state(e, st["reset_T"], st["reset_u"], st["reset_p"])
    assert oT.shape == (16, 12, 8, 12) and ou.shape == (16, 3, 12, 8, 12)
    assert np.abs(oT.numpy() - st["reset_obs_temperature"]).max() < 2e-5
    assert np.abs(ou.numpy() - st["reset_obs_velocity"]).max() < 2e-5
    assert np.abs(op.numpy() - st["reset_obs_pressure"]).max() < 2e-5
    # after the first env.step of the reference (its state -> its observation)
    oT, ou, _ = _obs_from_state(e, st["env0_T"], st["env0_u"], st["env0_p"])
    assert np.abs(oT.numpy() - st["step0_obs_temperature"]).max() < 2e-5
    assert np.abs(ou.numpy() - st["step0_obs_velocity"]).max() < 2e-5


def test_heater_profile_matches_reference(stub, golden):
    e, _ = stub
    st = golden("rbc3d_steps.npz")
    a = torch.from_numpy(st["actions"][0]).reshape(1, 4, 4)
    ctrl = RBC3DEnv._action_to_control(e, a)                       # [1, nz, nx]
    assert ctrl.shape == (1, 16, 16)
    assert np.abs(ctrl.reshape(-1).numpy() - st["env0_sb2"]).max() < 1e-6


def test_nusselt_rewards_match_reference(stub, golden):
    e, _ = stub
    st = golden("rbc3d_steps.npz")

    class S:                                          # the two solver fields the reward code reads
        T = torch.from_numpy(st["env0_T"]).reshape(1, -1)
        u = torch.from_numpy(st["env0_u"]).reshape(1, 3, -1)
    e.solver = S
    nu = RBC3DEnv.compute_global_nusselt(e)
    assert abs(float(nu[0]) - float(st["step0_info_nusselt"])) < 1e-5 * max(1.0, abs(float(st["step0_info_nusselt"])))
    local = RBC3DEnv._get_local_rewards(e)
    reward = e.local_reward_weight * local + (1 - e.local_reward_weight) * (e.nu_ref - nu)[:, None]
    assert np.abs(reward[0].numpy() - st["step0_reward"]).max() < 1e-5
    assert abs(float((e.nu_ref - nu)[0]) - float(st["step0_info_global_reward"][0])) < 1e-5


def test_initial_domain_pool_glue_on_a_stand_in_solver(stub, golden, tmp_path):
    """InitialDomains3D (envs/common.py) driven without a GPU: directory layout, per-environment state assignment from the
    reference-written file, save_initial_domain round trip."""
    import shutil
    from fluidgym_b200.domain_io import load_box_domain
    e, _ = stub

    class S:
        has_scalar, kappa = True, 0.0154
        u, p, bvel = torch.zeros(3, 3, 2560), torch.zeros(3, 2560), torch.ones(3, 3, 512)
        T, sbval = torch.zeros(3, 2560), torch.zeros(3, 512)
    e.solver, e.n_envs = S, 3
    e.rayleigh_number, e.prandtl_number = 6e3, 0.7
    e.initial_domains_path = str(tmp_path)
    e._np_rng = np.random.default_rng(0)
    e.__dict__.pop("_domain_pool", None)
    with pytest.raises(RuntimeError, match="Initial domain not found"):
        e._load_initial_domains_on_reset(False)
    d = tmp_path / e.initial_domain_id / "0"
    d.mkdir(parents=True)
    for ext in ("json", "npz"):
        shutil.copy(os.path.join(GOLDEN, f"rbc3d_domain.{ext}"), d / f"train.{ext}")
    assert e._load_initial_domains_on_reset(False) == [0, 0, 0]
    ref = golden("rbc3d_domain_state.npz")
    assert np.array_equal(S.T[2].numpy(), ref["T"]) and np.array_equal(S.u[1].numpy(), ref["u"]) and not S.bvel.any()
    assert np.array_equal(S.sbval[0, :256].numpy(), ref["sb2"])
    S.T[1] += 1.0
    path = e.save_initial_domain(4, env_index=1)
    assert path.endswith(os.path.join(e.initial_domain_id, "4", "train"))
    back = load_box_domain(path)
    assert np.array_equal(back["state"]["T"], S.T[1].numpy()) and np.array_equal(back["vertex"], e.dom.vertex)
    e.n_envs = 1
