"""Rayleigh-Benard path on the CPU: grid / metrics / oracle against the reference golden trace of
RBC2D-easy-v0 (tests/golden/rbc_*.npz, recorded with oracle/ref_harness.py on a B200), the table
compiler against the oracle, and the host-side helpers (moving windows, heater profile)."""
import numpy as np
import pytest
import torch

import table_eval as te
from conftest import rel_l2
from oracle import Oracle, ell_to_csr


@pytest.fixture(scope="module")
def rbc():
    from fluidgym_b200.envs.rbc_domain import make_rbc_domain
    spec, info = make_rbc_domain()
    return spec, info, spec.prepare()


def test_rbc_geometry_matches_reference(rbc, golden):
    spec, info, cd = rbc
    g = golden("rbc_geometry.npz")
    assert (info["nx"], info["ny"], cd.N, cd.NB) == (96, 61, 5856, 192)
    assert np.abs(cd.T - g["T"]).max() / np.abs(g["T"]).max() < 1e-6
    assert np.abs(cd.bT - g["bT"]).max() / np.abs(g["bT"]).max() < 1e-5
    assert np.abs(cd.alpha01).max() == 0.0          # orthogonal grid: every non-orthogonal table vanishes
    assert np.abs(cd.no_gP).max() == 0 and np.abs(cd.nob_w).max() == 0


def test_rbc_oracle_substep_matches_reference(rbc, golden):
    """Scalar transport + buoyancy + orthogonal-path PISO substep (SIM.py:1471-1563, 1650-1705, 1779-1831)."""
    spec, info, cd = rbc
    fx = golden("rbc_substep0.npz")
    dt = float(fx["dt"][0])
    orc = Oracle.from_compiled(cd, nonortho=False)
    orc.set_scalar(float(cd.scalar_visc), cd.sb_neumann[:cd.NB], fx["sbval_in"])
    val, idx, A = orc.build_C(fx["u_in"], dt, nonortho=False, for_scalar=True)
    cv, ci, cr = ell_to_csr(val, idx)
    assert np.array_equal(ci, fx["Cs_index"]) and np.array_equal(cr, fx["Cs_row"])
    assert rel_l2(cv, fx["Cs_value"]) < 2e-6
    srhs = orc.scalar_rhs(fx["u_in"], fx["T_in"], dt)
    assert rel_l2(srhs, fx["srhs"]) < 2e-6
    # the compiled tables reproduce the same scalar system
    off, diag, r = te.scalar_setup(cd, fx["u_in"], fx["T_in"], cd.bvel0, fx["sbval_in"], dt)
    assert rel_l2(diag, val[:, 0]) < 1e-6 and rel_l2(r, srhs) < 1e-6
    for f in range(4):
        assert rel_l2(off[f], val[:, f + 1]) < 1e-6
    u, p, T = fx["u_in"].astype(np.float32).copy(), fx["p_in"].astype(np.float32).copy(), fx["T_in"].astype(np.float32).copy()
    ures = fx["ures_in"].astype(np.float32).copy()
    bicg, cg, sit = orc.substep_scalar(u, p, T, ures, dt, 1.0)
    assert sit == int(fx["scalar_iters"][0]) and bicg == list(fx["bicg_iters"])
    assert [abs(a - int(b)) <= 2 for a, b in zip(cg, fx["cg_iters"])] == [True, True]
    assert rel_l2(T, fx["T_out"]) < 5e-6
    assert rel_l2(u, fx["u1"]) < 2e-5
    assert rel_l2(p, fx["p1"]) < 2e-5


def test_moving_window_matches_reference_semantics():
    """tests/env_utils/test_obs_extraction.py of the reference: circular windows centred on every agent."""
    from fluidgym_b200.envs.rbc import extract_moving_window_2d
    Y, n_agents, aw, win = 3, 6, 2, 3
    field = torch.arange(Y * n_agents * aw, dtype=torch.float32).reshape(Y, n_agents * aw)
    out = extract_moving_window_2d(field, n_agents, aw, win)
    assert out.shape == (n_agents, Y, win * aw)
    fa = field.view(Y, n_agents, aw)
    for i in range(n_agents):
        agents = [(i - 1) % n_agents, i, (i + 1) % n_agents]
        expect = torch.cat([fa[:, a, :] for a in agents], dim=1)
        assert torch.equal(out[i], expect)
    batched = extract_moving_window_2d(torch.stack([field, field + 100]), n_agents, aw, win)
    assert torch.equal(batched[0], out) and torch.equal(batched[1], out + 100)


def test_heater_profile_matches_reference_boundary_values(golden):
    """rbc_env_2d.py:210-282: zero-mean, clamp, cubic blend.  The recorded bottom-plate temperatures after the
    first env.step are the reference's output for the recorded action."""
    from fluidgym_b200.envs.rbc import RBC2DEnv
    st = golden("rbc_steps.npz")
    env = RBC2DEnv.__new__(RBC2DEnv)
    env.heater_width, env.nx, env.n_heaters = 8, 96, 12
    a = torch.from_numpy(st["actions"][0].reshape(1, 12))
    ctrl = RBC2DEnv._action_to_control(env, a)[0].numpy()
    assert np.abs(ctrl - st["env0_sbval"][:96]).max() < 1e-6
