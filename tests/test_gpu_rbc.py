"""GPU parity of the Rayleigh-Benard path (passive scalar + buoyancy + orthogonal PISO + Nusselt + sensors +
multi-agent windows) against the reference golden data of RBC2D-easy-v0 and the CPU oracle."""
import numpy as np
import pytest

from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _env(n_envs=2, **kw):
    import fluidgym_b200
    return fluidgym_b200.make("RBC2D-easy-v0", n_envs=n_envs, **kw)


def test_rbc_substep_matches_reference(golden):
    fx = golden("rbc_substep0.npz")
    env = _env(3)
    env.set_state(fx["u_in"], fx["p_in"], fx["T_in"], sbval=fx["sbval_in"], ures=fx["ures_in"])
    s = env.solver
    gen = torch.Generator(device="cuda").manual_seed(5)
    s.u[2] += 1e-3 * torch.randn(s.u[2].shape, device="cuda", generator=gen)
    u2_in = s.u[2].cpu().numpy().copy()
    s.piso_substep(float(fx["dt"][0]))
    torch.cuda.synchronize()
    its = s.buffer("iters").cpu().numpy()
    assert its[0, 7] == int(fx["scalar_iters"][0])
    assert list(its[0, :2]) == list(fx["bicg_iters"])
    assert abs(int(its[0, 2]) - int(fx["cg_iters"][0])) <= 2 and abs(int(its[0, 3]) - int(fx["cg_iters"][1])) <= 2
    assert rel_l2(s.T[0].cpu().numpy(), fx["T_out"]) < 5e-6
    assert rel_l2(s.u[1].cpu().numpy(), fx["u1"]) < 2e-5
    assert rel_l2(s.p[0].cpu().numpy(), fx["p1"]) < 2e-5
    assert rel_l2(s.vsrc[0, 1].cpu().numpy(), fx["T_out"]) < 5e-6 and float(s.vsrc[0, 0].abs().max()) == 0.0
    # perturbed batch entry against the CPU oracle on the same input
    from oracle import Oracle
    cd = env.cd
    orc = Oracle.from_compiled(cd, nonortho=False)
    orc.set_scalar(float(cd.scalar_visc), cd.sb_neumann[:cd.NB], fx["sbval_in"])
    u, p, T = u2_in, fx["p_in"].astype(np.float32).copy(), fx["T_in"].astype(np.float32).copy()
    ures = fx["ures_in"].astype(np.float32).copy()
    orc.substep_scalar(u, p, T, ures, float(fx["dt"][0]), 1.0)
    assert rel_l2(s.u[2].cpu().numpy(), u) < 2e-5 and rel_l2(s.T[2].cpu().numpy(), T) < 5e-6


def test_rbc_env_step_matches_reference(golden):
    """env.step from the reference's reset state: 20 adaptive sim steps, Nusselt reward, 48 x 8 sensors."""
    st = golden("rbc_steps.npz")
    env = _env(2)
    env.reset(seed=1)
    env.set_state(st["reset_u"], st["reset_p"], st["reset_T"], sbval=st["reset_sbval"], ures=st["reset_ures"])
    obs0 = env._get_global_obs()
    assert np.abs(obs0["temperature"][0].cpu().numpy() - st["reset_obs_temperature"]).max() < 2e-5
    assert np.abs(obs0["velocity"][1].cpu().numpy() - st["reset_obs_velocity"]).max() < 2e-5
    action = torch.from_numpy(st["actions"][0]).cuda().unsqueeze(0).repeat(2, 1, 1)
    obs, reward, term, trunc, info = env.step(action)
    torch.cuda.synchronize()
    assert np.abs(env.solver.sbval[0].cpu().numpy() - st["env0_sbval"]).max() < 1e-6
    assert rel_l2(env.solver.T[0].cpu().numpy(), st["env0_T"]) < 2e-3
    # the flow is still almost at rest here (|u| ~ 2e-3): the velocity is compared on the absolute scale set by
    # the solver tolerance (||r||/sqrt(N) < 1e-5 per solve, 20+ substeps), not relative to its tiny norm
    assert np.abs(env.solver.u[0].cpu().numpy() - st["env0_u"]).max() < 2e-4
    assert abs(float(info["nusselt"][0]) - float(st["step0_info_nusselt"])) < 2e-3 * abs(float(st["step0_info_nusselt"])) + 1e-4
    assert abs(float(reward[1]) - float(st["step0_reward"][0])) < 2e-3 * abs(float(st["step0_reward"][0])) + 1e-4
    assert obs["temperature"].shape == (2, 8, 48) and obs["velocity"].shape == (2, 2, 8, 48)
    assert np.abs(obs["temperature"][0].cpu().numpy() - st["step0_obs_temperature"]).max() < 5e-3
    assert np.abs(obs["velocity"][0].cpu().numpy() - st["step0_obs_velocity"]).max() < 5e-3


def test_rbc_marl_interface(golden):
    """use_marl=True: one agent per heater, local observation windows and blended local/global rewards
    (tests/envs/test_all_envs.py:78-99 of the reference: reward.shape[0] == n_agents, 'global_reward' in info)."""
    st = golden("rbc_steps.npz")
    env = _env(2, use_marl=True)
    obs, _ = env.reset(seed=3)
    assert env.n_agents == 12
    assert obs["temperature"].shape == (2, 12, 8, 11 * 4) and obs["velocity"].shape == (2, 12, 2, 8, 44)
    env.set_state(st["reset_u"], st["reset_p"], st["reset_T"], sbval=st["reset_sbval"], ures=st["reset_ures"])
    obs, reward, term, trunc, info = env.step(env.sample_action())
    assert reward.shape == (2, 12) and "global_reward" in info and info["global_reward"].shape == (2,)
    assert isinstance(term, bool) and isinstance(trunc, bool)
    # local rewards against a direct torch evaluation of rbc_env_2d.py:328-357
    s = env.solver
    T = s.T[0].reshape(61, 96)
    uy = s.u[0, 1].reshape(61, 96)
    cell = torch.from_numpy(env.cd.det.reshape(61, 96)).cuda()
    from fluidgym_b200.envs.rbc import extract_moving_window_2d
    lT = extract_moving_window_2d(T, 12, 8, 11)
    lu = extract_moving_window_2d(uy, 12, 8, 11)
    lc = cell[:, : 11 * 8].unsqueeze(0)
    nu = 1.0 + (8e4 * 0.7) ** 0.5 * (lu * lT * lc).sum(dim=(1, 2)) / lc.sum(dim=(1, 2))
    assert torch.allclose(env._get_local_rewards()[0], -nu, rtol=1e-4, atol=1e-5)


def test_error_strings_match_reference():
    """tests/env_utils/test_fluid_env.py of the reference."""
    env = _env(1)
    with pytest.raises(RuntimeError, match="Environment must be seeded before sampling actions"):
        env.sample_action()
    with pytest.raises(RuntimeError, match="Environment must be reset before stepping"):
        env.step(torch.zeros(1, 12, 1))


def test_randomized_reset_is_per_environment_and_reproducible():
    """reset(randomize=True) = rbc_env_base.py:336-393 per environment: mirror / periodic shift / noise, then 1-2 time units
    of settling.  Seeded -> reproducible; environments of one batch differ from each other; temperature stays physical."""
    env = _env(3, step_length=0.25)
    env.reset(seed=5, randomize=True)
    T1, u1 = env.solver.T.clone(), env.solver.u.clone()
    env.reset(seed=5, randomize=True)
    assert torch.equal(T1, env.solver.T) and torch.equal(u1, env.solver.u)
    assert torch.isfinite(u1).all() and float(T1.min()) > -0.2 and float(T1.max()) < 1.2
    assert not torch.allclose(T1[0], T1[1]) and not torch.allclose(T1[1], T1[2])
    env.reset(seed=5, randomize=False)
    assert not torch.allclose(T1, env.solver.T)
    obs, reward, *_ = env.step(torch.zeros(3, 12, 1, device="cuda"))
    assert torch.isfinite(reward).all()
