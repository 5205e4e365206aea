"""RBC3D on the GPU: the D = 3 orthogonal box path with passive scalar + buoyancy (fgb_ortho3_* with the scalar attached)
against the op trace of the unmodified reference (tests/golden/rbc3d_substep*.npz, 16 x 10 x 16 cells) and the batched
RBC3DEnv against the reference's env.step (observations, rewards, Nusselt number)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _full19(t7):
    T = np.zeros(t7.shape[:-1] + (19,), np.float32)
    for k, c in enumerate((0, 4, 8, 9, 13, 17, 18)):
        T[..., c] = t7[..., k]
    return T


def _ref_transforms(golden):
    g = golden("rbc3d_geometry.npz")
    return _full19(g["Tdiag"]), {2: _full19(g["bT2"]), 3: _full19(g["bT3"])}


@pytest.fixture(scope="module")
def setup(golden):
    from fluidgym_b200.box3d import BatchedPISO3D, Box3DDomain
    g = golden("rbc3d_geometry.npz")
    meta = json.load(open(os.path.join(GOLDEN, "rbc3d_meta.json")))
    T, bT = _ref_transforms(golden)
    dom = Box3DDomain(g["vertex"], closed=(False, True, False), viscosity=meta["viscosity"], transforms=T, btransforms=bT)
    sol = BatchedPISO3D(dom, 2, advection_tol=1e-5, pressure_tol=1e-5, non_orthogonal=False)
    sol.attach_scalar(meta["thermal_diffusivity"], 1.0)
    return dom, sol, meta


def _load(sol, fx):
    sol.u.copy_(torch.from_numpy(fx["u_in"]).cuda().unsqueeze(0).expand_as(sol.u))
    sol.p.copy_(torch.from_numpy(fx["p_in"]).cuda().unsqueeze(0).expand_as(sol.p))
    sol.T.copy_(torch.from_numpy(fx["T_in"]).cuda().unsqueeze(0).expand_as(sol.T))
    sol.bvel.zero_()
    nf = fx["sb2"].size
    sb = np.concatenate([fx["sb2"], np.full(nf, float(fx["sb3"][0]), np.float32)])
    sol.sbval.copy_(torch.from_numpy(sb).cuda().unsqueeze(0).expand_as(sol.sbval))
    sol.buffer("ures").copy_(torch.from_numpy(fx["ures_in"]).cuda().unsqueeze(0).expand_as(sol.buffer("ures")))


@pytest.mark.parametrize("s", [0, 1])
def test_ops_match_reference_trace(setup, golden, s):
    dom, sol, meta = setup
    fx = golden(f"rbc3d_substep{s}.npz")
    _load(sol, fx)
    dt = float(fx["dt"][0])
    sol.advect_scalar(dt)
    assert rel_l2(sol.buffer("rhs").reshape(-1)[: dom.N].cpu().numpy(), fx["scalar_rhs"]) < 5e-7
    it = sol.buffer("iters")[0].cpu().numpy()
    assert int(it[7]) == int(fx["bicg_iters"][0])
    assert rel_l2(sol.T[1].cpu().numpy(), fx["T_out"]) < 2e-6
    sol.setup_advection(dt)                               # the buoyancy source comes from the attached scalar
    assert rel_l2(sol.buffer("A")[1].cpu().numpy(), fx["A"]) < 5e-7
    assert rel_l2(sol.buffer("rhs")[1].cpu().numpy(), fx["rhs"]) < 5e-7
    sol.solve_advection(zero_init=False)                  # orthogonal path: started from the previous velocityResult
    it = sol.buffer("iters")[0].cpu().numpy()
    assert list(it[:3]) == list(fx["bicg_iters"][1:])
    assert rel_l2(sol.buffer("ures")[1].cpu().numpy(), fx["ustar"]) < 1e-5
    sol.setup_pressure(dt, with_matrix=True)
    assert rel_l2(sol.buffer("hbya")[0].cpu().numpy(), fx["hbya0"]) < 1e-5
    sol.solve_pressure(zero_init=True, reset_steps=0, slot=0)
    it = sol.buffer("iters")[0].cpu().numpy()
    assert abs(int(it[3]) - int(fx["cg_iters"][0])) <= 2
    assert rel_l2(sol.p[0].cpu().numpy(), fx["p0"]) < 2e-3
    sol.correct_velocity()
    assert rel_l2(sol.buffer("ures")[0].cpu().numpy(), fx["u0"]) < 5e-5


@pytest.mark.parametrize("s", [0, 1])
def test_substep_matches_reference(setup, golden, s):
    dom, sol, meta = setup
    fx = golden(f"rbc3d_substep{s}.npz")
    _load(sol, fx)
    sol.piso_substep(float(fx["dt"][0]))
    torch.cuda.synchronize()
    eu, ep, eT = rel_l2(sol.u[0].cpu().numpy(), fx["u1"]), rel_l2(sol.p[0].cpu().numpy(), fx["p1"]), rel_l2(sol.T[0].cpu().numpy(), fx["T_out"])
    print("rbc3d substep", s, ": u", eu, "p", ep, "T", eT, "iters", sol.buffer("iters")[0].tolist(), "ref", fx["bicg_iters"], fx["cg_iters"])
    assert eu < 5e-6 and ep < 1e-5 and eT < 2e-6          # observed on B200: 2.7e-7, 3.1e-7, 1.5e-7; iteration counts equal
    it = sol.buffer("iters")[0].cpu().numpy()
    assert list(it[:3]) == list(fx["bicg_iters"][1:]) and int(it[7]) == int(fx["bicg_iters"][0]) and list(it[3:5]) == list(fx["cg_iters"])
    assert torch.equal(sol.u[0], sol.u[1]) and torch.equal(sol.T[0], sol.T[1])


def test_env_step_matches_reference(golden):
    """Two env.step calls (2 x 5 solver steps, adaptive substeps, heater actuation on 4 x 4 heaters, 16 agents) from the
    reference's reset state with the reference's actions."""
    import fluidgym_b200 as fg
    st = golden("rbc3d_steps.npz")
    T, bT = _ref_transforms(golden)
    env = fg.make("RBC3D-easy-v0", n_envs=2, n_heaters=4, resolution=4, step_length=0.25, transforms=T, btransforms=bT)
    obs, _ = env.reset(seed=1)
    assert obs["temperature"].shape == (2, 16, 12, 8, 12) and obs["velocity"].shape == (2, 16, 3, 12, 8, 12)
    nf = st["reset_sb2"].size
    env.set_state(st["reset_u"], st["reset_p"], st["reset_T"], sbval=np.concatenate([st["reset_sb2"], np.zeros(nf, np.float32)]),
                  ures=st["reset_ures"])
    obs0 = env._get_local_obs()
    assert np.abs(obs0["temperature"][0].cpu().numpy() - st["reset_obs_temperature"]).max() < 2e-5
    assert np.abs(obs0["velocity"][1].cpu().numpy() - st["reset_obs_velocity"]).max() < 2e-5
    for k in range(2):
        action = torch.from_numpy(st["actions"][k]).cuda().unsqueeze(0).repeat(2, 1, 1)
        obs, reward, term, trunc, info = env.step(action)
        torch.cuda.synchronize()
        s = env.solver
        ref_T, ref_u = st[f"env{k}_T"], st[f"env{k}_u"]
        eT, eu = rel_l2(s.T[0].cpu().numpy(), ref_T), float(np.abs(s.u[0].cpu().numpy() - ref_u).max())
        d_nu = abs(float(info["nusselt"][0]) - float(st[f"step{k}_info_nusselt"]))
        d_r = float(np.abs(reward[1].cpu().numpy() - st[f"step{k}_reward"]).max())
        d_oT = float(np.abs(obs["temperature"][0].cpu().numpy() - st[f"step{k}_obs_temperature"]).max())
        d_ou = float(np.abs(obs["velocity"][0].cpu().numpy() - st[f"step{k}_obs_velocity"]).max())
        print("rbc3d env.step", k, ": substeps", env.last_substeps, "T", eT, "max|du|", eu, "nusselt", d_nu, "reward", d_r, "obs T", d_oT, "obs u", d_ou)
        if k == 0:
            assert np.abs(s.sbval[0, env._bottom].cpu().numpy() - st["env0_sb2"]).max() < 1e-6
        # observed on B200 after 2 x 5 solver steps: T 5.3e-7, |du| 1.5e-7, Nusselt / rewards 1.2e-7, observations 1.1e-6
        assert eT < 1e-5 and eu < 5e-6
        assert d_nu < 1e-5 and d_r < 1e-5              # north_star: rewards within 1e-4
        assert d_oT < 2e-5 and d_ou < 1e-5
    assert reward.shape == (2, 16)


def test_initial_domain_file_written_by_the_reference_drives_reset(golden, tmp_path):
    """load_initial_domain=True with the reference's directory layout initial_domains/<id>/<idx>/<mode>.{json,npz}
    (fluid_env.py:1044-1112): the file written by the unmodified reference becomes the reset state of every environment;
    save_initial_domain writes a file the reader (and the reference) loads back."""
    import shutil
    import fluidgym_b200 as fg
    T, bT = _ref_transforms(golden)
    kw = dict(n_envs=2, n_heaters=4, resolution=4, step_length=0.25, transforms=T, btransforms=bT)
    env = fg.make("RBC3D-easy-v0", load_initial_domain=True, initial_domains_path=str(tmp_path), **kw)
    assert env.initial_domain_id == "rbc_3d_Ra6000.0_Pr0.7_NH4_HW4"
    with pytest.raises(RuntimeError, match="Initial domain not found"):
        env.reset(seed=0)
    d = tmp_path / env.initial_domain_id / "0"
    d.mkdir(parents=True)
    for ext in ("json", "npz"):
        shutil.copy(os.path.join(GOLDEN, f"rbc3d_domain.{ext}"), d / f"train.{ext}")
    obs, _ = env.reset(seed=0, randomize=False)
    ref = golden("rbc3d_domain_state.npz")
    assert np.array_equal(env.solver.T[1].cpu().numpy(), ref["T"]) and np.array_equal(env.solver.u[0].cpu().numpy(), ref["u"])
    obs, reward, *_ = env.step(torch.zeros(2, 16, 1, device="cuda"))
    assert torch.isfinite(reward).all()
    path = env.save_initial_domain(3, env_index=1)
    from fluidgym_b200.domain_io import load_box_domain
    back = load_box_domain(path)
    assert np.array_equal(back["state"]["T"], env.solver.T[1].cpu().numpy())


def test_rbc3d_gradients_match_reference(golden):
    """Reverse mode of the D = 3 box substep with passive scalar + buoyancy against the UNMODIFIED reference run with differentiable=True
    (tools/r02_rbc3d_grad_golden.sh -> tests/golden/rbc3d_grad.npz; 16 x 10 x 16 cells, 2 x 2 heaters of 8 cells, single agent): one env.step = 5 solver
    steps, d sum(reward) / d (action, u0, T0) and the vector-Jacobian product of the outgoing (velocity, temperature) with sin cotangents.
    (With 4-cell heaters -- the grid of the other RBC3D fixtures -- the reference's own action gradient is NaN: its heater blend divides
    by round(0.1 * 4) = 0 in the branch torch.where does not select, rbc_env_3d.py:205-247; the state gradients agree there as well.)"""
    import fluidgym_b200 as fg
    fx = golden("rbc3d_grad.npz")
    T, bT = _ref_transforms(golden)
    env = fg.make("RBC3D-easy-v0", n_envs=1, n_heaters=2, resolution=8, step_length=0.25, use_marl=False, transforms=T, btransforms=bT,
                  differentiable=True)
    env.reset(seed=1)
    env.set_state(fx["pre_u"], fx["pre_p"], fx["pre_T"], sbval=fx["pre_sbval"])
    u0, T0 = env.mark_state_differentiable()
    act = torch.from_numpy(fx["action"]).cuda().reshape(1, 2, 2, 1).clone().requires_grad_(True)
    obs, reward, term, trunc, info = env.step(act)
    g_a, g_u, g_T = torch.autograd.grad(reward.sum(), [act, u0, T0], retain_graph=True)
    u1, T1 = env._dstate[0], env._dstate[2]
    cu = torch.sin(0.37 * torch.arange(u1.numel(), device="cuda", dtype=torch.float64)).to(torch.float32).reshape(u1.shape)
    cT = torch.sin(0.37 * torch.arange(T1.numel(), device="cuda", dtype=torch.float64) + 0.3).to(torch.float32).reshape(T1.shape)
    v_a, v_u, v_T = torch.autograd.grad([u1, T1], [act, u0, T0], grad_outputs=[cu, cT])
    torch.cuda.synchronize()

    def rel(a, b):
        a, b = a.detach().cpu().numpy().ravel().astype(np.float64), np.asarray(b, dtype=np.float64).ravel()
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))
    e = dict(reward=float(np.abs(reward.detach().cpu().numpy().ravel() - fx["reward"].ravel()).max()), u=rel(u1[0], fx["post_u"]), T=rel(T1[0], fx["post_T"]),
             dr_da=rel(g_a, fx["dreward_daction"]), dr_du=rel(g_u[0], fx["dreward_du"]), dr_dT=rel(g_T[0], fx["dreward_dT"]),
             vjp_da=rel(v_a, fx["vjp_daction"]), vjp_du=rel(v_u[0], fx["vjp_du"]), vjp_dT=rel(v_T[0], fx["vjp_dT"]), substeps=env.last_substeps)
    print("rbc3d gradients vs reference:", e)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(e, open("gpurun_out/rbc3d_gradients.json", "w"))
    assert e["reward"] < 1e-5 and e["u"] < 1e-4 and e["T"] < 1e-5
    for k in ("dr_da", "dr_du", "dr_dT", "vjp_da", "vjp_du", "vjp_dT"):          # north_star: gradients within 1e-3 relative
        assert e[k] < 1e-3, (k, e[k])
