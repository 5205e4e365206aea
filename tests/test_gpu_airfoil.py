"""Airfoil2D-medium (46 806 cells, 6 blocks; 2 advection / 4 pressure non-orthogonal iterations, tolerances
1e-6 / 1e-7) and CylinderRot2D on the GPU against fixtures recorded from the unmodified reference."""
import numpy as np
import pytest

from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(airfoil):
    from fluidgym_b200.envs.airfoil import Airfoil2DEnv
    e = Airfoil2DEnv(n_envs=2, compiled=airfoil)
    return e


def test_predictor_ops_match_reference(env, golden):
    fx = golden("airfoil_substep0.npz")
    s = env.solver
    dt = float(fx["dt"][0])
    for dst, src in ((s.u, fx["u_in"]), (s.p, fx["presres_in"]), (s.bvel, fx["bvel_in"])):
        dst.copy_(torch.from_numpy(np.ascontiguousarray(src)).cuda().unsqueeze(0).expand_as(dst))
    s.setup_advection(dt)
    assert rel_l2(s.buffer("A")[1].cpu().numpy(), fx["A"]) < 2e-6
    assert rel_l2(s.buffer("rhs")[1].cpu().numpy(), fx["rhs0"]) < 2e-6
    s.solve_advection(zero_init=True)
    s.setup_advection(dt, ures=s.buffer("ures"))
    # the second deferred-correction RHS amplifies the tolerance ball of the first solve (tol 1e-6) by 1/dt-sized terms
    assert rel_l2(s.buffer("rhs")[1].cpu().numpy(), fx["rhs1"]) < 2e-5
    s.solve_advection(zero_init=False)
    assert rel_l2(s.buffer("ures")[1].cpu().numpy(), fx["ustar"]) < 1e-5
    it = s.buffer("iters")[0].cpu().numpy()
    assert abs(int(it[0]) - int(fx["bicg_iters"][1][0])) <= 2


def test_substep_matches_reference(env, golden):
    """Whole substep: 2 predictor iterations, 2 correctors x 4 pressure solves.  The reference's CG does not reach
    1e-7 in fp32 on this mesh (it returns its best iterate after 5000 iterations, residual 1e-5..6e-5), so the
    comparison bar is that tolerance ball, not round-off."""
    fx = golden("airfoil_substep0.npz")
    s = env.solver
    for dst, src in ((s.u, fx["u_in"]), (s.p, fx["presres_in"]), (s.bvel, fx["bvel_in"])):
        dst.copy_(torch.from_numpy(np.ascontiguousarray(src)).cuda().unsqueeze(0).expand_as(dst))
    s.piso_substep(float(fx["dt"][0]))
    torch.cuda.synchronize()
    eu, ep = rel_l2(s.u[0].cpu().numpy(), fx["u1"]), rel_l2(s.p[0].cpu().numpy(), fx["p1"])
    it = s.buffer("iters")[0].cpu().numpy()
    rs = s.buffer("resid")[0].cpu().numpy()
    print("airfoil substep: rel err u", eu, "p", ep, "cg iters", it[2:8], "resid", rs[2:8], "ref iters", fx["cg_iters"], fx["cg_resid"])
    assert eu < 2e-2 and ep < 2e-2
    assert torch.equal(s.u[0], s.u[1])


def test_reset_and_step_match_reference(env, golden):
    st = golden("airfoil_steps.npz")
    obs, _ = env.reset(seed=42)
    s = env.solver
    assert obs["velocity"].shape == (2, 209, 2) and obs["pressure"].shape == (2, 209)
    print("reset: u", rel_l2(s.u[0].cpu().numpy(), st["reset_u"]), "bvel", rel_l2(s.bvel[0].cpu().numpy(), st["reset_bvel"]))
    # make_divergence_free = 4 deferred-correction pressure solves of <= 1000 CG iterations on a mesh whose cell
    # volumes span 6 decades: the projection is far from converged in BOTH implementations (remaining divergence
    # 1.2e-4 rms), and the velocity it produces is only determined to a few per cent (the literal CPU oracle differs
    # from the reference by 5.8 %).  What can be pinned is that our state is as divergence-free as the reference's.
    from oracle import Oracle
    import ctypes as C
    cd = env.cd
    orc = Oracle(cd.sizes, cd.btype, cd.bconn, cd.T, cd.bT, st["reset_bvel"], float(cd.visc))

    def div_rms(u):
        u = np.ascontiguousarray(u, np.float32)
        d = np.zeros(cd.N, np.float32)
        orc.lib.orc_div(orc.dom, u.ctypes.data_as(C.POINTER(C.c_float)), d.ctypes.data_as(C.POINTER(C.c_float)))
        return float(np.linalg.norm(d) / np.sqrt(cd.N))
    ours, ref = div_rms(s.u[0].cpu().numpy()), div_rms(st["reset_u"])
    print("reset divergence rms: ours", ours, "reference", ref, "zero field", div_rms(np.zeros((2, cd.N))))
    assert ours < 1.1 * ref
    assert rel_l2(s.u[0].cpu().numpy(), st["reset_u"]) < 0.1
    assert rel_l2(s.bvel[0].cpu().numpy(), st["reset_bvel"]) < 1e-5
    print("reset obs err", np.abs(obs["velocity"][0].cpu().numpy() - st["reset_obs_velocity"]).max(), np.abs(st["reset_obs_velocity"]).max())
    # one env.step (5 sim steps, ~27 substeps) from the reference's own reset state
    env.set_state(st["reset_u"], st["reset_p"], st["reset_bvel"])
    env.last_control.zero_()
    action = torch.tensor(st["actions"][0], device="cuda").reshape(1, 3).repeat(2, 1)
    obs, reward, term, trunc, info = env.step(action)
    torch.cuda.synchronize()
    print("step0: drag", float(info["drag"][0]), float(st["step0_info_drag"]), "lift", float(info["lift"][0]), float(st["step0_info_lift"]),
          "reward", float(reward[0]), float(st["step0_reward"]), "substeps", env.last_substeps,
          "u err", rel_l2(s.u[0].cpu().numpy(), st["env0_u"]), "bvel err", rel_l2(s.bvel[0].cpu().numpy(), st["env0_bvel"]))
    assert rel_l2(s.u[0].cpu().numpy(), st["env0_u"]) < 5e-3
    assert abs(float(info["drag"][0]) - float(st["step0_info_drag"])) < 3e-2 * abs(float(st["step0_info_drag"]))
    assert abs(float(info["lift"][0]) - float(st["step0_info_lift"])) < 3e-2 * abs(float(st["step0_info_lift"]))
    assert abs(float(reward[0]) - float(st["step0_reward"])) < 3e-2 * abs(float(st["step0_reward"]))
    assert np.abs(obs["velocity"][0].cpu().numpy() - st["step0_obs_velocity"]).max() < 5e-3


def test_rotating_cylinder_step_matches_reference(cyl24, golden):
    from fluidgym_b200.envs.cylinder import CylinderRot2DEnv
    st = golden("rot24_steps.npz")
    e = CylinderRot2DEnv(n_envs=2, compiled=cyl24)
    e.reset(seed=42)
    assert rel_l2(e.solver.u[0].cpu().numpy(), st["reset_u"]) < 2e-3
    e.set_state(st["reset_u"], st["reset_p"], st["reset_bvel"])
    action = torch.tensor(st["actions"][0], device="cuda").reshape(1, 1).repeat(2, 1)
    obs, reward, term, trunc, info = e.step(action)
    torch.cuda.synchronize()
    print("rot step0: drag", float(info["drag"][0]), float(st["step0_info_drag"]), "lift", float(info["lift"][0]), float(st["step0_info_lift"]))
    assert rel_l2(e.solver.bvel[0].cpu().numpy(), st["env0_bvel"]) < 1e-4
    assert rel_l2(e.solver.u[0].cpu().numpy(), st["env0_u"]) < 1e-3
    assert abs(float(info["drag"][0]) - float(st["step0_info_drag"])) < 1e-3 * abs(float(st["step0_info_drag"])) + 1e-4
    assert abs(float(info["lift"][0]) - float(st["step0_info_lift"])) < 1e-3 * abs(float(st["step0_info_drag"])) + 1e-4
    assert np.abs(obs["velocity"][0].cpu().numpy() - st["step0_obs_velocity"]).max() < 2e-3


def test_airfoil_differentiable_step_runs_and_gradient_is_finite(airfoil, golden):
    """Config 4 (gradient-based control): reward.backward() through one env.step (5 sim steps, ~60 substeps with 2 + 8 Krylov
    solves each) gives finite gradients of the size the reference reports (tests/golden/airfoil_grad.npz: |d reward / d action|
    ~ 0.05).  A finite-difference check is meaningless here, for the reference as for us: the differentiable backend starts every
    pressure solve from zero without residual reset and stops at "residual rising for 100 iterations" (residuals 2e-4 ... 4e-4,
    gpurun_out/r02/golden/grad_airfoil.log), so the forward pass is not a smooth function of the action at finite-difference
    step sizes (measured: central difference +0.40 vs autograd -0.17 along (1, 0, -1), eps 0.25).  The adjoint kernels of this
    path are checked against the float64 specification (test_gpu_autograd.py::test_vjp_with_several_nonorthogonal_iterations)
    and the gradient itself against the reference's (test_gpu_ref_gradients.py)."""
    from fluidgym_b200.envs.airfoil import Airfoil2DEnv
    st = golden("airfoil_steps.npz")
    e = Airfoil2DEnv(n_envs=1, compiled=airfoil, differentiable=True)
    e.reset(seed=0)
    e.set_state(st["env0_u"], st["env0_p"], st["env0_bvel"])
    a = torch.tensor([[0.5, 0.0, -0.5]], device="cuda", requires_grad=True)
    obs, r, *_ = e.step(a)
    r.sum().backward()
    e.detach()
    g = a.grad.cpu().numpy()
    print("airfoil d reward / d action:", g, "reward", float(r.detach()))
    assert np.isfinite(g).all() and 1e-4 < np.abs(g).max() < 5.0
