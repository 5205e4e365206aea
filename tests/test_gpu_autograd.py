"""Differentiability: the CUDA adjoint of the PISO substep (fgb_piso_substep_backward through
fluidgym_b200.autograd) against (i) the float64 numpy specification of the adjoint evaluated on the same
tape-free inputs and (ii) finite differences of the CUDA forward itself."""
import numpy as np
import pytest

import adjoint_eval as ae
from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
    from fluidgym_b200.solver import BatchedPISO
    cd = make_cylinder_domain(8).prepare()
    sol = BatchedPISO(cd, 2, cg_impl=6, advection_tol=1e-7, pressure_tol=1e-7)
    rng = np.random.default_rng(1)
    u = 0.3 * rng.standard_normal((2, cd.N)); u[0] += 1.0
    p0 = 0.1 * rng.standard_normal(cd.N)
    bvel = cd.bvel0[:, :cd.NB].astype(np.float64) + 0.05 * rng.standard_normal((2, cd.NB))
    return cd, sol, u, p0, bvel


def _run(sol, u, p0, bvel, dt, wu, wp):
    from fluidgym_b200.autograd import piso_substep
    tu = torch.tensor(np.stack([u, u]), dtype=torch.float32, device="cuda", requires_grad=True)
    tp = torch.tensor(np.stack([p0, p0]), dtype=torch.float32, device="cuda", requires_grad=True)
    tb = torch.tensor(np.stack([bvel, bvel]), dtype=torch.float32, device="cuda", requires_grad=True)
    uo, po = piso_substep(sol, tu, tp, tb, dt)
    J = (uo * torch.tensor(wu, dtype=torch.float32, device="cuda")).sum(dim=(1, 2)) + (po * torch.tensor(wp, dtype=torch.float32, device="cuda")).sum(dim=1)
    return tu, tp, tb, uo, po, J


def test_forward_record_equals_plain_substep(setup):
    cd, sol, u, p0, bvel = setup
    tu, tp, tb, uo, po, J = _run(sol, u, p0, bvel, 0.01, np.zeros((2, cd.N)), np.zeros(cd.N))
    sol.u.copy_(tu.detach()); sol.p.copy_(tp.detach()); sol.bvel.copy_(tb.detach())
    sol.piso_substep(0.01)
    torch.cuda.synchronize()
    # not bit for bit: the recorded pass follows the reference's differentiable backend (CG without the residual reset every
    # 100 iterations, DIFF.py:527-545), the plain substep its plain backend (SIM.py:1908) -- same system, same tolerance
    def rel(a, b):
        return float((a - b).norm() / b.norm())
    assert rel(uo.detach(), sol.u) < 2e-4 and rel(po.detach(), sol.p) < 2e-3


def test_vjp_matches_numpy_specification(setup):
    """u- and boundary-velocity gradients of a random linear functional of u_out (the pressure solution itself
    is only defined up to the CG tolerance ball, so the functional does not weight p_out here)."""
    cd, sol, u, p0, bvel = setup
    t = ae.T64(cd)
    rng = np.random.default_rng(7)
    wu, wp = rng.standard_normal((2, cd.N)), np.zeros(cd.N)
    tu, tp, tb, uo, po, J = _run(sol, u, p0, bvel, 0.01, wu, wp)
    J.sum().backward()
    torch.cuda.synchronize()
    uo64, po64, tape = ae.substep(t, u, p0, bvel, 0.01)
    assert rel_l2(uo[0].detach().cpu().numpy(), uo64) < 1e-4
    ub, pb, bb = ae.substep_vjp(t, u, p0, bvel, 0.01, tape, wu, wp)
    assert rel_l2(tu.grad[0].cpu().numpy(), ub) < 1e-3        # north_star: gradients within 1e-3 relative
    assert rel_l2(tb.grad[1].cpu().numpy(), bb) < 3e-3
    assert rel_l2(tp.grad[0].cpu().numpy(), pb) < 3e-3
    # both environments carry identical data; scatter-adds are atomic, so only equal up to summation order
    assert torch.allclose(tu.grad[0], tu.grad[1], rtol=1e-4, atol=1e-5)
    print('adjoint vs spec: u_out', rel_l2(uo[0].detach().cpu().numpy(), uo64), 'u_bar', rel_l2(tu.grad[0].cpu().numpy(), ub),
          'bvel_bar', rel_l2(tb.grad[1].cpu().numpy(), bb), 'p_bar', rel_l2(tp.grad[0].cpu().numpy(), pb))


def test_vjp_matches_finite_differences_of_the_cuda_forward(setup):
    cd, sol, u, p0, bvel = setup
    from fluidgym_b200.autograd import piso_substep
    rng = np.random.default_rng(3)
    wu = rng.standard_normal((2, cd.N)); wp = np.zeros(cd.N)
    tu, tp, tb, uo, po, J = _run(sol, u, p0, bvel, 0.01, wu, wp)
    J.sum().backward()
    twu = torch.tensor(wu, dtype=torch.float32, device="cuda")
    for name, base, grad in (("u", tu, tu.grad), ("bvel", tb, tb.grad)):
        d = torch.randn(base.shape[1:], device="cuda", generator=torch.Generator(device="cuda").manual_seed(11))
        d = d / d.norm()
        eps = 2e-2
        vals = []
        for sgn in (+1, -1):
            args = {"u": tu.detach(), "p": tp.detach(), "bvel": tb.detach()}
            args[name if name != "u" else "u"] = (base.detach() + sgn * eps * d).contiguous()
            uo2, po2 = piso_substep(sol, args["u"], args["p"], args["bvel"], 0.01)
            vals.append(float((uo2[0].double() * twu.double()).sum()))
        fd = (vals[0] - vals[1]) / (2 * eps)
        an = float((grad[0].double() * d.double()).sum())
        print('fd', name, fd, an)
        assert abs(fd - an) < 3e-2 * max(abs(fd), abs(an)) + 1e-3, (name, fd, an)


# ---- environment level (examples/interfaces/gradient_based_methods.py) ------------------------------------
@pytest.fixture(scope="module")
def developed_state():
    """A short spin-up of the easy cylinder so that drag/lift respond to the jets."""
    import fluidgym_b200 as fg
    env = fg.make("CylinderJet2D-easy-v0", n_envs=1)
    env.reset(seed=0)
    for _ in range(3):
        env.step(torch.zeros(1, 1, device="cuda"))
    return env.get_state()


def test_differentiable_step_equals_plain_step(developed_state):
    import fluidgym_b200 as fg
    a = torch.full((2, 1), 0.6, device="cuda")
    out = []
    for diff in (False, True):
        env = fg.make("CylinderJet2D-easy-v0", n_envs=2, differentiable=diff)
        env.reset(seed=0)
        st = developed_state
        env.set_state(st["u"][0], st["p"][0], st["bvel"][0], st["last_control"][0])
        obs, r, term, trunc, info = env.step(a)
        out.append((r.detach().cpu().numpy(), env.solver.u.clone(), info["drag"].cpu().numpy()))
    # the differentiable backend of the reference never resets the CG residual and starts every solve from zero (DIFF.py:527-545,
    # SIM.py:1436-1440): its own two backends differ by 1.3e-3 in the reward of this step (-9.93596 plain, -9.94843 differentiable,
    # tests/golden/cyl24_steps.npz / cyl24_grad.npz); the same holds here
    assert np.allclose(out[0][0], out[1][0], rtol=1e-2, atol=1e-3), (out[0][0], out[1][0])
    assert float((out[0][1] - out[1][1]).abs().max()) < 2e-2


def test_reward_gradient_wrt_action_matches_finite_differences(developed_state):
    """d reward / d action through 25 sim steps of the easy cylinder (one env.step), checked against a central
    difference of the non-differentiable environment."""
    import fluidgym_b200 as fg
    st = developed_state

    def run(a0, diff):
        env = fg.make("CylinderJet2D-easy-v0", n_envs=1, differentiable=diff)
        env.reset(seed=0)
        env.set_state(st["u"][0], st["p"][0], st["bvel"][0], st["last_control"][0])
        # tight solves so that the difference quotient is not dominated by the CG tolerance ball
        env.solver.set_options(pressure_tol=1e-7, advection_tol=1e-7)
        a = torch.full((1, 1), a0, device="cuda", requires_grad=diff)
        obs, r, *_ = env.step(a)
        return a, r, env

    a, r, env = run(0.3, True)
    r.sum().backward()
    g = float(a.grad)
    env.detach()
    eps = 0.1
    rp = float(run(0.3 + eps, False)[1])
    rm = float(run(0.3 - eps, False)[1])
    fd = (rp - rm) / (2 * eps)
    print("d reward / d action: autograd", g, "central difference", fd)
    assert abs(g - fd) < 0.05 * max(abs(g), abs(fd)) + 1e-4


def test_vjp_with_several_nonorthogonal_iterations(setup):
    """advect_non_ortho_steps = 2, pressure_non_ortho_steps = 3 (the airfoil runs 2 / 4): CUDA adjoint against the
    float64 specification."""
    cd, sol0, u, p0, bvel = setup
    from fluidgym_b200.solver import BatchedPISO
    sol = BatchedPISO(cd, 2, cg_impl=6, advection_tol=1e-7, pressure_tol=1e-7, advect_non_ortho_steps=2, pressure_non_ortho_steps=3)
    t = ae.T64(cd)
    rng = np.random.default_rng(5)
    wu, wp = rng.standard_normal((2, cd.N)), np.zeros(cd.N)
    tu, tp, tb, uo, po, J = _run(sol, u, p0, bvel, 0.01, wu, wp)
    J.sum().backward()
    torch.cuda.synchronize()
    uo64, po64, tape = ae.substep(t, u, p0, bvel, 0.01, n_adv=2, n_p=3)
    assert rel_l2(uo[0].detach().cpu().numpy(), uo64) < 1e-4
    ub, pb, bb = ae.substep_vjp(t, u, p0, bvel, 0.01, tape, wu, wp)
    print("multi-iteration adjoint vs spec: u_bar", rel_l2(tu.grad[0].cpu().numpy(), ub), "bvel_bar", rel_l2(tb.grad[0].cpu().numpy(), bb),
          "p_bar", rel_l2(tp.grad[0].cpu().numpy(), pb), np.abs(pb).max())
    assert rel_l2(tu.grad[0].cpu().numpy(), ub) < 1e-3
    assert rel_l2(tb.grad[0].cpu().numpy(), bb) < 3e-3
    # the previous pressure only enters through three nested deferred corrections: its gradient is tiny, compare absolutely
    assert np.abs(tp.grad[0].cpu().numpy() - pb).max() < 3e-3 * max(np.abs(ub).max(), 1e-3)
