"""Extruded D = 3 path (CylinderJet3D / Airfoil3D) and the opt-in kernels on the GPU.  Every case runs in its OWN PROCESS with a time
limit (a fault in a cooperative kernel cannot take the pytest process down) and is an ordinary strict test: all of them passed on a
B200 in round 2 (profiles/r02_first_run_checks.md).

1. tools/extruded_check.py: substep, reset and a whole env.step of CylinderJet3D (res 8) against the unmodified reference's goldens
   (tests/golden/cyl3d_*; envs/cylinder/jet_cylinder_env_3d.py:399-424), both environments of the batch equal, timing.
2. tests/zz_first_run_worker.py: multi-environment assembly kernels (FGB_ASM_ENVS) and the fused-update cooperative CG
   (FGB_K3_CG_FUSED) -- bit-identity with the default kernels; hook / force / sensor kernels of the extruded environments
   (FGB_X3_HOOKS=cuda) against the torch expressions and a float64 evaluation of them."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_cylinder3d_substep_reset_and_env_step_match_reference(tmp_path):
    out = tmp_path / "extruded_check.json"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "extruded_check.py"), "--json", str(out)], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    print(r.stdout[-4000:])
    print(r.stderr[-2000:], file=sys.stderr)
    try:                                                       # keep the evidence where gpurun brings it back
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "extruded_check.log"), "w") as f:
            f.write(r.stdout + "\n--- stderr ---\n" + r.stderr)
    except OSError:
        pass
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    verdict = json.load(open(out))
    assert verdict["ok"]


@pytest.mark.parametrize("case", ["asm2", "asm4", "asm8", "hooks", "forces", "cg_fused"])
def test_opt_in_kernels(case):
    """tests/zz_first_run_worker.py, one process per case:
    asm<E>  -- k_setup_advection_multi / k_setup_pressure_matrix_multi / k_pressure_div_multi<E> (FGB_ASM_ENVS; one thread = the same
               cell of E environments, tables loaded once) are BIT-IDENTICAL to the default kernels on 11 noisy environments (not a
               multiple of E), per buffer and after a whole substep with 2 + 2 deferred-correction iterations;
    hooks   -- kx3_balance_fluxes / kx3_update_outflow / kx3_max_velocity (FGB_X3_HOOKS=cuda) against the default torch expressions of
               ExtrudedStepping (which the CPU tests pin to the reference's boundary values), both judged by their distance from a
               float64 evaluation (on random +- data the flux sums cancel: comparing the two fp32 paths with each other at 2e-6
               failed in round 1 at 7e-6 absolute -- tolerance, not a bug);
    forces  -- kx3_wall_forces and k_sample_sensors on the extruded layout against the torch expressions (per-plane drag / lift, global
               observation of CylinderJet3D);
    cg_fused -- k3_cg_fused (FGB_K3_CG_FUSED=1: 2 instead of 3 grid.sync per CG iteration) is BIT-IDENTICAL to k3_cg on an extruded substep
               (8 solves of 1 400 - 2 300 iterations) and on the reset projection."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "zz_first_run_worker.py"), case], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:], file=sys.stderr)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]


def test_airfoil3d_env_step_matches_reference(golden):
    """Airfoil3D environment-level parity (VERDICT r01 missing #6): one ``env.step`` (5 solver steps, 56 substeps in the reference) from
    the reference's own reset state at the reduced spanwise resolution res_z = 8 (tests/golden/airfoil3d_env.npz, written by the
    unmodified reference with ``AirfoilEnvBase._res_z = 8``, tools/r02_airfoil3d_golden.sh).  As on the 2-D airfoil the pressure
    solves of both codes end at their iteration cap (reference: CG mean 2 026 / max 4 998 iterations, tolerance 1e-8), so states agree
    to the 1e-3 ... 1e-2 of that regime; forces and reward to a few percent."""
    import numpy as np
    import torch
    import fluidgym_b200 as fg
    fx = golden("airfoil3d_env.npz")
    nz = int(fx["nz"])
    env = fg.make("Airfoil3D-easy-v0", n_envs=1, res_z=nz, n_agents=4, init_from_2d=False)
    env.reset(seed=42)
    s = env.solver
    # our own reset (z-invariant projection of the impulsive start) against the reference's z-average
    ru = torch.from_numpy(fx["reset_u"]).mean(dim=1)
    print("reset: rel L2 of the z-mean in-plane velocity", float((s.u[0].reshape(3, nz, -1).mean(dim=1).cpu()[:2] - ru[:2]).norm() / ru[:2].norm()))
    env.set_state(fx["reset_u"], fx["reset_p"], fx["reset_bvel"], last_control=0.0)
    act = torch.from_numpy(fx["actions"][0]).cuda().reshape(env._zero_action.shape)
    obs, reward, term, trunc, info = env.step(act)
    torch.cuda.synchronize()
    u = s.u[0].reshape(3, nz, -1).cpu().numpy()
    planes = [int(k) for k in fx["env0_planes"]]

    def rel(a, b):
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))
    e = dict(u_planes=rel(u[:, planes], fx["env0_u"]), plane_norms=float(np.abs(np.linalg.norm(u, axis=(0, 2)) / fx["env0_u_plane_norms"] - 1).max()),
             bvel=rel(s.bvel[0].cpu().numpy(), fx["env0_bvel"]),
             reward=abs(float(reward[0]) - float(fx["step0_reward"])), drag=abs(float(info["drag"][0]) / float(fx["step0_info_drag"]) - 1),
             lift=abs(float(info["lift"][0]) / float(fx["step0_info_lift"]) - 1),
             all_cds=float(np.abs(info["all_cds"][0].cpu().numpy() - fx["step0_info_all_cds"]).max()),
             obs_velocity=float(np.abs(obs["velocity"][0].cpu().numpy() - fx["step0_obs_velocity"]).max()),
             obs_pressure=float(np.abs(obs["pressure"][0].cpu().numpy() - fx["step0_obs_pressure"]).max()))
    print("Airfoil3D env.step vs the reference:", e, "substeps", env.last_substeps, "reward", float(reward[0]), float(fx["step0_reward"]))
    try:
        import json
        with open(os.path.join(ROOT, "gpurun_out", "airfoil3d_env_step.json"), "w") as f:
            json.dump({**e, "substeps": int(env.last_substeps), "reward": float(reward[0]), "ref_reward": float(fx["step0_reward"])}, f, indent=1)
    except OSError:
        pass
    assert tuple(obs["velocity"].shape[1:]) == tuple(fx["step0_obs_velocity"].shape)
    assert e["u_planes"] < 2e-2 and e["plane_norms"] < 1e-2 and e["bvel"] < 1e-3
    assert e["drag"] < 5e-2 and e["lift"] < 5e-2 and e["reward"] < 5e-2
    assert e["obs_velocity"] < 2e-2 and e["obs_pressure"] < 2e-2


def test_cylinder3d_gradients_match_reference(golden):
    """Reverse mode of the z-extruded substep against the UNMODIFIED reference run with differentiable=True (tools/r02_cyl3d_grad_golden.sh
    -> tests/golden/cyl3d_grad.npz): CylinderJet3D at its test resolution 8 (5 blocks x 8 planes), one env.step = 25 solver steps with
    1 + 2 x 4 Krylov solves each; d reward / d action, d reward / d u0 and the state vjp with the reference's cotangent."""
    import numpy as np
    import torch
    import fluidgym_b200 as fg
    fx = golden("cyl3d_grad.npz")
    env = fg.make("CylinderJet3D-easy-v0", n_envs=1, resolution=8, n_jets=8, differentiable=True)
    env.seed(42)
    env.set_state(fx["pre_u"], fx["pre_p"], fx["pre_bvel"], last_control=np.zeros((1, 8), np.float32))
    u0 = env.mark_state_differentiable()
    act = torch.from_numpy(fx["action"]).cuda().reshape(1, 8, 1).clone().requires_grad_(True)
    obs, reward, term, trunc, info = env.step(act)
    g_a, g_u = torch.autograd.grad(reward.sum(), [act, u0], retain_graph=True)
    u1 = env._dstate[0]
    cot = torch.from_numpy(fx["cotangent_u"]).cuda().reshape(u1.shape)
    v_a, v_u = torch.autograd.grad([u1], [act, u0], grad_outputs=[cot])
    torch.cuda.synchronize()

    def rel(a, b):
        a, b = a.detach().cpu().numpy().ravel().astype(np.float64), np.asarray(b, dtype=np.float64).ravel()
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))
    e = dict(reward=abs(float(reward.detach().sum()) - float(fx["reward"][0])) / abs(float(fx["reward"][0])), u=rel(u1[0], fx["post_u"]),
             dr_da=rel(g_a, fx["dreward_daction"]), dr_du=rel(g_u[0], fx["dreward_du"]), vjp_da=rel(v_a, fx["vjp_daction"]),
             vjp_du=rel(v_u[0], fx["vjp_du"]), substeps=env.last_substeps, drag=float(info["drag"][0].detach()), ref_drag=float(fx["info_drag"][0]))
    print("cyl3d gradients vs reference:", e)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(e, open("gpurun_out/cyl3d_gradients.json", "w"))
    # north_star: rewards within 1e-4, gradients within 1e-3 relative.  Observed on a B200 (profiles/r02_cyl3d_gradients.json):
    # reward 1.2e-6, u 1.3e-6, d reward / d action 3.5e-5, d reward / d u0 1.6e-4, state vjp 2.9e-6 / 2.2e-6
    assert e["reward"] < 1e-4 and e["u"] < 1e-4
    assert e["dr_da"] < 1e-3 and e["vjp_da"] < 1e-3 and e["dr_du"] < 1e-3 and e["vjp_du"] < 1e-3


def test_extruded_adjoint_equals_the_2d_adjoint_on_a_z_invariant_state(golden):
    """Property: for a z-invariant state (w = 0) and a z-invariant cotangent the reverse mode of the extruded substep must reproduce, plane
    by plane, the reverse mode of the 2-D substep on the same compiled domain (which is pinned to the reference's gradients); the z
    components of the gradients vanish.  Catches in-plane porting errors independently of the golden."""
    import numpy as np
    import torch
    from fluidgym_b200.autograd import piso_substep, piso_substep_extruded
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
    from fluidgym_b200.extruded3d import ExtrudedPISO3D
    from fluidgym_b200.solver import BatchedPISO
    fx = golden("cyl3d_grad.npz")
    cd = make_cylinder_domain(8).prepare()
    nz, kw = 8, dict(corrector_steps=2, advect_non_ortho_steps=1, pressure_non_ortho_steps=4, advection_tol=1e-5, pressure_tol=5e-7, max_iter=5000)
    s3, s2 = ExtrudedPISO3D(cd, nz, 4.0 / nz, 1, **kw), BatchedPISO(cd, 1, cg_impl=6, **kw)
    N2, NB = s3.N2, s3.NB2
    gen = torch.Generator(device="cuda").manual_seed(0)
    u2 = torch.from_numpy(fx["pre_u"][:2, 0]).cuda().reshape(1, 2, N2).contiguous()
    p2 = torch.from_numpy(fx["pre_p"][0]).cuda().reshape(1, N2).contiguous()
    b2 = torch.from_numpy(fx["pre_bvel"][:2, 0]).cuda().reshape(1, 2, NB).contiguous()
    cu2, cp2 = torch.randn(1, 2, N2, device="cuda", generator=gen), torch.randn(1, N2, device="cuda", generator=gen)
    u, p, b = u2.clone().requires_grad_(True), p2.clone().requires_grad_(True), b2.clone().requires_grad_(True)
    g2 = torch.autograd.grad(piso_substep(s2, u, p, b, 0.01), [u, p, b], grad_outputs=[cu2, cp2])
    u3 = torch.zeros(1, 3, nz, N2, device="cuda"); u3[:, :2] = u2[:, :, None, :]
    b3 = torch.zeros(1, 3, nz, NB, device="cuda"); b3[:, :2] = b2[:, :, None, :]
    c3 = torch.zeros(1, 3, nz, N2, device="cuda"); c3[:, :2] = cu2[:, :, None, :]
    u = u3.reshape(1, 3, -1).clone().requires_grad_(True)
    p = p2[:, None, :].expand(1, nz, N2).reshape(1, -1).clone().requires_grad_(True)
    b = b3.clone().requires_grad_(True)
    g3 = torch.autograd.grad(piso_substep_extruded(s3, u, p, b, 0.01), [u, p, b],
                             grad_outputs=[c3.reshape(1, 3, -1), cp2[:, None, :].expand(1, nz, N2).reshape(1, -1).contiguous()])
    torch.cuda.synchronize()
    rel = lambda a, r: float((a.detach() - r.detach()).norm() / r.detach().norm())
    gu3, gb3 = g3[0].view(1, 3, nz, N2), g3[2]
    for k in (0, nz // 2):
        e = (rel(gu3[:, :2, k], g2[0]), rel(gb3[:, :2, k], g2[2]))
        print("plane", k, "d/du, d/dbvel vs the 2-D adjoint:", e)
        assert e[0] < 2e-3 and e[1] < 2e-3                  # observed 1.7e-4 / 1.2e-4 (two independent sets of 9 Krylov solves each way)
    assert float(gu3[:, 2].abs().max()) < 1e-2 * float(gu3[:, :2].abs().max())


def test_airfoil3d_differentiable_step_runs_and_responds_to_the_action(golden):
    """Airfoil3D with differentiable=True (the extruded reverse mode pinned on CylinderJet3D, airfoil jets and flux balance as functional
    torch expressions): one env.step of a single solver step from the reference's reset state at res_z = 8 -- the recorded forward pass
    is finite, d reward / d action and d reward / d u0 are finite and not identically zero.  (No reference gradient golden for this
    family: its differentiable env.step takes the reference several minutes; stated as unpinned in DESIGN.md.)"""
    import numpy as np
    import torch
    import fluidgym_b200 as fg
    fx = golden("airfoil3d_env.npz")
    nz = int(fx["nz"])
    env = fg.make("Airfoil3D-easy-v0", n_envs=1, res_z=nz, n_agents=4, init_from_2d=False, step_length=0.05, differentiable=True)
    env.seed(42)
    env.set_state(fx["reset_u"], fx["reset_p"], fx["reset_bvel"], last_control=0.0)
    assert env.n_sim_steps == 1
    u0 = env.mark_state_differentiable()
    act = torch.from_numpy(fx["actions"][0]).cuda().reshape(env._zero_action.shape).clone().requires_grad_(True)
    obs, reward, term, trunc, info = env.step(act)
    g_a, g_u = torch.autograd.grad(reward.sum(), [act, u0])
    torch.cuda.synchronize()
    print("airfoil3d differentiable step: substeps", env.last_substeps, "reward", float(reward.detach().sum()), "|dR/da|", float(g_a.abs().max()),
          "|dR/du0|", float(g_u.abs().max()))
    assert torch.isfinite(reward).all() and torch.isfinite(g_a).all() and torch.isfinite(g_u).all()
    assert float(g_a.abs().max()) > 0 and float(g_u.abs().max()) > 0


def test_airfoil3d_gradients_match_reference(golden):
    """Airfoil3D reverse mode against the UNMODIFIED reference (differentiable=True, res_z = 8, one env.step = 5 solver steps;
    tools/r02_airfoil3d_grad_golden.sh -> tests/golden/airfoil3d_grad.npz; the reference needs 166 s forward + 221 s backward).  As in 2-D the
    airfoil's pressure solves end at their iteration cap in both codes, so forward states agree to ~1e-3 and gradients to the percent
    level of that regime (the bars below); the well-conditioned paths are pinned at 1e-5 on the cylinder, the channel and RBC3D."""
    import numpy as np
    import torch
    import fluidgym_b200 as fg
    from fluidgym_b200.envs.airfoil_domain import make_airfoil_domain
    fx = golden("airfoil3d_grad.npz")
    nz = int(fx["pre_u"].shape[1])
    env = fg.make("Airfoil3D-easy-v0", n_envs=1, res_z=nz, n_agents=4, init_from_2d=False, differentiable=True)
    env.seed(42)
    env.set_state(fx["pre_u"], fx["pre_p"], fx["pre_bvel"], last_control=0.0)
    u0 = env.mark_state_differentiable()
    act = torch.from_numpy(fx["action"]).cuda().reshape(env._zero_action.shape).clone().requires_grad_(True)
    obs, reward, term, trunc, info = env.step(act)
    g_a, g_u = torch.autograd.grad(reward.sum(), [act, u0], retain_graph=True)
    u1 = env._dstate[0]
    spec = make_airfoil_domain()
    N2 = env.solver.N2
    cot = np.zeros((3, nz, N2), np.float32)
    off = 0
    for bi, blk in enumerate(spec.blocks):                                  # ref_grad_harness.cotangent_like, block by block
        n2 = blk.nx * blk.ny
        cot[:, :, off:off + n2] = np.sin(0.37 * np.arange(3 * nz * n2, dtype=np.float64) + 0.1 * bi).astype(np.float32).reshape(3, nz, n2)
        off += n2
    v_a, v_u = torch.autograd.grad([u1], [act, u0], grad_outputs=[torch.from_numpy(cot).cuda().reshape(u1.shape)])
    torch.cuda.synchronize()

    def rel(a, b):
        a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))
    pl = [int(k) for k in fx["planes"]]
    gu, vu, uu = (x.detach()[0].reshape(3, nz, N2).cpu().numpy() for x in (g_u, v_u, u1))
    e = dict(reward=abs(float(reward.detach().sum()) - float(fx["reward"][0])) / abs(float(fx["reward"][0])), u=rel(uu[:, pl], fx["post_u_planes"]),
             dr_da=rel(g_a.detach().cpu().numpy(), fx["dreward_daction"]), vjp_da=rel(v_a.detach().cpu().numpy(), fx["vjp_daction"]),
             dr_du=rel(gu[:, pl], fx["dreward_du_planes"]), vjp_du=rel(vu[:, pl], fx["vjp_du_planes"]),
             dr_du_norms=rel(np.sqrt((gu.astype(np.float64) ** 2).sum(axis=(0, 2))), fx["dreward_du_norms"]), substeps=env.last_substeps)
    print("airfoil3d gradients vs reference:", e)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(e, open("gpurun_out/airfoil3d_gradients.json", "w"))
    assert e["u"] < 5e-3 and e["reward"] < 5e-2
    assert e["vjp_da"] < 5e-2 and e["vjp_du"] < 5e-2
    assert e["dr_da"] < 0.2 and e["dr_du"] < 0.2
