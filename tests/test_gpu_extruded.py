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
