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
