"""FIRST GPU RUN of the extruded D = 3 path (fluidgym_b200/extruded3d.py): one PISO substep from the state the unmodified
reference traced on CylinderJet3D-easy (resolution 8; tests/golden/cyl3d_substep*.npz) against the reference's result.
    python tools/extruded_check.py
Expected: u within the CG tolerance ball (~1e-5), BiCGStab iterations 3,3,1 / CG iterations within a few percent of the reference."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain  # noqa: E402
from fluidgym_b200.extruded3d import ExtrudedPISO3D  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(np.asarray(b, np.float64)))


def main():
    cd = make_cylinder_domain(8).prepare()
    for s in (0, 1):
        fx = np.load(os.path.join(ROOT, "tests", "golden", f"cyl3d_substep{s}.npz"))
        nz, N2 = fx["A"].shape
        sol = ExtrudedPISO3D(cd, nz, float(fx["hz"][0]), n_envs=2)
        sol.u.copy_(torch.from_numpy(fx["u_in"]).reshape(1, 3, -1).cuda().expand_as(sol.u))
        sol.p.copy_(torch.from_numpy(fx["p_in"]).reshape(1, -1).cuda().expand_as(sol.p))
        sol.bvel.copy_(torch.from_numpy(fx["bvel"]).cuda().unsqueeze(0).expand_as(sol.bvel))
        sol.piso_substep(float(fx["dt"][0]))
        torch.cuda.synchronize()
        u, p = sol.u[0].cpu().numpy().reshape(3, nz, N2), sol.p[0].cpu().numpy().reshape(nz, N2)
        print(json.dumps({"substep": s, "rel_l2_u": rel(u, fx["u1"]), "rel_l2_p": rel(p, fx["p1"]), "ref_bicg_iters": fx["bicg_iters"].tolist(),
                          "ref_cg_iters": fx["cg_iters"].tolist(), "envs_equal": bool(torch.equal(sol.u[0], sol.u[1]))}))


if __name__ == "__main__":
    main()
