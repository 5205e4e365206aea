#!/usr/bin/env python
"""Reduce the output of the UNMODIFIED reference, run on a B200 as

    python oracle/ref_harness.py --env TCFSmall3D-both-easy-v0 --tag tcf32_sgs --perturb 0.05 --env-steps 1 --time-steps 0 \
        --trace-substeps 1 --lean --gradients \
        --kw '{"resolution_x_z":32,"resolution_y":33,"init_with_noise":false,"C_smag":0.1,"use_van_driest":true}'

(32 x 33 x 32 channel, Smagorinsky model with van Driest damping, tcf_env.py:441-472) to ``tests/golden/tcf32_sgs_*.npz``.  The grid is
that of ``tcf32_geometry.npz``.  Test infrastructure only."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main(src, tag="tcf32_sgs"):
    L = lambda n: np.load(os.path.join(src, f"{tag}_{n}.npz"))
    tr, st, e0, gr = L("trace"), L("steps"), L("state_step0"), L("gradients")
    N = tr["s0_in_b0_u"][0, 0].size

    def fld(a, c):
        return np.asarray(a)[0].reshape(c, N) if c > 1 else np.asarray(a)[0].reshape(N)

    meta = json.load(open(os.path.join(src, f"{tag}_meta.json")))
    its = [m for m in meta["trace_meta"] if m["substep"] == 0]
    p = "s0_"
    np.savez_compressed(
        os.path.join(HERE, f"{tag}_substep0.npz"), dt=tr[p + "dt"], u_in=fld(tr[p + "in_b0_u"], 3), p_in=fld(tr[p + "in_b0_p"], 1),
        bvel2=tr[p + "in_b0_f2_velocity"].reshape(3, -1), bvel3=tr[p + "in_b0_f3_velocity"].reshape(3, -1),
        src=np.asarray(tr[p + "b0_velocitySource"]).reshape(-1)[:3], visc=fld(tr[p + "in_b0_viscosity"], 1), A=tr[p + "A"],
        rhs=tr[p + "velocityRHS0"].reshape(3, N), ustar=tr[p + "solve0_x"].reshape(3, N), hbya0=tr[p + "pressureRHS0"].reshape(3, N),
        p1=tr[p + "pressureResult1"], u1=tr[p + "velocityResult1"].reshape(3, N),
        bicg_iters=np.array([i[1] for i in its[0]["infos"]]), cg_iters=np.array([its[1]["infos"][0][1], its[2]["infos"][0][1]]))
    fx = {k: st[k] for k in st.files}
    fx.update(env0_u=fld(e0["b0_u"], 3), env0_p=fld(e0["b0_p"], 1), env0_bvel2=e0["b0_f2_velocity"].reshape(3, -1),
              env0_bvel3=e0["b0_f3_velocity"].reshape(3, -1),
              # ComputeSpatialVelocityGradients of the state after the env.step: [list index = component c][channel = direction d][N]
              env0_grad=np.stack([fld(gr[f"b0_d{d}"], 3) for d in range(3)]))
    np.savez_compressed(os.path.join(HERE, f"{tag}_steps.npz"), **fx)
    keep = {k: meta[k] for k in ("env", "seed", "torch", "gpu", "n_sim_steps", "dt", "viscosity", "mean_iters", "max_iters", "n_solves",
                                 "substeps_in_env_steps", "perturb")}
    keep["kwargs"] = {"resolution_x_z": 32, "resolution_y": 33, "init_with_noise": False, "C_smag": 0.1, "use_van_driest": True}
    json.dump(keep, open(os.path.join(HERE, f"{tag}_meta.json"), "w"), indent=1)
    for f in sorted(os.listdir(HERE)):
        if f.startswith(tag):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "..", "gpurun_out", "golden"))
