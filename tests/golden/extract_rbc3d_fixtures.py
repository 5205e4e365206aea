#!/usr/bin/env python
"""Reduce the output of ``oracle/ref_harness.py --env RBC3D-easy-v0 --tag rbc3d --kw '{"n_heaters":4,"resolution":4,
"step_length":0.25}' --env-steps 2 --time-steps 1 --trace-substeps 2 --lean`` (the UNMODIFIED reference run on a B200,
16 x 10 x 16 cells, 16 agents) to the small fixtures ``tests/golden/rbc3d_*.npz``.  Test infrastructure only."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DIAG = (0, 4, 8, 9, 13, 17, 18)        # M00, M11, M22, Mi00, Mi11, Mi22, det of the 19-float transform


def main(src):
    L = lambda n: np.load(os.path.join(src, f"rbc3d_{n}.npz"))
    g, tr, rs, st = L("geometry"), L("trace"), L("state_reset"), L("steps")
    e0, e1, s0 = L("state_step0"), L("state_step1"), L("simstep0")
    np.savez_compressed(os.path.join(HERE, "rbc3d_geometry.npz"), vertex=g["b0_vertex"][0], Tdiag=g["b0_transform"][0][..., DIAG],
                        bT2=g["b0_f2_transform"][0].reshape(-1, 19)[:, DIAG], bT3=g["b0_f3_transform"][0].reshape(-1, 19)[:, DIAG])
    N = rs["b0_u"][0, 0].size

    def fld(a, c):
        return np.asarray(a)[0].reshape(c, N) if c > 1 else np.asarray(a)[0].reshape(N)

    meta = json.load(open(os.path.join(src, "rbc3d_meta.json")))
    for s in (0, 1):
        p = f"s{s}_"
        its = [m for m in meta["trace_meta"] if m["substep"] == s]
        np.savez_compressed(
            os.path.join(HERE, f"rbc3d_substep{s}.npz"), dt=tr[p + "dt"], u_in=fld(tr[p + "in_b0_u"], 3), p_in=fld(tr[p + "in_b0_p"], 1),
            T_in=fld(tr[p + "in_b0_s"], 1), sb2=tr[p + "in_b0_f2_scalar"].reshape(-1), sb3=tr[p + "in_b0_f3_scalar"].reshape(-1),
            ures_in=tr[p + "in_velocityResult"].reshape(3, N), scalar_rhs=tr[p + "scalarRHS0"], T_out=tr[p + "solve0_x"],
            A=tr[p + "A"], rhs=tr[p + "velocityRHS0"].reshape(3, N), vsrc=fld(tr[p + "b0_velocitySource"], 3),
            ustar=tr[p + "solve1_x"].reshape(3, N), hbya0=tr[p + "pressureRHS0"].reshape(3, N), div0=tr[p + "pressureRHSdiv0"],
            p0=tr[p + "pressureResult0"], u0=tr[p + "velocityResult0"].reshape(3, N), div1=tr[p + "pressureRHSdiv1"],
            p1=tr[p + "pressureResult1"], u1=tr[p + "velocityResult1"].reshape(3, N),
            bicg_iters=np.array([i[1] for i in its[0]["infos"]] + [i[1] for i in its[1]["infos"]]),
            cg_iters=np.array([its[2]["infos"][0][1], its[3]["infos"][0][1]]))
    fx = {k: st[k] for k in st.files}
    fx.update(reset_u=fld(rs["b0_u"], 3), reset_p=fld(rs["b0_p"], 1), reset_T=fld(rs["b0_s"], 1), reset_sb2=rs["b0_f2_scalar"].reshape(-1),
              reset_ures=rs["velocityResult"].reshape(3, N), reset_obs_temperature=rs["obs_temperature"],
              reset_obs_velocity=rs["obs_velocity"], reset_obs_pressure=rs["obs_pressure"],
              sim0_u=fld(s0["b0_u"], 3), sim0_T=fld(s0["b0_s"], 1), sim0_substeps=s0["substeps_so_far"],
              env0_u=fld(e0["b0_u"], 3), env0_p=fld(e0["b0_p"], 1), env0_T=fld(e0["b0_s"], 1), env0_sb2=e0["b0_f2_scalar"].reshape(-1),
              env1_u=fld(e1["b0_u"], 3), env1_T=fld(e1["b0_s"], 1))
    np.savez_compressed(os.path.join(HERE, "rbc3d_steps.npz"), **fx)
    keep = {k: meta[k] for k in ("env", "seed", "torch", "gpu", "n_sim_steps", "dt", "viscosity", "thermal_diffusivity", "timing",
                                 "mean_iters", "max_iters", "n_solves", "substeps_in_env_steps", "reset_seconds")}
    keep["kw"] = {"n_heaters": 4, "resolution": 4, "step_length": 0.25}
    json.dump(keep, open(os.path.join(HERE, "rbc3d_meta.json"), "w"), indent=1)
    for f in sorted(os.listdir(HERE)):
        if f.startswith("rbc3d"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "..", "gpurun_out", "golden"))
