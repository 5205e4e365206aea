#!/usr/bin/env python
"""Reduce the output of ``oracle/ref_harness.py --env CylinderJet3D-easy-v0 --tag cyl3d --kw '{"resolution":8,"n_jets":8}'
--env-steps 1 --time-steps 1 --trace-substeps 2`` (the UNMODIFIED reference run on a B200: 5 blocks x 8 z-planes = 15 872
cells, pressure_non_ortho_steps = 4) to ``tests/golden/cyl3d_substep{0,1}.npz``.  Fields are stored in the "planes" layout of
oracle/extruded_eval.py ([C, nz, N2], N2 = block-major 2-D cell index of make_cylinder_domain(8)); the CSR matrices of substep 0
keep the reference's global ordering together with the permutation ``glob`` [nz, N2].  Test infrastructure only."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
f32 = np.float32


def main(src):
    from fluidgym_b200.domain import FIXED
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
    t = np.load(os.path.join(src, "cyl3d_trace.npz"))
    g = np.load(os.path.join(src, "cyl3d_geometry.npz"))
    meta = json.load(open(os.path.join(src, "cyl3d_meta.json")))
    spec = make_cylinder_domain(8)
    cd = spec.prepare()
    nz, N2 = 8, cd.N
    N3 = nz * N2
    for bi, b in enumerate(spec.blocks):
        assert np.array_equal(g[f"b{bi}_vertex"][0][:2, 0], b.vertex), "in-plane vertices differ from the 2-D generator"
    offs = np.concatenate([[0], np.cumsum([b.nx * b.ny for b in spec.blocks])])
    glob = np.zeros((nz, N2), np.int64)
    for bi, b in enumerate(spec.blocks):
        n2 = b.nx * b.ny
        glob[:, offs[bi]:offs[bi + 1]] = nz * offs[bi] + np.arange(nz)[:, None] * n2 + np.arange(n2)[None, :]

    def planes(v, comps=1):
        out = np.asarray(v).reshape(comps, N3)[:, glob.reshape(-1)].reshape(comps, nz, N2)
        return out[0] if comps == 1 else out

    def blocks(prefix, name, comps):
        out = np.zeros((comps, nz, N2), f32)
        for bi in range(len(spec.blocks)):
            out[:, :, offs[bi]:offs[bi + 1]] = t[f"{prefix}b{bi}_{name}"][0].reshape(comps, nz, -1)
        return out

    for s in (0, 1):
        p_ = f"s{s}_"
        bvel = np.zeros((3, nz, cd.NB), f32)
        o = 0
        for bi, b in enumerate(spec.blocks):
            for f in range(4):
                if b.bounds[f].type == FIXED:
                    n = b.size(1 - (f >> 1))
                    v = t[f"{p_}in_b{bi}_f{f}_velocity"][0]
                    bvel[:, :, o:o + n] = v.reshape(3, nz, n) if v.ndim > 1 else np.broadcast_to(v.reshape(3, 1, 1), (3, nz, n))
                    o += n
        its = [m for m in meta["trace_meta"] if m["substep"] == s]
        fx = dict(dt=t[p_ + "dt"], hz=np.array([4.0 / nz], f32), u_in=blocks(p_ + "in_", "u", 3), p_in=blocks(p_ + "in_", "p", 1)[0], bvel=bvel,
                  A=planes(t[p_ + "A"]), rhs=planes(t[p_ + "velocityRHS0"], 3), ustar=planes(t[p_ + "solve0_x"], 3),
                  hbya0=planes(t[p_ + "pressureRHS0"], 3), div0=planes(t[p_ + "pressureRHSdiv0"]), x1=planes(t[p_ + "solve1_x"]),
                  div1=planes(t[p_ + "pressureRHSdiv1"]), p0=planes(t[p_ + "pressureResult0"]), u0=planes(t[p_ + "velocityResult0"], 3),
                  p1=planes(t[p_ + "pressureResult1"]), u1=planes(t[p_ + "velocityResult1"], 3),
                  bicg_iters=np.array([i[1] for i in its[0]["infos"]]), cg_iters=np.array([m["infos"][0][1] for m in its[1:]]))
        if s == 0:
            fx.update(glob=glob.astype(np.int32), C_value=t[p_ + "C_value"], C_index=t[p_ + "C_index"].astype(np.int32),
                      C_row=t[p_ + "C_row"].astype(np.int32), P_value=t[p_ + "P_value0"], P_index=t[p_ + "P_index"].astype(np.int32),
                      P_row=t[p_ + "P_row"].astype(np.int32))
        np.savez_compressed(os.path.join(HERE, f"cyl3d_substep{s}.npz"), **fx)
    # environment level (for the CylinderJet3D environment still to be built): reset state, state after the first env.step with
    # the recorded action (25 solver steps), observations [8 jets, 2 sensor layers, ...], drag / lift / per-jet coefficients
    st, rs, e0 = (np.load(os.path.join(src, f"cyl3d_{n}.npz")) for n in ("steps", "state_reset", "state_step0"))

    def state(z):
        u = np.zeros((3, nz, N2), f32)
        p = np.zeros((nz, N2), f32)
        for bi in range(len(spec.blocks)):
            u[:, :, offs[bi]:offs[bi + 1]] = z[f"b{bi}_u"][0].reshape(3, nz, -1)
            p[:, offs[bi]:offs[bi + 1]] = z[f"b{bi}_p"][0].reshape(nz, -1)
        bv = np.zeros((3, nz, cd.NB), f32)
        o = 0
        for bi, b in enumerate(spec.blocks):
            for f in range(4):
                if b.bounds[f].type == FIXED:
                    n = b.size(1 - (f >> 1))
                    v = z[f"b{bi}_f{f}_velocity"][0]
                    bv[:, :, o:o + n] = v.reshape(3, nz, n) if v.ndim > 1 else np.broadcast_to(v.reshape(3, 1, 1), (3, nz, n))
                    o += n
        return u, p, bv
    ru, rp, rb = state(rs)
    eu, ep, eb = state(e0)
    env_fx = {k: st[k] for k in st.files}
    env_fx.update(reset_u=ru, reset_p=rp, reset_bvel=rb, reset_obs_velocity=rs["obs_velocity"], reset_obs_pressure=rs["obs_pressure"],
                  env0_u=eu, env0_p=ep, env0_bvel=eb)
    np.savez_compressed(os.path.join(HERE, "cyl3d_env.npz"), **env_fx)
    keep = {k: meta[k] for k in ("env", "seed", "torch", "gpu", "n_sim_steps", "dt", "viscosity", "timing", "mean_iters", "max_iters",
                                 "n_solves", "substeps_in_env_steps", "reset_seconds")}
    keep["kw"] = {"resolution": 8, "n_jets": 8}
    json.dump(keep, open(os.path.join(HERE, "cyl3d_meta.json"), "w"), indent=1)
    for f in sorted(os.listdir(HERE)):
        if f.startswith("cyl3d"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "..", "gpurun_out", "cyl3d"))
