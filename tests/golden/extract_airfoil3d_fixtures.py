#!/usr/bin/env python
"""Reduce the output of tools/r02_airfoil3d_golden.sh (the UNMODIFIED reference's Airfoil3D-easy-v0 with the class attribute
``AirfoilEnvBase._res_z = 8`` and ``n_agents = 4``, ``init_from_2d = False``: 6 blocks x 8 z-planes = 374 448 cells, one reset and one
``env.step`` = 5 solver steps / 56 substeps on a B200) to ``tests/golden/airfoil3d_env.npz``.  Layout: "planes" [C, nz, N2] with N2 the
block-major 2-D cell index of make_airfoil_domain().  The reference's reset state is NOT z-invariant (it perturbs the spanwise
direction), so it is stored completely -- it is the input of the test; of the state after the step the planes 0 and 5 (two different
agents) and the per-plane norms are kept.  Test infrastructure only."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
f32 = np.float32


def main(src):
    from fluidgym_b200.domain import FIXED
    from fluidgym_b200.envs.airfoil_domain import make_airfoil_domain
    spec = make_airfoil_domain()
    cd = spec.prepare()
    meta = json.load(open(os.path.join(src, "airfoil3d_meta.json")))
    nz, N2 = int(meta.get("res_z", 8)), cd.N
    offs = np.concatenate([[0], np.cumsum([b.nx * b.ny for b in spec.blocks])])
    g = np.load(os.path.join(src, "airfoil3d_geometry.npz"))
    for bi, b in enumerate(spec.blocks):
        assert np.array_equal(g[f"b{bi}_vertex"][0][:2, 0], b.vertex), "in-plane vertices differ from the 2-D generator"
    st, rs, e0 = (np.load(os.path.join(src, f"airfoil3d_{n}.npz")) for n in ("steps", "state_reset", "state_step0"))

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
                    bv[:, :, o:o + n] = v.reshape(3, nz, n) if v.size == 3 * nz * n else np.broadcast_to(v.reshape(3, 1, -1), (3, nz, n))
                    o += n
        pr = z["pressureResult"].ravel()
        return u, p, bv, pr

    ru, rp, rb, rpr = state(rs)
    eu, ep, eb, _ = state(e0)
    zinv = max(float(np.abs(ru - ru[:, :1]).max()), float(np.abs(rp - rp[:1]).max()), float(np.abs(rb - rb[:, :1]).max()))
    print("reset state: max deviation from z-invariance", zinv)
    fx = {k: st[k] for k in st.files}
    fx.update(nz=np.array(nz), reset_u=ru, reset_p=rp, reset_presres=rpr, reset_bvel=rb, reset_obs_velocity=rs["obs_velocity"], reset_obs_pressure=rs["obs_pressure"],
              env0_planes=np.array([0, 5]), env0_u=eu[:, [0, 5]], env0_p=ep[[0, 5]], env0_bvel=eb,
              env0_u_norm=np.array(np.linalg.norm(eu)), env0_u_plane_norms=np.linalg.norm(eu.reshape(3, nz, -1), axis=(0, 2)))
    np.savez_compressed(os.path.join(HERE, "airfoil3d_env.npz"), **fx)
    keep = {k: meta[k] for k in meta if k in ("env", "seed", "torch", "gpu", "n_sim_steps", "dt", "viscosity", "mean_iters", "max_iters", "n_solves",
                                              "substeps_in_env_steps", "reset_seconds", "res_z")}
    keep["kw"] = {"init_from_2d": False, "n_agents": 4, "res_z": nz}
    json.dump(keep, open(os.path.join(HERE, "airfoil3d_meta.json"), "w"), indent=1)
    print({k: (v.shape, float(np.abs(v).max()) if v.size else None) for k, v in fx.items()})
    print(os.path.getsize(os.path.join(HERE, "airfoil3d_env.npz")))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "..", "gpurun_out", "r02", "airfoil3d"))
