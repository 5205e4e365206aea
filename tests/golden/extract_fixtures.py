"""Reduces the raw reference dumps of oracle/ref_harness.py (gpurun_out/golden/, produced by the
UNMODIFIED reference on a B200) to the small flat fixtures committed next to this script.

    python tests/golden/extract_fixtures.py [gpurun_out/golden] [CylinderJet2D_easy_v0]

Layout of the fixtures = layout of the product: cells of all blocks concatenated (x fastest),
fields component-major ``[2, N]``, boundary faces concatenated in (block, face) order ``[2, NB]``.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from fluidgym_b200.domain import FIXED  # noqa: E402
from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain  # noqa: E402


def rbc(src, tag="RBC2D_easy_v0", out="rbc"):
    """Single block [61 x 96]: cells are already in product order; boundary faces = (-y: 96, +y: 96)."""
    g = np.load(os.path.join(src, f"{tag}_geometry.npz"))
    np.savez_compressed(os.path.join(HERE, f"{out}_geometry.npz"), T=g["b0_transform"].reshape(-1, 9),
                        bT=np.concatenate([g["b0_f2_transform"].reshape(-1, 9), g["b0_f3_transform"].reshape(-1, 9)]))
    tr = np.load(os.path.join(src, f"{tag}_trace.npz"))
    meta = json.load(open(os.path.join(src, f"{tag}_meta.json")))

    def sb(d, pre):
        return np.concatenate([np.broadcast_to(d[pre + "f2_scalar"].ravel(), (96,)), np.broadcast_to(d[pre + "f3_scalar"].ravel(), (96,))]).astype(np.float32)

    k = "s0_"
    its = [m for m in meta["trace_meta"] if m["substep"] == 0]
    fx = dict(dt=tr[k + "dt"], u_in=tr[k + "in_b0_u"].reshape(2, -1), p_in=tr[k + "in_pressureResult"].ravel(),
              T_in=tr[k + "in_b0_s"].ravel(), sbval_in=sb(tr, k + "in_b0_"), ures_in=tr[k + "in_velocityResult"].reshape(2, -1),
              Cs_value=tr[k + "Cs_value"], Cs_index=tr[k + "Cs_index"], Cs_row=tr[k + "Cs_row"], srhs=tr[k + "scalarRHS0"],
              T_out=tr[k + "solve0_x"], A=tr[k + "A"], rhs=tr[k + "velocityRHS0"].reshape(2, -1), ustar=tr[k + "solve1_x"].reshape(2, -1),
              hbya0=tr[k + "pressureRHS0"].reshape(2, -1), div0=tr[k + "pressureRHSdiv0"], p0=tr[k + "pressureResult0"],
              u1=tr[k + "velocityResult1"].reshape(2, -1), p1=tr[k + "pressureResult1"],
              scalar_iters=np.array([its[0]["infos"][0][1]]), bicg_iters=np.array([i[1] for i in its[1]["infos"]]),
              cg_iters=np.array([its[2]["infos"][0][1], its[3]["infos"][0][1]]))
    np.savez_compressed(os.path.join(HERE, f"{out}_substep0.npz"), **fx)
    rs = np.load(os.path.join(src, f"{tag}_state_reset.npz"))
    st = np.load(os.path.join(src, f"{tag}_steps.npz"))
    e0 = np.load(os.path.join(src, f"{tag}_state_step0.npz"))
    s0 = np.load(os.path.join(src, f"{tag}_simstep0.npz"))
    fx = {k2: st[k2] for k2 in st.files if not k2.startswith("step2")}
    fx.update(reset_u=rs["b0_u"].reshape(2, -1), reset_p=rs["pressureResult"], reset_T=rs["b0_s"].ravel(), reset_sbval=sb(rs, "b0_"),
              reset_ures=rs["velocityResult"].reshape(2, -1), reset_obs_temperature=rs["obs_temperature"],
              reset_obs_velocity=rs["obs_velocity"], reset_obs_pressure=rs["obs_pressure"],
              sim0_u=s0["b0_u"].reshape(2, -1), sim0_T=s0["b0_s"].ravel(), sim0_substeps=s0["substeps_so_far"],
              env0_u=e0["b0_u"].reshape(2, -1), env0_p=e0["b0_p"].ravel(), env0_T=e0["b0_s"].ravel(), env0_sbval=sb(e0, "b0_"))
    np.savez_compressed(os.path.join(HERE, f"{out}_steps.npz"), **fx)
    keep = {k2: meta[k2] for k2 in ("env", "seed", "gpu", "n_sim_steps", "dt", "viscosity", "thermal_diffusivity", "timing",
                                    "mean_iters", "max_iters", "n_solves", "substeps_in_env_steps")}
    json.dump(keep, open(os.path.join(HERE, f"{out}_meta.json"), "w"), indent=1)


def _multiblock_helpers(spec):
    nb = len(spec.blocks)
    faces = [(bi, f) for bi, b in enumerate(spec.blocks) for f in range(4) if b.bounds[f].type == FIXED]

    def cells(d, key, comps):
        return np.concatenate([d[key.format(bi)].reshape(comps, -1) for bi in range(nb)], axis=1).astype(np.float32)

    def bfaces(d, key):
        outv = []
        for bi, f in faces:
            v = d[key.format(bi, f)].reshape(2, -1)
            n = spec.blocks[bi].size(1 - (f >> 1))
            outv.append(np.broadcast_to(v, (2, n)) if v.shape[1] == 1 else v)
        return np.concatenate(outv, axis=1).astype(np.float32)
    return nb, faces, cells, bfaces


def rotating_cylinder(src, tag="CylinderRot2D_easy_v0", out="rot24"):
    """First env.step of CylinderRot2D-easy from its reset state (same domain as cyl24)."""
    spec = make_cylinder_domain(24)
    nb, faces, cells, bfaces = _multiblock_helpers(spec)
    rs = np.load(os.path.join(src, f"{tag}_state_reset.npz"))
    st = np.load(os.path.join(src, f"{tag}_steps.npz"))
    e0 = np.load(os.path.join(src, f"{tag}_state_step0.npz"))
    fx = {k2: st[k2] for k2 in st.files if k2.startswith("step0") or k2 == "actions"}
    fx.update(reset_u=cells(rs, "b{}_u", 2), reset_p=rs["pressureResult"].ravel(), reset_bvel=bfaces(rs, "b{}_f{}_velocity"),
              env0_u=cells(e0, "b{}_u", 2), env0_bvel=bfaces(e0, "b{}_f{}_velocity"))
    np.savez_compressed(os.path.join(HERE, f"{out}_steps.npz"), **fx)


def airfoil(src, tag="Airfoil2D_medium_v0", out="airfoil"):
    """Airfoil2D-medium: vertices of the six blocks, the first traced substep (2 advection and 2 x 4 pressure
    non-orthogonal iterations) and the first env.step from the reset state."""
    from fluidgym_b200.envs.airfoil_domain import make_airfoil_domain
    spec = make_airfoil_domain()
    nb, faces, cells, bfaces = _multiblock_helpers(spec)
    g = np.load(os.path.join(src, f"{tag}_geometry.npz"))
    np.savez_compressed(os.path.join(HERE, f"{out}_vertices.npz"), **{f"b{bi}": g[f"b{bi}_vertex"][0] for bi in range(nb)})
    T = np.concatenate([g[f"b{bi}_transform"].reshape(-1, 9) for bi in range(nb)])
    bT = np.concatenate([g[f"b{bi}_f{f}_transform"].reshape(-1, 9) for bi, f in faces])
    np.savez_compressed(os.path.join(HERE, f"{out}_geometry.npz"), T=T, bT=bT)
    tr = np.load(os.path.join(src, f"{tag}_trace.npz"))
    meta = json.load(open(os.path.join(src, f"{tag}_meta.json")))
    k = "s0_"
    its = [m for m in meta["trace_meta"] if m["substep"] == 0]
    fx = dict(dt=tr[k + "dt"], u_in=cells(tr, k + "in_b{}_u", 2), p_in=cells(tr, k + "in_b{}_p", 1)[0],
              bvel_in=bfaces(tr, k + "in_b{}_f{}_velocity"), presres_in=tr[k + "in_pressureResult"].ravel(),
              A=tr[k + "A"].ravel(), rhs0=tr[k + "velocityRHS0"].reshape(2, -1), rhs1=tr[k + "velocityRHS1"].reshape(2, -1),
              ustar=tr[k + "solve1_x"].reshape(2, -1), div0=tr[k + "pressureRHSdiv0"].ravel(), div3=tr[k + "pressureRHSdiv3"].ravel(),
              p_c0=tr[k + "pressureResult0"].ravel(), u1=tr[k + "velocityResult1"].reshape(2, -1), p1=tr[k + "pressureResult1"].ravel(),
              bicg_iters=np.array([[i[1] for i in m["infos"]] for m in its[:2]]),
              cg_iters=np.array([m["infos"][0][1] for m in its[2:10]]), cg_resid=np.array([m["infos"][0][0] for m in its[2:10]]))
    np.savez_compressed(os.path.join(HERE, f"{out}_substep0.npz"), **fx)
    rs = np.load(os.path.join(src, f"{tag}_state_reset.npz"))
    st = np.load(os.path.join(src, f"{tag}_steps.npz"))
    e0 = np.load(os.path.join(src, f"{tag}_state_step0.npz"))
    fx = {k2: st[k2] for k2 in st.files if k2.startswith("step0") or k2 == "actions"}
    fx.update(reset_u=cells(rs, "b{}_u", 2), reset_p=rs["pressureResult"].ravel(), reset_bvel=bfaces(rs, "b{}_f{}_velocity"),
              reset_obs_velocity=rs["obs_velocity"], reset_obs_pressure=rs["obs_pressure"],
              env0_u=cells(e0, "b{}_u", 2), env0_p=cells(e0, "b{}_p", 1)[0], env0_bvel=bfaces(e0, "b{}_f{}_velocity"))
    np.savez_compressed(os.path.join(HERE, f"{out}_steps.npz"), **fx)
    keep = {k2: meta[k2] for k2 in ("env", "seed", "torch", "gpu", "n_sim_steps", "dt", "viscosity", "timing",
                                    "mean_iters", "max_iters", "n_solves", "substeps_in_env_steps", "reset_seconds")}
    json.dump(keep, open(os.path.join(HERE, f"{out}_meta.json"), "w"), indent=1)


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden"
    if os.path.exists(os.path.join(src, "Airfoil2D_medium_v0_trace.npz")):
        return airfoil(src)
    if os.path.exists(os.path.join(src, "CylinderRot2D_easy_v0_steps.npz")):
        return rotating_cylinder(src)
    tag = sys.argv[2] if len(sys.argv) > 2 else "CylinderJet2D_easy_v0"
    if os.path.exists(os.path.join(src, "RBC2D_easy_v0_trace.npz")):
        rbc(src)
    out = "cyl24"
    spec = make_cylinder_domain(24)
    nb = len(spec.blocks)
    faces = [(bi, f) for bi, b in enumerate(spec.blocks) for f in range(4) if b.bounds[f].type == FIXED]

    def cells(d, key, comps):
        return np.concatenate([d[key.format(bi)].reshape(comps, -1) for bi in range(nb)], axis=1).astype(np.float32)

    def bfaces(d, key):
        outv = []
        for bi, f in faces:
            v = d[key.format(bi, f)]
            n = spec.blocks[bi].size(1 - (f >> 1))
            v = v.reshape(2, -1)
            outv.append(np.broadcast_to(v, (2, n)) if v.shape[1] == 1 else v)
        return np.concatenate(outv, axis=1).astype(np.float32)

    g = np.load(os.path.join(src, f"{tag}_geometry.npz"))
    T = np.concatenate([g[f"b{bi}_transform"].reshape(-1, 9) for bi in range(nb)])
    bT = np.concatenate([g[f"b{bi}_f{f}_transform"].reshape(-1, 9) for bi, f in faces])
    np.savez_compressed(os.path.join(HERE, f"{out}_geometry.npz"), T=T, bT=bT)

    tr = np.load(os.path.join(src, f"{tag}_trace.npz"))
    meta = json.load(open(os.path.join(src, f"{tag}_meta.json")))
    for s in (0, 1):
        k = f"s{s}_"
        fx = dict(
            dt=tr[k + "dt"], u_in=cells(tr, k + "in_b{}_u", 2), p_in=cells(tr, k + "in_b{}_p", 1)[0],
            bvel_in=bfaces(tr, k + "in_b{}_f{}_velocity"), presres_in=tr[k + "in_pressureResult"].ravel(),
            C_value=tr[k + "C_value"], C_index=tr[k + "C_index"], C_row=tr[k + "C_row"], A=tr[k + "A"].ravel(),
            rhs=tr[k + "velocityRHS0"].reshape(2, -1), ustar=tr[k + "solve0_x"].reshape(2, -1),
            P_value=tr[k + "P_value0"], P_index=tr[k + "P_index"], P_row=tr[k + "P_row"],
            hbya0=tr[k + "pressureRHS0"].reshape(2, -1), div0=tr[k + "pressureRHSdiv0"].ravel(),
            p0=tr[k + "pressureResult0"].ravel(), u0=tr[k + "velocityResult0"].reshape(2, -1),
            hbya1=tr[k + "pressureRHS1"].reshape(2, -1), div1=tr[k + "pressureRHSdiv1"].ravel(),
            p1=tr[k + "pressureResult1"].ravel(), u1=tr[k + "velocityResult1"].reshape(2, -1),
        )
        its = [m for m in meta["trace_meta"] if m["substep"] == s]
        fx["bicg_iters"] = np.array([i[1] for i in its[0]["infos"]])
        fx["cg_iters"] = np.array([its[1]["infos"][0][1], its[2]["infos"][0][1]])
        if s == 0:  # keep the first substep small: inputs + final outputs only
            fx = {k2: fx[k2] for k2 in ("dt", "u_in", "p_in", "bvel_in", "presres_in", "u1", "p1", "bicg_iters", "cg_iters")}
        np.savez_compressed(os.path.join(HERE, f"{out}_substep{s}.npz"), **fx)

    rs = np.load(os.path.join(src, f"{tag}_state_reset.npz"))
    np.savez_compressed(os.path.join(HERE, f"{out}_reset.npz"), u=cells(rs, "b{}_u", 2), p=cells(rs, "b{}_p", 1)[0],
                        bvel=bfaces(rs, "b{}_f{}_velocity"), presres=rs["pressureResult"].ravel(),
                        obs_velocity=rs["obs_velocity"], obs_pressure=rs["obs_pressure"])
    st = np.load(os.path.join(src, f"{tag}_steps.npz"))
    s0 = np.load(os.path.join(src, f"{tag}_simstep0.npz"))
    e0 = np.load(os.path.join(src, f"{tag}_state_step0.npz"))
    e3 = np.load(os.path.join(src, f"{tag}_state_step3.npz"))
    fx = {k2: st[k2] for k2 in st.files}
    fx.update(sim0_u=cells(s0, "b{}_u", 2), sim0_p=cells(s0, "b{}_p", 1)[0],
              env0_u=cells(e0, "b{}_u", 2), env0_p=cells(e0, "b{}_p", 1)[0], env0_bvel=bfaces(e0, "b{}_f{}_velocity"),
              env3_u=cells(e3, "b{}_u", 2), env3_p=cells(e3, "b{}_p", 1)[0])
    np.savez_compressed(os.path.join(HERE, f"{out}_steps.npz"), **fx)
    keep = {k2: meta[k2] for k2 in ("env", "seed", "torch", "gpu", "n_sim_steps", "dt", "viscosity", "timing",
                                    "mean_iters", "max_iters", "n_solves", "substeps_in_env_steps", "reset_seconds")}
    json.dump(keep, open(os.path.join(HERE, f"{out}_meta.json"), "w"), indent=1)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
