"""Reduces the gradient dumps of oracle/ref_grad_harness.py (gpurun_out/r02/golden/, produced by the UNMODIFIED reference with
differentiable=True on a B200, tools/r02_goldens.sh) and its tight-tolerance trace to the flat fixtures committed next to this script.

    python tests/golden/extract_grad_fixtures.py [gpurun_out/r02/golden]

Layout = layout of the product (see extract_fixtures.py): cells of all blocks concatenated, fields component-major.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from extract_fixtures import _multiblock_helpers  # noqa: E402


def multiblock(src, tag, out, spec, with_vjp=True):
    nb, faces, cells, bfaces = _multiblock_helpers(spec)
    d = np.load(os.path.join(src, f"{tag}_grad.npz"))
    meta = json.load(open(os.path.join(src, f"{tag}_grad_meta.json")))
    fx = dict(action=d["action"], reward=d["reward"], dreward_daction=d["dreward_daction"],
              pre_u=cells(d, "pre_b{}_u", 2), pre_p=d["pre_pressureResult"].ravel().astype(np.float32), pre_bvel=bfaces(d, "pre_b{}_f{}_velocity"),
              post_u=cells(d, "post_b{}_u", 2), dreward_du=cells(d, "dreward_db{}_u", 2),
              info_drag=d["info_drag"] if "info_drag" in d.files else np.zeros(0), info_lift=d["info_lift"] if "info_lift" in d.files else np.zeros(0),
              forward_cg_mean=np.array(meta["forward_iters"]["cg"]["mean"]), forward_cg_max=np.array(meta["forward_iters"]["cg"]["max"]))
    if with_vjp:
        fx.update(vjp_daction=d["vjp_daction"], vjp_du=cells(d, "vjp_db{}_u", 2),
                  cotangent_u=np.concatenate([d[f"cotangent{bi}"].reshape(2, -1) for bi in range(nb)], axis=1).astype(np.float32))
    np.savez_compressed(os.path.join(HERE, f"{out}_grad.npz"), **fx)
    print(out, {k: (v.shape, float(np.abs(v).max()) if v.size else None) for k, v in fx.items()})


def rbc(src, tag="rbc", out="rbc"):
    d = np.load(os.path.join(src, f"{tag}_grad.npz"))

    def sb(pre):
        return np.concatenate([np.broadcast_to(d[pre + "f2_scalar"].ravel(), (96,)), np.broadcast_to(d[pre + "f3_scalar"].ravel(), (96,))]).astype(np.float32)

    fx = dict(action=d["action"], reward=d["reward"], dreward_daction=d["dreward_daction"], vjp_daction=d["vjp_daction"],
              pre_u=d["pre_b0_u"].reshape(2, -1), pre_p=d["pre_pressureResult"].ravel(), pre_T=d["pre_b0_s"].ravel(), pre_sbval=sb("pre_b0_"),
              pre_ures=d["pre_velocityResult"].reshape(2, -1),
              post_u=d["post_b0_u"].reshape(2, -1), post_T=d["post_b0_s"].ravel(),
              dreward_du=d["dreward_db0_u"].reshape(2, -1), dreward_dT=d["dreward_db0_s"].ravel(),
              vjp_du=d["vjp_db0_u"].reshape(2, -1), vjp_dT=d["vjp_db0_s"].ravel(),
              cotangent_u=d["cotangent0"].reshape(2, -1), cotangent_T=d["cotangent1"].ravel())
    np.savez_compressed(os.path.join(HERE, f"{out}_grad.npz"), **{k: np.asarray(v, dtype=np.float32) for k, v in fx.items()})
    print(out, {k: (np.asarray(v).shape, float(np.abs(v).max())) for k, v in fx.items()})


def tight_trace(src, tag="cyl24_t32", out="cyl24_tight"):
    """First two substeps of the reference run with pressure / advection tolerance 1e-7 (its CG then returns the best iterate after
    5000 iterations at residual ~1.6e-7): input state, dt, u / p after the substep, iteration counts."""
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
    spec = make_cylinder_domain(24)
    nb, faces, cells, bfaces = _multiblock_helpers(spec)
    tr = np.load(os.path.join(src, f"{tag}_trace.npz"))
    meta = json.load(open(os.path.join(src, f"{tag}_meta.json")))
    for k in (0, 1):
        pre = f"s{k}_"
        its = [m for m in meta["trace_meta"] if m["substep"] == k]
        fx = dict(dt=tr[pre + "dt"], u_in=cells(tr, pre + "in_b{}_u", 2), bvel_in=bfaces(tr, pre + "in_b{}_f{}_velocity"),
                  presres_in=tr[pre + "in_pressureResult"].ravel(), ures_in=tr[pre + "in_velocityResult"].reshape(2, -1),
                  u1=tr[pre + "velocityResult1"].reshape(2, -1), p1=tr[pre + "pressureResult1"].ravel(),
                  p0=tr[pre + "pressureResult0"].ravel(),
                  cg_iters=np.array([m["infos"][0][1] for m in its if not m["bicg"]]), cg_resid=np.array([m["infos"][0][0] for m in its if not m["bicg"]]),
                  bicg_iters=np.array([[i[1] for i in m["infos"]] for m in its if m["bicg"]]),
                  pressure_tol=np.array(meta["pressure_tol"]), advection_tol=np.array(meta["advection_tol"]))
        np.savez_compressed(os.path.join(HERE, f"{out}_substep{k}.npz"), **fx)
        print(out, k, fx["cg_iters"], fx["cg_resid"], fx["bicg_iters"])


def tcf(src, tag="tcf32", out="tcf32"):
    """3-D channel, 32 x 33 x 32 (tools/r02_tcf_grad_golden.sh; the grid of tcf32_geometry.npz): one env.step = 10 solver steps.
    The cotangent of the state vjp is sin(0.37 i + 0) over the flat index (ref_grad_harness.cotangent_like), not stored."""
    d = np.load(os.path.join(src, f"{tag}_grad.npz"))
    meta = json.load(open(os.path.join(src, f"{tag}_grad_meta.json")))
    fx = dict(action=d["action"].reshape(-1), reward=d["reward"].reshape(-1)[:1], dreward_daction=d["dreward_daction"].reshape(-1),
              vjp_daction=d["vjp_daction"].reshape(-1), pre_u=d["pre_b0_u"].reshape(3, -1), post_u=d["post_b0_u"].reshape(3, -1),
              dreward_du=d["dreward_db0_u"].reshape(3, -1), vjp_du=d["vjp_db0_u"].reshape(3, -1),
              info_wall_stress=d["info_wall_stress"].reshape(-1), forward_cg_mean=np.array(meta["forward_iters"]["cg"]["mean"]),
              forward_cg_n=np.array(meta["forward_iters"]["cg"]["n"]))
    assert np.abs(d["cotangent0"].reshape(-1) - np.sin(0.37 * np.arange(d["cotangent0"].size)).astype(np.float32)).max() < 1e-6
    np.savez_compressed(os.path.join(HERE, f"{out}_grad.npz"), **{k: np.asarray(v, dtype=np.float32) for k, v in fx.items()})
    print(out, {k: (np.asarray(v).shape, float(np.abs(v).max())) for k, v in fx.items()})


def rbc3d(src, tag="rbc3d", out="rbc3d"):
    """RBC3D-easy with n_heaters 2, resolution 8, step_length 0.25, use_marl False (16 x 10 x 16 cells; the grid of rbc3d_geometry.npz): one
    env.step = 5 solver steps.  Cotangents: sin(0.37 i) (velocity), sin(0.37 i + 0.3) (temperature)."""
    d = np.load(os.path.join(src, f"{tag}_grad.npz"))
    meta = json.load(open(os.path.join(src, f"{tag}_grad_meta.json")))
    N = d["pre_b0_s"].size
    nface = d["pre_b0_f2_scalar"].size

    def sb(pre):
        return np.concatenate([np.broadcast_to(d[pre + "f2_scalar"].ravel(), (nface,)), np.broadcast_to(d[pre + "f3_scalar"].ravel(), (nface,))])

    fx = dict(action=d["action"].reshape(-1), reward=d["reward"].reshape(-1), dreward_daction=d["dreward_daction"].reshape(-1),
              vjp_daction=d["vjp_daction"].reshape(-1), pre_u=d["pre_b0_u"].reshape(3, N), pre_p=d["pre_pressureResult"].reshape(N),
              pre_T=d["pre_b0_s"].reshape(N), pre_sbval=sb("pre_b0_"), post_u=d["post_b0_u"].reshape(3, N), post_T=d["post_b0_s"].reshape(N),
              dreward_du=d["dreward_db0_u"].reshape(3, N), dreward_dT=d["dreward_db0_s"].reshape(N), vjp_du=d["vjp_db0_u"].reshape(3, N),
              vjp_dT=d["vjp_db0_s"].reshape(N), forward_cg_n=np.array(meta["forward_iters"]["cg"]["n"]))
    np.savez_compressed(os.path.join(HERE, f"{out}_grad.npz"), **{k: np.asarray(v, dtype=np.float32) for k, v in fx.items()})
    print(out, {k: (np.asarray(v).shape, float(np.abs(v).max())) for k, v in fx.items()})


def cyl3d(src, tag="cyl3d", out="cyl3d"):
    """CylinderJet3D-easy at resolution 8 (5 blocks x 8 planes = 15 872 cells, 8 jets; tools/r02_cyl3d_grad_golden.sh): one env.step =
    25 solver steps.  Fields in the "planes" layout [C, nz, N2] of extract_cyl3d_fixtures.py; the cotangent is stored (per block
    sin(0.37 i + 0.1 block) over the block's flat index)."""
    from fluidgym_b200.domain import FIXED
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
    d = np.load(os.path.join(src, f"{tag}_grad.npz"))
    meta = json.load(open(os.path.join(src, f"{tag}_grad_meta.json")))
    spec = make_cylinder_domain(8)
    cd = spec.prepare()
    nz, N2 = 8, cd.N
    offs = np.concatenate([[0], np.cumsum([b.nx * b.ny for b in spec.blocks])])

    def blocks(fmt, comps):
        o = np.zeros((comps, nz, N2), np.float32)
        for bi in range(len(spec.blocks)):
            o[:, :, offs[bi]:offs[bi + 1]] = d[fmt.format(bi)][0].reshape(comps, nz, -1)
        return o

    def bvel(pre):
        o, k = np.zeros((3, nz, cd.NB), np.float32), 0
        for bi, b in enumerate(spec.blocks):
            for f in range(4):
                if b.bounds[f].type == FIXED:
                    n = b.size(1 - (f >> 1))
                    v = d[f"{pre}b{bi}_f{f}_velocity"][0]
                    o[:, :, k:k + n] = v.reshape(3, nz, n) if v.size == 3 * nz * n else np.broadcast_to(v.reshape(3, 1, -1), (3, nz, n))
                    k += n
        return o

    fx = dict(action=d["action"].reshape(-1), reward=d["reward"].reshape(-1), dreward_daction=d["dreward_daction"].reshape(-1),
              vjp_daction=d["vjp_daction"].reshape(-1), pre_u=blocks("pre_b{}_u", 3), pre_p=blocks("pre_b{}_p", 1)[0], pre_bvel=bvel("pre_"),
              post_u=blocks("post_b{}_u", 3), dreward_du=blocks("dreward_db{}_u", 3), vjp_du=blocks("vjp_db{}_u", 3),
              cotangent_u=blocks("cotangent{}", 3), info_drag=d["info_drag"].reshape(-1), info_lift=d["info_lift"].reshape(-1),
              forward_cg_n=np.array(meta["forward_iters"]["cg"]["n"]), forward_cg_mean=np.array(meta["forward_iters"]["cg"]["mean"]))
    np.savez_compressed(os.path.join(HERE, f"{out}_grad.npz"), **{k: np.asarray(v, dtype=np.float32) for k, v in fx.items()})
    print(out, {k: (np.asarray(v).shape, float(np.abs(v).max())) for k, v in fx.items()})


def airfoil3d(src, tag="airfoil3d", out="airfoil3d"):
    """Airfoil3D-easy at res_z 8 (6 blocks x 8 planes = 374 448 cells, 4 agents x 3 jets; tools/r02_airfoil3d_grad_golden.sh): one
    env.step = 5 solver steps.  The incoming state (the differentiable backend's own reset) is stored completely; of the field gradients
    the planes 0 and 5 and the per-plane norms are kept (4.5 MB per full field)."""
    from fluidgym_b200.envs.airfoil_domain import make_airfoil_domain
    d = np.load(os.path.join(src, f"{tag}_grad.npz"))
    meta = json.load(open(os.path.join(src, f"{tag}_grad_meta.json")))
    spec = make_airfoil_domain()
    cd = spec.prepare()
    nz, N2 = int(meta.get("res_z", 8)), cd.N
    offs = np.concatenate([[0], np.cumsum([b.nx * b.ny for b in spec.blocks])])

    def blocks(fmt, comps):
        o = np.zeros((comps, nz, N2), np.float32)
        for bi in range(len(spec.blocks)):
            o[:, :, offs[bi]:offs[bi + 1]] = d[fmt.format(bi)][0].reshape(comps, nz, -1)
        return o

    from fluidgym_b200.domain import FIXED
    # the incoming state is stored: the reset of the differentiable backend (zero-started projection solves without residual reset)
    # differs from the plain backend's reset state in airfoil3d_env.npz by 27 % in the velocity
    pre_u, pre_p = blocks("pre_b{}_u", 3), blocks("pre_b{}_p", 1)[0]
    pre_bvel, k = np.zeros((3, nz, cd.NB), np.float32), 0
    for bi, b in enumerate(spec.blocks):
        for f in range(4):
            if b.bounds[f].type == FIXED:
                n = b.size(1 - (f >> 1))
                v = d[f"pre_b{bi}_f{f}_velocity"][0]
                pre_bvel[:, :, k:k + n] = v.reshape(3, nz, n) if v.size == 3 * nz * n else np.broadcast_to(v.reshape(3, 1, -1), (3, nz, n))
                k += n
    gu, vu, cot, post = blocks("dreward_db{}_u", 3), blocks("vjp_db{}_u", 3), blocks("cotangent{}", 3), blocks("post_b{}_u", 3)
    fx = dict(action=d["action"].reshape(-1), reward=d["reward"].reshape(-1), dreward_daction=d["dreward_daction"].reshape(-1),
              vjp_daction=d["vjp_daction"].reshape(-1), pre_u=pre_u, pre_p=pre_p, pre_bvel=pre_bvel, planes=np.array([0, 5]), dreward_du_planes=gu[:, [0, 5]], vjp_du_planes=vu[:, [0, 5]],
              dreward_du_norms=np.sqrt((gu.astype(np.float64) ** 2).sum(axis=(0, 2))), vjp_du_norms=np.sqrt((vu.astype(np.float64) ** 2).sum(axis=(0, 2))),
              post_u_planes=post[:, [0, 5]], cot_phase=np.array([0.1 * bi for bi in range(len(spec.blocks))]),
              info_drag=d["info_drag"].reshape(-1), info_lift=d["info_lift"].reshape(-1), forward_cg_n=np.array(meta["forward_iters"]["cg"]["n"]),
              forward_cg_mean=np.array(meta["forward_iters"]["cg"]["mean"]), forward_seconds=np.array(meta["forward_seconds"]),
              backward_seconds=np.array(meta["backward_seconds"]))
    # the cotangent is sin(0.37 i + 0.1 block) over every block's flat index: reproducible from the block sizes, check one block
    b0 = d["cotangent0"].reshape(-1)
    assert np.abs(b0 - np.sin(0.37 * np.arange(b0.size)).astype(np.float32)).max() < 1e-6
    np.savez_compressed(os.path.join(HERE, f"{out}_grad.npz"), **{k: np.asarray(v, dtype=np.float32) for k, v in fx.items()})
    print(out, {k: (np.asarray(v).shape, float(np.abs(v).max())) for k, v in fx.items()})


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02/golden"
    if len(sys.argv) > 2 and sys.argv[2] == "airfoil3d":
        return airfoil3d(src)
    if len(sys.argv) > 2 and sys.argv[2] == "cyl3d":
        return cyl3d(src)
    if len(sys.argv) > 2 and sys.argv[2] == "tcf":
        return tcf(src)
    if len(sys.argv) > 2 and sys.argv[2] == "tcf_sgs":          # the same channel with C_smag = 0.1 + van Driest damping
        return tcf(src, tag="tcf32_sgs", out="tcf32_sgs")
    if len(sys.argv) > 2 and sys.argv[2] == "rbc3d":
        return rbc3d(src)
    from fluidgym_b200.envs.airfoil_domain import make_airfoil_domain
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
    multiblock(src, "cyl24", "cyl24", make_cylinder_domain(24))
    multiblock(src, "cyl24_tight", "cyl24_tight", make_cylinder_domain(24))
    multiblock(src, "airfoil", "airfoil", make_airfoil_domain(), with_vjp=False)
    rbc(src)
    tight_trace(src)


if __name__ == "__main__":
    main()
