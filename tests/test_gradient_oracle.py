"""CPU: the numpy restatement of the reference's velocity-gradient and Smagorinsky kernels (oracle/gradient_eval.py) against outputs of
the unmodified reference (goldens written on a B200 by oracle/ref_harness.py --gradients)."""
import json
import os
import sys

import numpy as np

from conftest import GOLDEN, ROOT, rel_l2

sys.path.insert(0, os.path.join(ROOT, "oracle"))


def test_velocity_gradients_2d_multiblock(golden):
    """five connected blocks with flipped axes, prescribed faces with the 1.5-cell one-sided difference"""
    import gradient_eval as ge
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
    fx = golden("cyl24_velocity_gradients.npz")
    cd = make_cylinder_domain(24).prepare()
    N = cd.N
    minv = np.asarray(cd.minv).reshape(2, 2, N)                    # minv[(2 i + j) N + g] = Minv[i][j]
    G = ge.velocity_gradients(fx["u"], fx["bvel"], np.asarray(cd.nbr), minv)
    for c in range(2):
        for d in range(2):
            assert rel_l2(G[c, d], fx["grad"][c, d]) < 5e-6, (c, d)


def test_velocity_gradients_and_smagorinsky_viscosity_3d_channel(golden):
    import gradient_eval as ge
    from fluidgym_b200.box3d import Box3DDomain
    g, st, sub = golden("tcf32_geometry.npz"), golden("tcf32_sgs_steps.npz"), golden("tcf32_sgs_substep0.npz")
    nu = json.load(open(os.path.join(GOLDEN, "tcf32_meta.json")))["viscosity"]

    def full(t7):
        T = np.zeros(t7.shape[:-1] + (19,), np.float32)
        for k, c in enumerate((0, 4, 8, 9, 13, 17, 18)):
            T[..., c] = t7[..., k]
        return T
    dom = Box3DDomain(g["vertex"], closed=(False, True, False), viscosity=nu, transforms=full(g["Tdiag"]),
                      btransforms={2: full(g["bT2"]), 3: full(g["bT3"])})
    N = dom.N
    mi = np.asarray(dom.minv).reshape(3, N)
    minv = np.zeros((3, 3, N), np.float32)
    for i in range(3):
        minv[i, i] = mi[i]
    nbr = np.asarray(dom.nbr).reshape(6, N)
    # (1) ComputeSpatialVelocityGradients of the state after the reference's env.step
    G = ge.velocity_gradients(st["env0_u"], np.concatenate([st["env0_bvel2"], st["env0_bvel3"]], axis=1), nbr, minv)
    for c in range(3):
        assert rel_l2(G[c], st["env0_grad"][c]) < 2e-6, c
    # (2) the per-cell viscosity the reference's prep function set before its first substep
    G0 = ge.velocity_gradients(sub["u_in"], np.concatenate([sub["bvel2"], sub["bvel3"]], axis=1), nbr, minv)
    h2 = (1.0 / mi.astype(np.float32) ** 2).max(axis=0)
    damp = ge.van_driest_sqr(dom.cell_centres()[1].reshape(-1), 180.0 * nu, nu)
    visc = ge.smagorinsky_viscosity(G0, h2, 0.1, nu, damp)
    assert rel_l2(visc, sub["visc"]) < 2e-6
    assert float((sub["visc"] / nu).max()) > 1.5                   # the model is active in this golden
