"""z-extruded multi-block domains (CylinderJet3D / Airfoil3D): the float32 numpy specification
tests/extruded_eval.py -- 2-D compiled tables + uniform periodic z faces -- against an op trace of the UNMODIFIED reference on
CylinderJet3D-easy-v0 with resolution 8 (tests/golden/cyl3d_substep*.npz, generated on a B200 by oracle/ref_harness.py and
tests/golden/extract_cyl3d_fixtures.py).  This pins the operator the D = 3 non-orthogonal kernels have to implement."""
import numpy as np
import pytest
import scipy.sparse as sp

import extruded_eval as ee
from conftest import rel_l2

f32 = np.float32


@pytest.fixture(scope="module")
def cyl3d():
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
    return make_cylinder_domain(8).prepare()


@pytest.mark.parametrize("s", [0, 1])
def test_extruded_operators_match_reference_trace(cyl3d, golden, s):
    cd = cyl3d
    fx = golden(f"cyl3d_substep{s}.npz")
    dt, hz = float(fx["dt"][0]), float(fx["hz"][0])
    u, p, bvel = fx["u_in"], fx["p_in"], fx["bvel"]
    off, offz, A = ee.assemble(cd, u, bvel, dt, hz)
    assert rel_l2(A, fx["A"]) < 5e-7
    assert rel_l2(ee.adv_rhs(cd, u, u, bvel, dt), fx["rhs"]) < 5e-7           # first deferred-correction iteration reads u itself
    hb = ee.hbya(cd, u, fx["ustar"], off, offz, A, bvel, dt)
    assert rel_l2(hb, fx["hbya0"]) < 1e-6
    # divergence incl. the deferred non-orthogonal pressure term: first iteration reads the incoming pressure, the second one
    # the (mean-free) result of the first solve -- pressure_non_ortho_steps = 4 in 3-D (cylinder_env_base.py:317)
    assert rel_l2(ee.divergence(cd, hb, bvel, hz, pres=p, A=A), fx["div0"]) < 3e-5
    x1 = fx["x1"]
    assert rel_l2(ee.divergence(cd, hb, bvel, hz, pres=x1 - x1.mean(), A=A), fx["div1"]) < 3e-5
    assert rel_l2(ee.correct(cd, hb, fx["p0"], A, hz), fx["u0"]) < 1e-6
    assert list(fx["cg_iters"].shape) == [8]                                # 2 correctors x 4 pressure iterations


def test_extruded_matrices_match_reference_csr(cyl3d, golden):
    cd = cyl3d
    fx = golden("cyl3d_substep0.npz")
    nz, N2 = fx["A"].shape
    N3 = nz * N2
    glob = fx["glob"].astype(np.int64).reshape(-1)
    off, offz, A = ee.assemble(cd, fx["u_in"], fx["bvel"], float(fx["dt"][0]), float(fx["hz"][0]))
    Poff, Poffz, Pd = ee.build_P(cd, A, float(fx["hz"][0]))
    C = sp.csr_matrix((fx["C_value"], fx["C_index"], fx["C_row"]), shape=(N3, N3))
    P = sp.csr_matrix((fx["P_value"], fx["P_index"], fx["P_row"]), shape=(N3, N3))
    assert C.nnz < 7 * N3                                                    # ELL(7) minus the prescribed-boundary faces
    rng = np.random.default_rng(0)
    for _ in range(2):
        x = rng.standard_normal((nz, N2)).astype(f32)
        xg = np.zeros(N3, f32)
        xg[glob] = x.reshape(-1)
        assert rel_l2(ee.spmv(cd, off, offz, A, x), (C @ xg)[glob].reshape(nz, N2)) < 5e-7
        assert rel_l2(ee.spmv(cd, Poff, Poffz, Pd, x), (P @ xg)[glob].reshape(nz, N2)) < 5e-7
