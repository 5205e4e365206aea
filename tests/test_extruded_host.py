"""The per-cell code of the D = 3 extruded-domain kernels (fluidgym_b200/csrc/extruded3_b200.cuh, __host__ __device__) executed
ON THE CPU through tests/cpu_harness/extruded_host.cu and compared with the op trace of the unmodified reference on
CylinderJet3D-easy (tests/golden/cyl3d_substep*.npz) and with the numpy specification oracle/extruded_eval.py.  This verifies
the arithmetic of the kernels without a GPU; their launch glue has not run on a GPU yet."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import extruded_eval as ee
from conftest import ROOT, rel_l2

f32 = np.float32
sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_harness"))
from extruded_standin import bicgstab as _bicgstab, build_harness, cg as _cg, csr as _csr, host_tables  # noqa: E402


@pytest.fixture(scope="module")
def harness():
    return build_harness()


@pytest.fixture(scope="module")
def tables():
    """fgb_tables with HOST pointers into the numpy arrays of the compiled 2-D cylinder domain (resolution 8)"""
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
    cd = make_cylinder_domain(8).prepare()
    t, keep = host_tables(cd)
    return cd, t, keep


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("s", [0, 1])
def test_cell_functions_match_reference_trace_and_specification(harness, tables, golden, s):
    cd, t, _keep = tables
    fx = golden(f"cyl3d_substep{s}.npz")
    nz, N2 = fx["A"].shape
    N3 = nz * N2
    dt, hz = float(fx["dt"][0]), float(fx["hz"][0])
    u = np.ascontiguousarray(fx["u_in"])
    bvel = np.ascontiguousarray(fx["bvel"])
    p_in = np.ascontiguousarray(fx["p_in"])
    coff, A, rhs = np.zeros((6, nz, N2), f32), np.zeros((nz, N2), f32), np.zeros((3, nz, N2), f32)
    harness.xh_setup_advection(C.byref(t), nz, C.c_float(hz), _p(u), _p(u), _p(bvel), C.c_float(dt), _p(coff), _p(A), _p(rhs), 1)
    assert rel_l2(A, fx["A"]) < 5e-7 and rel_l2(rhs, fx["rhs"]) < 5e-7
    off, offz, A_spec = ee.assemble(cd, u, bvel, dt, hz)
    assert rel_l2(coff[:4], off) < 5e-7 and rel_l2(coff[4:], offz) < 5e-7
    poff, pdiag = np.zeros((6, nz, N2), f32), np.zeros((nz, N2), f32)
    harness.xh_pressure_matrix(C.byref(t), nz, C.c_float(hz), _p(A), _p(poff), _p(pdiag))
    Po, Pz, Pd = ee.build_P(cd, A, hz)
    assert rel_l2(poff[:4], Po) < 5e-7 and rel_l2(poff[4:], Pz) < 5e-7 and rel_l2(pdiag, Pd) < 5e-7
    ustar = np.ascontiguousarray(fx["ustar"])
    hb = np.zeros((3, nz, N2), f32)
    harness.xh_hbya(C.byref(t), nz, C.c_float(hz), _p(u), _p(ustar), _p(bvel), _p(coff), _p(A), C.c_float(dt), _p(hb))
    assert rel_l2(hb, fx["hbya0"]) < 1e-6
    div = np.zeros((nz, N2), f32)
    harness.xh_divergence(C.byref(t), nz, C.c_float(hz), _p(hb), _p(bvel), _p(p_in), _p(A), _p(div))
    assert rel_l2(div, fx["div0"]) < 3e-5
    x1 = fx["x1"]
    pm = np.ascontiguousarray((x1 - x1.mean()).astype(f32))
    harness.xh_divergence(C.byref(t), nz, C.c_float(hz), _p(hb), _p(bvel), _p(pm), _p(A), _p(div))
    assert rel_l2(div, fx["div1"]) < 3e-5
    harness.xh_divergence(C.byref(t), nz, C.c_float(hz), _p(hb), _p(bvel), None, _p(A), _p(div))
    assert rel_l2(div, ee.divergence(cd, hb, bvel, hz)) < 1e-5
    p0 = np.ascontiguousarray(fx["p0"])
    uout = np.zeros((3, nz, N2), f32)
    harness.xh_correct(C.byref(t), nz, C.c_float(hz), _p(hb), _p(p0), _p(A), _p(uout))
    assert rel_l2(uout, fx["u0"]) < 1e-6
    # second deferred-correction iteration of the predictor (rhs only): reads the previous iterate
    rhs2 = np.zeros((3, nz, N2), f32)
    harness.xh_setup_advection(C.byref(t), nz, C.c_float(hz), _p(u), _p(ustar), _p(bvel), C.c_float(dt), None, None, _p(rhs2), 0)
    assert rel_l2(rhs2, ee.adv_rhs(cd, u, ustar, bvel, dt)) < 5e-7


def test_substep_orchestration_reaches_the_reference_result(harness, tables, golden):
    """The sequence of fgb_extruded3_piso_substep (predictor; per corrector: HbyA, 4 x [divergence with the deferred term of the
    current pressure, CG started from zero in the first iteration and from the previous pressure afterwards, mean removal],
    corrector) replayed on the CPU with the host-executed kernel cell code and numpy Krylov solvers, against the reference's
    velocity / pressure after the substep and its iteration counts."""
    from fluidgym_b200.extruded3d import extruded_neighbours
    cd, t, _keep = tables
    fx = golden("cyl3d_substep0.npz")
    nz, N2 = fx["A"].shape
    dt, hz = float(fx["dt"][0]), float(fx["hz"][0])
    hzc, dtc = C.c_float(hz), C.c_float(dt)
    u, bvel = np.ascontiguousarray(fx["u_in"]), np.ascontiguousarray(fx["bvel"])
    p = np.ascontiguousarray(fx["p_in"]).copy()
    nbr6 = extruded_neighbours(np.asarray(cd.nbr), nz)
    coff, A, rhs = np.zeros((6, nz, N2), f32), np.zeros((nz, N2), f32), np.zeros((3, nz, N2), f32)
    harness.xh_setup_advection(C.byref(t), nz, hzc, _p(u), _p(u), _p(bvel), dtc, _p(coff), _p(A), _p(rhs), 1)
    Cm = _csr(nbr6, coff, A)
    ures = np.zeros((3, nz, N2), f32)
    bicg_its = []
    for c in range(3):
        x, it = _bicgstab(Cm, rhs[c].reshape(-1), 1e-5)
        ures[c] = x.reshape(nz, N2)
        bicg_its.append(it)
    assert rel_l2(ures, fx["ustar"]) < 2e-5
    assert all(abs(a - (b + 1)) <= 1 for a, b in zip(bicg_its, fx["bicg_iters"]))          # the reference reports the last index
    poff, pdiag = np.zeros((6, nz, N2), f32), np.zeros((nz, N2), f32)
    harness.xh_pressure_matrix(C.byref(t), nz, hzc, _p(A), _p(poff), _p(pdiag))
    Pm = _csr(nbr6, poff, pdiag)
    hb, div, cg_its = np.zeros((3, nz, N2), f32), np.zeros((nz, N2), f32), []
    for cs in range(2):
        harness.xh_hbya(C.byref(t), nz, hzc, _p(u), _p(ures), _p(bvel), _p(coff), _p(A), dtc, _p(hb))
        for ps in range(4):
            harness.xh_divergence(C.byref(t), nz, hzc, _p(hb), _p(bvel), _p(p), _p(A), _p(div))
            x, it = _cg(Pm, div.reshape(-1), np.zeros(nz * N2, f32) if ps == 0 else p.reshape(-1), 5e-7)
            p = np.ascontiguousarray(x.reshape(nz, N2))
            cg_its.append(it)
        out = np.zeros((3, nz, N2), f32)
        harness.xh_correct(C.byref(t), nz, hzc, _p(hb), _p(p), _p(A), _p(out))
        ures = out
        if cs == 0:
            assert rel_l2(ures, fx["u0"]) < 1e-4 and rel_l2(p, fx["p0"]) < 2e-3
    print("extruded substep on the CPU: u", rel_l2(ures, fx["u1"]), "p", rel_l2(p, fx["p1"]), "CG iterations", cg_its, "reference", fx["cg_iters"].tolist())
    # observed: u 1.3e-6, p 3.3e-5, CG iterations 1577, 1376, 1788, 1969, 2276, 1967, 1967, 1967 = the reference's (it reports the
    # index of the last iteration, i.e. one less)
    assert rel_l2(ures, fx["u1"]) < 1e-5 and rel_l2(p, fx["p1"]) < 3e-4
    ref = fx["cg_iters"].astype(np.float64) + 1
    assert np.all(np.abs(np.array(cg_its) - ref) <= 0.02 * ref)


def test_cell_code_reduces_to_the_2d_operators_on_the_airfoil_mesh(harness):
    """Size-independent property on the mesh of the next extruded environment (Airfoil3D: 6 blocks, 46 806 cells per plane, other
    block connections / corner rules than the cylinder): on a z-INVARIANT state with w = 0 every in-plane coefficient the extruded
    cell code produces must equal the 2-D operator of oracle/table_eval.py (the specification of the GPU-verified 2-D kernels), and
    the z faces must contribute exactly -nu / hz^2 (matrix), det / (hz A) (pressure) and nothing to right-hand sides, divergence and
    the in-plane velocity correction."""
    import table_eval as te
    from fluidgym_b200.envs.airfoil_domain import make_airfoil_domain
    cd = make_airfoil_domain().prepare()
    t, _keep = host_tables(cd)
    N, NB, nz, hz, dt = cd.N, cd.NB, 3, 0.37, 0.004
    rng = np.random.default_rng(5)
    u2 = (0.3 + 0.1 * rng.standard_normal((2, N))).astype(f32)
    us2 = (u2 + 0.01 * rng.standard_normal((2, N))).astype(f32)
    b2 = np.ascontiguousarray(cd.bvel0[:, :NB] + 0.05 * rng.standard_normal((2, NB))).astype(f32)
    p2 = rng.standard_normal(N).astype(f32)

    def ext(a, comps):                                       # [c, N] -> z-invariant [comps, nz, N] (missing components zero)
        out = np.zeros((comps, nz) + a.shape[1:], f32)
        out[:a.shape[0]] = a[:, None]
        return np.ascontiguousarray(out)

    u3, us3, b3 = ext(u2, 3), ext(us2, 3), ext(b2, 3)
    p3 = np.ascontiguousarray(np.broadcast_to(p2, (nz, N))).astype(f32)
    hzc, dtc, nu = C.c_float(hz), C.c_float(dt), f32(cd.visc)
    dz = nu / f32(hz * hz)
    coff, A3, rhs = np.zeros((6, nz, N), f32), np.zeros((nz, N), f32), np.zeros((3, nz, N), f32)
    harness.xh_setup_advection(C.byref(t), nz, hzc, _p(u3), _p(us3), _p(b3), dtc, _p(coff), _p(A3), _p(rhs), 1)
    off2, A2 = te.assemble_C(cd, u2, b2, dt)
    rhs2 = te.adv_rhs(cd, u2, us2, b2, dt)
    for k in range(nz):
        assert rel_l2(coff[:4, k], off2) < 1e-6
        assert np.abs(coff[4:, k] + dz).max() < 1e-6 * dz
        assert rel_l2(A3[k] - 2 * dz, A2) < 1e-6
        assert rel_l2(rhs[:2, k], rhs2) < 1e-6 and not rhs[2, k].any()
    poff, pdiag = np.zeros((6, nz, N), f32), np.zeros((nz, N), f32)
    harness.xh_pressure_matrix(C.byref(t), nz, hzc, _p(A3), _p(poff), _p(pdiag))
    Po2, Pd2 = te.build_P(cd, A3[0])
    assert rel_l2(poff[:4, 1], Po2 * f32(hz)) < 1e-6
    assert rel_l2(poff[4, 1], cd.det / f32(hz) / A3[0]) < 1e-6 and rel_l2(poff[5, 1], poff[4, 1]) < 1e-7
    assert rel_l2(pdiag[1] + poff[4, 1] + poff[5, 1], Pd2 * f32(hz)) < 1e-5
    hb = np.zeros((3, nz, N), f32)
    harness.xh_hbya(C.byref(t), nz, hzc, _p(u3), _p(us3), _p(b3), _p(coff), _p(A3), dtc, _p(hb))
    hb2 = te.hbya(cd, u2, us2, off2, A3[0], b2, dt)          # in-plane part with the same diagonal ...
    assert rel_l2(hb[:2, 2], hb2 + 2 * dz * us2 / A3[0]) < 2e-6    # ... plus the two z neighbours of a z-invariant iterate
    assert not hb[2].any()
    div = np.zeros((nz, N), f32)
    harness.xh_divergence(C.byref(t), nz, hzc, _p(hb), _p(b3), _p(p3), _p(A3), _p(div))
    d2 = te.divergence(cd, hb[:2, 0], b2, pres=p2, A=A3[0])
    assert rel_l2(div[1], d2 * f32(hz)) < 2e-5
    out = np.zeros((3, nz, N), f32)
    harness.xh_correct(C.byref(t), nz, hzc, _p(hb), _p(p3), _p(A3), _p(out))
    assert rel_l2(out[:2, 1], te.correct(cd, hb[:2, 0], p2, A3[0])) < 1e-6 and not out[2].any()
