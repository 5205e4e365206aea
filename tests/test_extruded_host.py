"""The per-cell code of the D = 3 extruded-domain kernels (fluidgym_b200/csrc/extruded3_b200.cuh, __host__ __device__) executed
ON THE CPU through tests/cpu_harness/extruded_host.cu and compared with the op trace of the unmodified reference on
CylinderJet3D-easy (tests/golden/cyl3d_substep*.npz) and with the numpy specification tests/extruded_eval.py.  This verifies
the arithmetic of the kernels without a GPU; their launch glue has not run on a GPU yet."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import extruded_eval as ee
from conftest import ROOT, rel_l2
from fluidgym_b200 import native

f32 = np.float32
FIELDS = ["nbr", "fl_comp", "minv", "det", "Cd", "Wp", "no_idx", "no_face", "no_gP", "no_gN", "no_wv", "nob_idx", "nob_w", "b_minv", "b_det",
          "b_alpha", "b_cell", "b_face"]


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(ROOT, "tests", "cpu_harness", "extruded_host.cu")
    out = os.path.join(ROOT, "tests", "_build", "libextruded_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    hdr = os.path.join(ROOT, "fluidgym_b200", "csrc", "extruded3_b200.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.check_call([nvcc, "-x", "cu", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC", "-shared",
                               "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "fluidgym_b200", "csrc"), "-o", out, src])
    return C.CDLL(out)


@pytest.fixture(scope="module")
def tables():
    """fgb_tables with HOST pointers into the numpy arrays of the compiled 2-D cylinder domain (resolution 8)"""
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
    cd = make_cylinder_domain(8).prepare()
    keep = {}
    t = native.Tables()
    t.N, t.NB, t.K_no, t.K_nob, t.viscosity = cd.N, cd.NB, cd.K_no, cd.K_nob, float(cd.visc)
    for name in FIELDS:
        a = np.ascontiguousarray(getattr(cd, name))
        if name == "b_face":
            a = np.ascontiguousarray(a.astype(np.int8))
        keep[name] = a
        setattr(t, name, a.ctypes.data)
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
