"""CPU emulation of the control flow of k3_cg_fused (csrc/ortho3_b200.cuh, opt-in FGB_K3_CG_FUSED=1): the search-direction update
p <- r + beta p is folded into the next matrix-vector product (every row forms r[n] + beta p_old[n] for itself and its neighbours and
stores its own value into a second buffer; the buffers swap after the step) with the residual reset, best-iterate tracking,
100-rising-steps cut-off and iteration cap of k3_cg.  Both recurrences are written here in float32 numpy with the same expressions:
iterates, iteration counts and results must be bit-identical.  (The GPU check of the kernels themselves:
tests/zz_first_run_worker.py cg_fused.)"""
import numpy as np
import pytest
import scipy.sparse as sp

f32 = np.float32


def _poisson(n, seed):
    rng = np.random.default_rng(seed)
    k = (0.5 + rng.random(n)).astype(f32)
    main = np.zeros(n, f32)
    off = np.zeros(n - 1, f32)
    off[:] = -(k[:-1] + k[1:]) * f32(0.5)
    main[:-1] -= off
    main[1:] -= off
    jump = -(rng.random(n - 7).astype(f32))                                # a second, wider coupling: 5 entries per row
    main[:-7] -= jump
    main[7:] -= jump
    A = sp.diags([main, off, off, jump, jump], [0, 1, -1, 7, -7], format="csr", dtype=f32)
    b = rng.standard_normal(n).astype(f32)
    return A, (b - b.mean()).astype(f32)                                   # singular, consistent system like the pressure matrix


def _crit(rr, n):
    return f32(np.sqrt(rr)) * f32(1.0 / np.sqrt(f32(n)))


def cg_reference(A, b, tol, reset_steps, maxit):
    """statement order of k3_cg"""
    n = b.size
    x = np.zeros(n, f32)
    r = b.copy(); p = r.copy()
    rho = np.dot(r, r)
    bestc = lastc = f32(0); best_it, rising, used = -1, 0, -1
    best = x.copy()
    until_reset = reset_steps - 1 if reset_steps > 0 else -1
    for i in range(maxit):
        do_reset = until_reset == 0
        if until_reset >= 0:
            until_reset = reset_steps - 1 if do_reset else until_reset - 1
        if do_reset:
            r = (b - A @ x).astype(f32); p = r.copy(); rho = np.dot(r, r)
        ap = (A @ p).astype(f32)
        alpha = f32(rho / np.dot(p, ap))
        x = (x + alpha * p).astype(f32)
        r = (r - alpha * ap).astype(f32)
        rr = np.dot(r, r)
        crit = _crit(rr, n)
        if i == 0 or crit < bestc:
            bestc, best_it, best = crit, i, x.copy()
        rising = rising + 1 if (i > 0 and crit >= lastc) else 0
        lastc, used = crit, i
        if crit < tol:
            break
        if i == maxit - 1 or rising >= 100:
            x, used = best.copy(), best_it
            break
        beta = f32(rr / rho); rho = rr
        p = (r + beta * p).astype(f32)
    return x, used


def cg_fused(A, b, tol, reset_steps, maxit):
    """statement order of k3_cg_fused: two direction buffers, `fresh` = the buffer holds the direction itself"""
    n = b.size
    x = np.zeros(n, f32)
    r = b.copy(); p = r.copy(); p2 = np.full(n, np.nan, f32)              # the spare buffer starts with garbage
    rho = np.dot(r, r)
    beta = f32(0)
    fresh = True
    bestc = lastc = f32(0); best_it, rising, used = -1, 0, -1
    best = x.copy()
    until_reset = reset_steps - 1 if reset_steps > 0 else -1
    for i in range(maxit):
        do_reset = until_reset == 0
        if until_reset >= 0:
            until_reset = reset_steps - 1 if do_reset else until_reset - 1
        if do_reset:
            r = (b - A @ x).astype(f32); p[:] = r; rho = np.dot(r, r)
            fresh = True
        if fresh:
            pc = p
            ap = (A @ p).astype(f32)
        else:
            pn = (r + beta * p).astype(f32)                                # what every row forms for itself and its neighbours
            ap = (A @ pn).astype(f32)
            p2[:] = pn
            pc = p2
        alpha = f32(rho / np.dot(pc, ap))
        x = (x + alpha * pc).astype(f32)
        r = (r - alpha * ap).astype(f32)
        rr = np.dot(r, r)
        if not fresh:
            p, p2 = p2, p
        fresh = False
        crit = _crit(rr, n)
        if i == 0 or crit < bestc:
            bestc, best_it, best = crit, i, x.copy()
        rising = rising + 1 if (i > 0 and crit >= lastc) else 0
        lastc, used = crit, i
        if crit < tol:
            break
        if i == maxit - 1 or rising >= 100:
            x, used = best.copy(), best_it
            break
        beta = f32(rr / rho); rho = rr
    return x, used


@pytest.mark.parametrize("n, tol, reset_steps, maxit", [(400, 1e-6, 100, 5000), (400, 1e-6, 0, 5000), (900, 1e-9, 100, 350), (257, 1e-5, 7, 5000)])
def test_fused_direction_update_is_bit_identical(n, tol, reset_steps, maxit):
    A, b = _poisson(n, seed=n)
    x0, it0 = cg_reference(A, b, f32(tol), reset_steps, maxit)
    x1, it1 = cg_fused(A, b, f32(tol), reset_steps, maxit)
    assert it0 == it1 and it0 > 20
    assert np.array_equal(x0, x1)


def test_fused_update_on_the_extruded_pressure_system_of_the_reference_trace(golden):
    """The same comparison on the real thing: pressure matrix (kernel cell code on the host) and right-hand side of the first
    pressure solve the unmodified reference traced on CylinderJet3D -- 1 576 iterations with a residual reset every 100, the
    reference's own count."""
    import ctypes as C
    import os
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_harness"))
    from extruded_standin import _p, build_harness, csr, host_tables
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
    from fluidgym_b200.extruded3d import extruded_neighbours
    cd = make_cylinder_domain(8).prepare()
    t, _keep = host_tables(cd)
    fx = golden("cyl3d_substep0.npz")
    nz, N2 = fx["A"].shape
    A = np.ascontiguousarray(fx["A"])
    poff, pd = np.zeros((6, nz, N2), f32), np.zeros((nz, N2), f32)
    build_harness().xh_pressure_matrix(C.byref(t), nz, C.c_float(float(fx["hz"][0])), _p(A), _p(poff), _p(pd))
    P = csr(extruded_neighbours(np.asarray(cd.nbr), nz), poff, pd)
    b = np.ascontiguousarray(fx["div0"]).reshape(-1)
    x0, i0 = cg_reference(P, b, f32(5e-7), 100, 5000)
    x1, i1 = cg_fused(P, b, f32(5e-7), 100, 5000)
    assert i0 == i1 == int(fx["cg_iters"][0]) and np.array_equal(x0, x1)
