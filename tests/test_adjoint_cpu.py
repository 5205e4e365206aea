"""The hand-derived reverse-mode adjoint of one PISO substep (oracle/adjoint_eval.py, the float64 numpy
specification of the CUDA adjoint kernels) against central finite differences: directional derivatives of a
random linear functional of (u_out, p_out) w.r.t. u, p_prev and the boundary velocities."""
import numpy as np
import pytest

import adjoint_eval as ae


@pytest.fixture(scope="module")
def small():
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
    cd = make_cylinder_domain(8).prepare()
    return cd, ae.T64(cd)


@pytest.mark.parametrize("n_adv,n_p", [(1, 1), (2, 3)])
def test_substep_vjp_matches_finite_differences(small, n_adv, n_p):
    cd, t = small
    rng = np.random.default_rng(1)
    u = 0.3 * rng.standard_normal((2, t.N)); u[0] += 1.0
    p0 = 0.1 * rng.standard_normal(t.N)
    bvel = cd.bvel0[:, :t.NB].astype(np.float64) + 0.05 * rng.standard_normal((2, t.NB))
    dt = 0.01
    wu, wp = rng.standard_normal((2, t.N)), rng.standard_normal(t.N)

    def J(u_, p_, b_):
        uo, po, _ = ae.substep(t, u_, p_, b_, dt, n_adv=n_adv, n_p=n_p)
        return float((wu * uo).sum() + (wp * po).sum())

    uo, po, tape = ae.substep(t, u, p0, bvel, dt, n_adv=n_adv, n_p=n_p)
    ub, pb, bb = ae.substep_vjp(t, u, p0, bvel, dt, tape, wu, wp)
    for name, grad, shape in (("u", ub, u.shape), ("p_prev", pb, p0.shape), ("bvel", bb, bvel.shape)):
        for trial in range(2):
            d = rng.standard_normal(shape)
            eps = 1e-6
            if name == "u":
                fd = (J(u + eps * d, p0, bvel) - J(u - eps * d, p0, bvel)) / (2 * eps)
            elif name == "p_prev":
                fd = (J(u, p0 + eps * d, bvel) - J(u, p0 - eps * d, bvel)) / (2 * eps)
            else:
                fd = (J(u, p0, bvel + eps * d) - J(u, p0, bvel - eps * d)) / (2 * eps)
            an = float((grad * d).sum())
            # the absolute term covers the round-off of the finite difference itself: the pressure operator is
            # nearly singular, so J carries ~1e-12 noise that eps = 1e-6 amplifies to ~1e-6
            assert abs(fd - an) <= 2e-5 * max(abs(fd), abs(an)) + 5e-6, (name, trial, fd, an)


def test_scalar_substep_vjp_matches_finite_differences():
    """RBC substep (passive scalar with the old velocity, buoyancy from the new temperature, orthogonal PISO):
    gradients w.r.t. u, the boundary velocities, the temperature and the heater (boundary) temperatures."""
    from fluidgym_b200.envs.rbc_domain import make_rbc_domain
    cd = make_rbc_domain(n_heaters=3, heater_width=4)[0].prepare()
    t = ae.T64(cd)
    rng = np.random.default_rng(2)
    u = 0.05 * rng.standard_normal((2, t.N))
    p0 = np.zeros(t.N)
    bvel = 0.01 * rng.standard_normal((2, t.NB))
    T = np.clip(0.5 + 0.3 * rng.standard_normal(t.N), 0, 1)
    sb = cd.sb_val0[:t.NB].astype(np.float64) + 0.2 * rng.standard_normal(t.NB)
    dt, beta = 0.05, 1.0
    wu, wp, wT = rng.standard_normal((2, t.N)), rng.standard_normal(t.N), rng.standard_normal(t.N)

    def J(u_, b_, T_, s_):
        uo, po, Tn, _ = ae.substep_scalar(t, u_, p0, b_, T_, s_, dt, beta)
        return float((wu * uo).sum() + (wp * po).sum() + (wT * Tn).sum())

    uo, po, Tn, tape = ae.substep_scalar(t, u, p0, bvel, T, sb, dt, beta)
    ub, pb, bb, Tb, sbb = ae.substep_scalar_vjp(t, u, p0, bvel, T, sb, dt, beta, tape, wu, wp, wT)
    args = [u, bvel, T, sb]
    for k, (name, grad) in enumerate((("u", ub), ("bvel", bb), ("T", Tb), ("sbval", sbb))):
        for trial in range(2):
            d = rng.standard_normal(args[k].shape)
            eps = 1e-6 if name in ("u", "bvel") else 1e-3      # the substep is linear in T and in the boundary temperatures
            hi = [a + eps * d if i == k else a for i, a in enumerate(args)]
            lo = [a - eps * d if i == k else a for i, a in enumerate(args)]
            fd = (J(*hi) - J(*lo)) / (2 * eps)
            an = float((grad * d).sum())
            assert abs(fd - an) <= 2e-5 * max(abs(fd), abs(an)) + 5e-6, (name, trial, fd, an)


def test_substep_vjp_is_exact_at_the_junction_cells(small):
    """Single-cell derivatives at the cells where a block connection ends on a prescribed boundary (domain corner left|top at
    inflow + wall; right|bottom|wake at the lower wall): these are the cells that carry 88 % of the remaining difference to the
    REFERENCE's d reward / d u0 (profiles/r02_cyl24_gradient_difference_map.md).  The hand-written adjoint equals central
    differences of the forward substep there, so that difference is not an error of the adjoint."""
    cd, t = small
    offs, sizes = np.array(cd.offsets), cd.sizes

    def cell(bi, x, y):
        return int(offs[bi] + (y % sizes[bi][1]) * sizes[bi][0] + (x % sizes[bi][0]))

    junction = [cell(0, 0, -1), cell(1, 0, -1), cell(2, -1, 0), cell(3, -1, 0), cell(4, 0, 0), cell(0, 1, -2)]
    rng = np.random.default_rng(3)
    u = 0.3 * rng.standard_normal((2, t.N)); u[0] += 1.0
    p0 = 0.1 * rng.standard_normal(t.N)
    bvel = cd.bvel0[:, :t.NB].astype(np.float64) + 0.05 * rng.standard_normal((2, t.NB))
    dt = 0.01
    wu, wp = rng.standard_normal((2, t.N)), rng.standard_normal(t.N)

    def J(u_):
        uo, po, _ = ae.substep(t, u_, p0, bvel, dt, n_adv=2, n_p=3)
        return float((wu * uo).sum() + (wp * po).sum())

    uo, po, tape = ae.substep(t, u, p0, bvel, dt, n_adv=2, n_p=3)
    ub, _, _ = ae.substep_vjp(t, u, p0, bvel, dt, tape, wu, wp)
    scale = np.abs(ub).max()
    for c in junction:
        for comp in (0, 1):
            d = np.zeros_like(u); d[comp, c] = 1.0
            fd = (J(u + 1e-6 * d) - J(u - 1e-6 * d)) / 2e-6
            assert abs(ub[comp, c] - fd) < 2e-5 * scale, (c, comp, ub[comp, c], fd)
