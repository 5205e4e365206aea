"""Static plan of the register-blocked on-chip CG (fluidgym_b200/strip_plan.py, cg_impl 11) on the CPU: the padded
strip layout + ghost replicas reproduce the table-driven operator, the export lists deliver every ghost, and the
kernel's iteration (ghost slots updated redundantly, only r exchanged) is the textbook CG of CG.cu:225-446."""
import numpy as np
import pytest

from fluidgym_b200 import strip_plan as sp
from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
from fluidgym_b200.envs.rbc_domain import make_rbc_domain


def _laplacian(cd, seed=0):
    """Symmetric positive semi-definite 5-point matrix on the neighbour table: off[f][g] = -w(edge), diag = sum w."""
    rng = np.random.default_rng(seed)
    nbr = np.asarray(cd.nbr)
    N = cd.N
    off = np.zeros((4, N))
    for f in range(4):
        nb = nbr[f]
        ok = nb >= 0
        g = np.arange(N)
        lo, hi = np.minimum(g, nb), np.maximum(g, nb)
        w = 0.5 + ((lo * 7919 + hi * 104729) % 1000) / 1000.0      # same weight seen from both sides of an edge
        off[f] = np.where(ok, -w, 0.0)
    diag = -off.sum(axis=0)
    return nbr, diag, off


def _table_spmv(nbr, diag, off, v):
    out = diag * v
    for f in range(4):
        nb = nbr[f]
        out = out + off[f] * np.where(nb >= 0, v[np.maximum(nb, 0)], 0.0)
    return out


DOMAINS = {"cyl8": lambda: make_cylinder_domain(8).prepare(), "rbc": lambda: make_rbc_domain()[0].prepare()}


@pytest.fixture(scope="module", params=["cyl8", "rbc"])
def dom(request):
    return request.param, DOMAINS[request.param]()


@pytest.mark.parametrize("cs", [None, 2, 4])
def test_plan_covers_domain_and_reproduces_operator(dom, cs):
    name, cd = dom
    plan = sp.plan_for_domain(cd, cs=cs)
    if plan is None:                      # a forced cluster size may be too small for the default shape
        assert cs is not None and sp.plan_for_domain(cd) is not None
        return
    real = plan.cell[plan.cell >= 0]
    assert np.array_equal(np.sort(real), np.arange(cd.N))                      # every cell exactly once
    assert int(plan.cnt[:, 0].sum()) == int(plan.cnt[:, 1].sum())                # every remote ghost has exactly one exporter
    assert int(plan.cnt[:, 0].sum() + plan.cnt[:, 2].sum()) == int((plan.cell <= -2).sum())
    assert ((plan.cpt - 1) * plan.thread[:, :, 1] % 32 == 0).all()            # conflict-free band transitions (CPT S = S mod 32)
    nbr, diag, off = _laplacian(cd)
    em = sp.StripEmulator(plan, nbr, diag, off)
    rng = np.random.default_rng(1)
    v = rng.standard_normal(cd.N)
    regs = em.scatter(v)
    # the two delivery paths of the kernel: remote ghosts receive along the remote export lists, local ghosts are mirrored by
    # their owners when the vector is published
    owners_only = np.where(em.real, regs, 0)
    rs = em.push_remote(owners_only)
    for r in range(plan.cs):
        assert np.array_equal(rs[r][em.gidx[r][em.rghost[r]]], regs[r][em.rghost[r]])
    vss = em.publish(np.where(em.lghost, 0, regs))
    assert np.array_equal(em.reload_local_ghosts(np.where(em.lghost, 0, regs), vss), regs)
    got = em.gather(em.spmv(regs, vss))
    want = _table_spmv(nbr, diag, off, v)
    assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()
    # slots stay inside the shared-memory arrays, one row of slack on both sides
    s = em.slot[plan.thread[:, :, 1] > 0]
    st = plan.thread[:, :, 1][plan.thread[:, :, 1] > 0][:, None]
    assert (s - st).min() >= 0 and (s + st).max() < plan.slots


def test_kernel_iteration_is_textbook_cg(dom):
    name, cd = dom
    plan = sp.plan_for_domain(cd, cs=4 if name == "cyl8" else None)
    nbr, diag, off = _laplacian(cd)
    em = sp.StripEmulator(plan, nbr, diag, off)
    rng = np.random.default_rng(2)
    f = rng.standard_normal(cd.N)
    f -= f.mean()
    x_regs, used = em.cg(f, 1e-9, 3000, reset_steps=100)
    # ghost replicas never drift from their owners (they receive only r)
    xg = em.gather(x_regs)
    assert np.array_equal(em.scatter(xg)[em.ghost], x_regs[em.ghost])
    # plain CG on the table operator, same reset rule
    x = np.zeros(cd.N); r = f.copy(); p = r.copy(); rho = r @ r
    for i in range(3000):
        if (i + 1) % 100 == 0:
            r = f - _table_spmv(nbr, diag, off, x); p = r.copy(); rho = r @ r
        ap = _table_spmv(nbr, diag, off, p)
        a = rho / (p @ ap)
        x += a * p; r -= a * ap
        rr = r @ r
        if np.sqrt(rr / cd.N) < 1e-9:
            break
        p = r + (rr / rho) * p; rho = rr
    assert abs(i - used) <= 2
    assert np.abs(xg - x).max() < 1e-6 * np.abs(x).max()
    assert np.abs(_table_spmv(nbr, diag, off, xg) - f).max() < 1e-6
