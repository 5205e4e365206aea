"""The table compiler (fluidgym_b200.domain) against the literal oracle: every table-driven op,
evaluated in numpy exactly as the CUDA kernels evaluate it (oracle/table_eval.py), must reproduce the
oracle's per-cell walk.  Bar: fp32 round-off (1e-6 relative)."""
import numpy as np
import pytest

import table_eval as te
from conftest import rel_l2
from oracle import Oracle


@pytest.fixture(scope="module")
def ctx(cyl24_own, golden):
    spec, cd = cyl24_own
    fx = golden("cyl24_substep1.npz")
    orc = Oracle.from_compiled(cd)
    orc.bvel[:] = fx["bvel_in"]
    return cd, fx, orc


def test_tables_shapes(ctx):
    cd, fx, orc = ctx
    assert cd.N == 14232 and cd.NB == orc.NB == fx["bvel_in"].shape[1]
    assert cd.nbr.shape == (4, cd.N) and cd.Wp.shape == (5, 5, cd.N)
    # every boundary face is referenced by exactly one cell face
    refs = np.sort(-1 - cd.nbr[cd.nbr < 0])
    assert np.array_equal(refs, np.arange(cd.NB))
    # neighbour relation is symmetric
    for f in range(4):
        nb = cd.nbr[f]
        inner = np.nonzero(nb >= 0)[0]
        back = np.stack([cd.nbr[k][nb[inner]] for k in range(4)])
        assert (back == inner[None, :]).any(axis=0).all()


def test_assembly_ops_match_oracle(ctx):
    cd, fx, orc = ctx
    dt = float(fx["dt"][0])
    u, bvel = fx["u_in"], fx["bvel_in"]
    val, idx, A = orc.build_C(u, dt)
    off, A2 = te.assemble_C(cd, u, bvel, dt)
    assert rel_l2(A2, A) < 1e-6
    for f in range(4):
        assert np.array_equal(np.where(cd.nbr[f] >= 0, cd.nbr[f], -1), idx[:, f + 1])
        assert rel_l2(off[f], val[:, f + 1]) < 1e-6
    assert rel_l2(te.adv_rhs(cd, u, u, bvel, dt), orc.adv_rhs(u, u, dt)) < 1e-6
    Pv, Pi = orc.build_P(A)
    Po, Pd = te.build_P(cd, A)
    assert rel_l2(Pd, Pv[:, 0]) < 1e-6
    for f in range(4):
        assert rel_l2(Po[f], Pv[:, f + 1]) < 1e-6
    ures = fx["ustar"]
    h = orc.pressure_rhs(u, ures, val, idx, A, dt)
    assert rel_l2(te.hbya(cd, u, ures, off, A2, bvel, dt), h) < 1e-6
    d = orc.div(h, fx["presres_in"], A)
    d2 = te.divergence(cd, h, bvel, fx["presres_in"], A)
    assert np.abs(d2 - d).max() < 1e-6 * np.abs(d).max() + 1e-7
    assert rel_l2(te.correct(cd, h, fx["p0"], A), orc.correct(h, fx["p0"], A)) < 1e-6
    assert abs(te.max_velocity(cd, u, bvel) - orc.max_velocity(u)) < 1e-4


def test_single_block_periodic_box():
    """Rayleigh-Benard style domain: one block, periodic in x, walls in y, orthogonal grid -> all
    non-orthogonal tables vanish and the stencil is the classic 5-point Laplacian."""
    from fluidgym_b200.domain import DomainSpec
    from fluidgym_b200.grids import uniform_box_grid
    spec = DomainSpec(0.01)
    b = spec.create_block(uniform_box_grid(12, 7, (0.0, 0.0), (3.0, 1.0)))
    spec.make_periodic(b, 0)
    spec.close_boundary(b, "-y")
    spec.close_boundary(b, "+y")
    cd = spec.prepare()
    assert cd.N == 84 and cd.NB == 24
    assert np.abs(cd.no_gP).max() == 0 and np.abs(cd.no_wv).max() == 0
    assert (cd.nbr[0][::12] == np.arange(84)[11::12]).all()      # -x neighbour of column 0 wraps to column 11
    rng = np.random.default_rng(0)
    u = rng.standard_normal((2, cd.N)).astype(np.float32)
    orc = Oracle.from_compiled(cd)
    val, idx, A = orc.build_C(u, 0.05)
    off, A2 = te.assemble_C(cd, u, cd.bvel0, 0.05)
    assert rel_l2(A2, A) < 1e-6
    for f in range(4):
        assert rel_l2(off[f], val[:, f + 1]) < 1e-6
    Pv, Pi = orc.build_P(A)
    Po, Pd = te.build_P(cd, A)
    assert rel_l2(Pd, Pv[:, 0]) < 1e-6
    assert np.abs(Pd + Po.sum(0)).max() < 1e-4 * np.abs(Pd).max()   # pure Neumann: zero row sums
