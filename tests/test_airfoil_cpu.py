"""Airfoil2D (config 4): grid generator, domain tables and the CPU oracle on the six-block C-mesh, against
fixtures recorded from the unmodified reference on a B200 (tests/golden/airfoil_*.npz)."""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import Oracle


def test_vertex_grids_are_bit_identical(golden):
    from fluidgym_b200.envs.airfoil_domain import airfoil_vertex_grids
    ref = golden("airfoil_vertices.npz")
    grids = airfoil_vertex_grids()
    assert [g.shape for g in grids] == [(2, 22, 72), (2, 22, 96), (2, 96, 70), (2, 96, 70), (2, 96, 160), (2, 96, 160)]
    for i, g in enumerate(grids):
        assert np.array_equal(g, ref[f"b{i}"]), i


def test_domain_layout_and_transforms(golden):
    from fluidgym_b200.domain import FIXED, cell_transforms
    from fluidgym_b200.envs.airfoil_domain import BOT, make_airfoil_domain
    spec = make_airfoil_domain()
    assert sum(b.nx * b.ny for b in spec.blocks) == 46806
    # closing bot:+y (airfoil) also closed bot:-y, the tunnel wall, which the reference never closes explicitly
    assert spec.blocks[BOT].bounds[2].type == FIXED and spec.blocks[BOT].bounds[3].type == FIXED
    g = golden("airfoil_geometry.npz")
    T = np.concatenate([cell_transforms(b.vertex).reshape(-1, 9) for b in spec.blocks])
    scale = np.abs(g["T"]).max(axis=1, keepdims=True)
    assert (np.abs(T - g["T"]) / scale).max() < 1e-6


@pytest.fixture(scope="module")
def setup(airfoil, golden):
    spec, cd = airfoil
    fx = golden("airfoil_substep0.npz")
    orc = Oracle(cd.sizes, cd.btype, cd.bconn, cd.T, cd.bT, fx["bvel_in"], float(cd.visc), adv_nonortho_steps=2,
                 p_nonortho_steps=4, adv_tol=1e-6, p_tol=1e-7)
    return cd, fx, orc


def test_oracle_predictor_with_two_nonorthogonal_iterations(setup):
    """C, A, RHS of both deferred-correction iterations and the BiCGStab solves (SIM.py:1662-1757)."""
    cd, fx, orc = setup
    dt = float(fx["dt"][0])
    u = fx["u_in"]
    val, idx, A = orc.build_C(u, dt)
    assert rel_l2(A, fx["A"]) < 2e-6
    rhs0 = orc.adv_rhs(u, u, dt)
    assert rel_l2(rhs0, fx["rhs0"]) < 2e-6
    x = np.zeros_like(u)
    for c in range(2):
        x[c], it, res, conv = orc.bicgstab(val, idx, rhs0[c], tol=1e-6)
        assert conv and abs(it - int(fx["bicg_iters"][0][c])) <= 1
    rhs1 = orc.adv_rhs(u, x, dt)
    assert rel_l2(rhs1, fx["rhs1"]) < 5e-6
    for c in range(2):
        x[c], it, res, conv = orc.bicgstab(val, idx, rhs1[c], x0=x[c], tol=1e-6)
        assert conv and abs(it - int(fx["bicg_iters"][1][c])) <= 1
    assert rel_l2(x, fx["ustar"]) < 5e-6


def test_table_kernels_specification_matches_oracle_on_the_c_mesh(setup):
    """The neighbour-table formulation the CUDA kernels implement (oracle/table_eval.py) against the literal
    oracle on a domain with rotated block connections and a wake cut."""
    import table_eval as te
    cd, fx, orc = setup
    dt = float(fx["dt"][0])
    u = fx["u_in"]
    Coff, A = te.assemble_C(cd, u, fx["bvel_in"], dt)
    rhs = te.adv_rhs(cd, u, u, fx["bvel_in"], dt)
    assert rel_l2(A, fx["A"]) < 2e-6 and rel_l2(rhs, fx["rhs0"]) < 2e-6
