"""D = 3 (turbulent channel flow): grid generator, box-domain tables and the float32 numpy specification of the 3-D
orthogonal PISO operators (oracle/box3d_eval.py) against a trace of the unmodified reference on a 32 x 32 x 32 channel
(tests/golden/tcf32_*.npz: TCFSmall3D-both-easy-v0 with resolution 32/33, Reichardt profile + 5 % noise)."""
import json
import os

import numpy as np
import pytest

import box3d_eval as be
from conftest import GOLDEN, rel_l2


@pytest.fixture(scope="module")
def tcf(golden):
    g, fx = golden("tcf32_geometry.npz"), golden("tcf32_substep0.npz")
    meta = json.load(open(os.path.join(GOLDEN, "tcf32_meta.json")))

    def full(t7):                      # [.., 7] = (M00, M11, M22, Mi00, Mi11, Mi22, det) -> [.., 19]
        T = np.zeros(t7.shape[:-1] + (19,), np.float32)
        for k, c in enumerate((0, 4, 8, 9, 13, 17, 18)):
            T[..., c] = t7[..., k]
        return T
    T = full(g["Tdiag"])
    bT = {2: full(g["bT2"]).reshape(32, 1, 32, 19), 3: full(g["bT3"]).reshape(32, 1, 32, 19)}
    box = be.Box3D(g["vertex"], closed=(False, True, False), viscosity=meta["viscosity"], T=T, bT=bT)
    bvel = {2: fx["bvel2"].reshape(3, 32, 1, 32), 3: fx["bvel3"].reshape(3, 32, 1, 32)}
    return box, fx, bvel, T, bT, meta


def test_channel_grid_is_bit_identical(golden):
    from fluidgym_b200.grids import channel_vertex_grid
    v = channel_vertex_grid(2.0, np.pi, np.pi / 2, 32, 16, 2, 32)
    assert np.array_equal(v, golden("tcf32_geometry.npz")["vertex"])


def test_box_domain_metrics_and_tables(golden):
    from fluidgym_b200.box3d import Box3DDomain
    g = golden("tcf32_geometry.npz")
    dom = Box3DDomain(g["vertex"], closed=(False, True, False), viscosity=1e-3)
    T = g["Tdiag"]
    for d in range(3):
        assert np.array_equal(dom.h[d], T[..., d])
        assert (np.abs(dom.minv[d] - T[..., 3 + d]) / T[..., 3 + d]).max() < 1e-6      # fast-math reciprocal in the reference
    assert (np.abs(dom.det - T[..., 6]) / T[..., 6]).max() < 1e-6
    # boundary faces: one-sided differences, different summation order -> cancellation error of the first cell height
    assert (np.abs(dom.b_det[:1024] - g["bT2"][:, 6]) / g["bT2"][:, 6]).max() < 1e-5
    assert dom.N == 32768 and dom.NB == 2048
    cells = np.arange(dom.N).reshape(dom.shape)
    assert np.array_equal(dom.nbr[1].reshape(dom.shape), np.roll(cells, -1, axis=2))          # +x periodic
    assert np.array_equal(dom.nbr[2].reshape(dom.shape)[:, 0, :], -1 - np.arange(1024).reshape(32, 32))   # -y wall faces


def test_specification_matches_reference_operators(tcf):
    box, fx, bvel, *_ = tcf
    dt, u = float(fx["dt"][0]), fx["u_in"]
    off, A, fl = be.assemble(box, u, bvel, dt)
    assert rel_l2(A, fx["A"]) < 5e-7
    rhs, Sb = be.adv_rhs(box, u, bvel, dt, fx["src"])
    assert rel_l2(rhs, fx["rhs"]) < 5e-7
    ustar = fx["ustar"].reshape(3, *box.shape)
    assert np.linalg.norm(be.spmv(box, off, A, ustar) - rhs) / np.sqrt(3 * box.N) < 5e-5   # reference solution solves our system
    hb = be.hbya(box, u, ustar, off, A, Sb, dt, fx["src"])
    assert rel_l2(hb, fx["hbya0"]) < 5e-7
    div = be.divergence(box, hb, bvel)
    assert rel_l2(div, fx["div0"]) < 5e-6                                                   # cancellation in the flux differences
    Po, Pd = be.build_P(box, A)
    r = be.spmv(box, Po, Pd, fx["p0"].reshape(1, *box.shape))[0] - fx["div0"].reshape(box.shape)
    assert np.linalg.norm(r) / np.sqrt(box.N) < 2e-6                                        # reference pressure solves our system to its tol (1e-6)
    u0 = be.correct(box, hb, fx["p0"].reshape(1, *box.shape), A)
    assert rel_l2(u0, fx["u0"]) < 5e-7


def test_agent_window_means_matches_loop_definition():
    """extract_moving_window_2d_x_z (obs_extraction.py:255-343) restated without unfold: agent (ix, iz) sees the patch means
    of agents (ix - pad_x + k, iz - pad_z + j), circular, ordered x-major."""
    import torch
    from fluidgym_b200.envs.tcf import agent_window_means
    nax, naz, aw, wx, wz, px, pz = 4, 3, 2, 3, 2, 1, 1
    f = torch.arange(naz * aw * nax * aw, dtype=torch.float32).reshape(naz * aw, nax * aw) ** 1.5
    got = agent_window_means(f[None], nax, naz, aw, wx, wz, px, pz)[0]
    means = f.reshape(naz, aw, nax, aw).mean(dim=(1, 3))
    for ix in range(nax):
        for iz in range(naz):
            for j in range(wz):
                for k in range(wx):
                    assert got[ix * naz + iz, j, k] == means[(iz - pz + j) % naz, (ix - px + k) % nax]


@pytest.mark.parametrize("world", [1, 2, 4])
def test_slab_tables_reproduce_the_global_stencil(golden, world):
    """Slab decomposition (host logic): with the halo planes filled from the z-neighbours, gathering through the local
    neighbour table of every rank returns exactly what the global table returns; metrics incl. halos are copies."""
    from fluidgym_b200.box3d import Box3DDomain, SlabTables
    g = golden("tcf32_geometry.npz")
    dom = Box3DDomain(g["vertex"], viscosity=1e-3)
    x = np.random.default_rng(0).random(dom.N).astype(np.float32)
    nz, P = dom.nz, dom.nx * dom.ny
    xg = x.reshape(nz, P)
    for r in range(world):
        tb = SlabTables(dom, r, world)
        assert tb.NS == tb.N + 2 * P and tb.N * world == dom.N
        loc = tb.take_cells(x)
        loc[tb.N:tb.N + P] = xg[(tb.z0 - 1) % nz]
        loc[tb.N + P:] = xg[(tb.z0 + tb.nzl) % nz]
        for f in range(6):
            n = tb.nbr[f, :tb.N]
            gn = dom.nbr[f].reshape(nz, P)[tb.z0:tb.z0 + tb.nzl].reshape(-1)
            inner = n >= 0
            assert np.array_equal(inner, gn >= 0)
            assert np.array_equal(loc[n[inner]], x[gn[inner]])
        assert np.array_equal(tb.det[:tb.N], dom.det.reshape(nz, P)[tb.z0:tb.z0 + tb.nzl].reshape(-1))
        assert np.array_equal(tb.det[tb.N:tb.N + P], dom.det.reshape(nz, P)[(tb.z0 - 1) % nz])
        faces = tb.take_faces(np.arange(3 * nz * dom.nx, dtype=np.float32).reshape(3, -1), 2)
        assert faces.shape == (3, tb.nzl * dom.nx)
    with pytest.raises(ValueError):
        SlabTables(dom, 0, 5)


@pytest.mark.parametrize("closed", [(False, True, False), (True, True, False), (False, False, False), (True, False, True)])
def test_structured_neighbour_arithmetic_equals_the_tables(closed):
    """The Krylov kernels compute neighbour indices of structured boxes instead of loading them (o3_nbrs): the host mirror of that
    formula reproduces the neighbour tables entry for entry, prescribed-face indices included."""
    from fluidgym_b200.box3d import Box3DDomain, SlabTables, structured_neighbours
    nz, ny, nx = 6, 5, 7
    v = np.stack(np.meshgrid(np.linspace(0, 1, nz + 1), np.linspace(0, 1, ny + 1), np.linspace(0, 2, nx + 1), indexing="ij")[::-1]).astype(np.float32)
    dom = Box3DDomain(v, closed=closed, viscosity=1e-3)
    assert (dom.nz, dom.ny, dom.nx) == (nz, ny, nx)
    n = structured_neighbours(nx, ny, nz, closed, boff=dom.boff)
    tab = np.asarray(dom.nbr).reshape(6, -1)
    assert np.array_equal(n, tab)
    if not closed[2]:
        for world in (1, 2, 3):
            for r in range(world):
                tb = SlabTables(dom, r, world)
                ns = structured_neighbours(nx, ny, tb.nzl, (closed[0], closed[1], False), halo=True, boff=tb.boff)
                assert np.array_equal(ns, tb.nbr[:, :tb.N])


# ---- passive scalar + buoyancy on the 3-D box (RBC3D) -----------------------------------------------------------------
def _full19(t7):
    T = np.zeros(t7.shape[:-1] + (19,), np.float32)
    for k, c in enumerate((0, 4, 8, 9, 13, 17, 18)):
        T[..., c] = t7[..., k]
    return T


@pytest.fixture(scope="module")
def rbc3d(golden):
    g = golden("rbc3d_geometry.npz")
    meta = json.load(open(os.path.join(GOLDEN, "rbc3d_meta.json")))
    nz, ny, nx = g["Tdiag"].shape[:3]
    bT = {2: _full19(g["bT2"]).reshape(nz, 1, nx, 19), 3: _full19(g["bT3"]).reshape(nz, 1, nx, 19)}
    box = be.Box3D(g["vertex"], closed=(False, True, False), viscosity=meta["viscosity"], T=_full19(g["Tdiag"]), bT=bT)
    return box, meta, (nz, ny, nx)


@pytest.mark.parametrize("s", [0, 1])
def test_scalar_specification_matches_reference_trace(rbc3d, golden, s):
    """oracle/box3d_eval.py::assemble_scalar / buoyancy_source and the operators fed by the buoyancy source field against
    the op trace of the unmodified reference on RBC3D (16 x 10 x 16 cells, tests/golden/rbc3d_substep*.npz)."""
    box, meta, (nz, ny, nx) = rbc3d
    fx = golden(f"rbc3d_substep{s}.npz")
    dt, u = float(fx["dt"][0]), fx["u_in"]
    bvel = {2: np.zeros((3, nz, 1, nx), np.float32), 3: np.zeros((3, nz, 1, nx), np.float32)}
    sb = {2: fx["sb2"].reshape(nz, 1, nx), 3: np.broadcast_to(fx["sb3"].reshape(1, 1, 1), (nz, 1, nx)).astype(np.float32)}
    off, A, rhs = be.assemble_scalar(box, u, bvel, fx["T_in"], sb, dt, meta["thermal_diffusivity"])
    assert rel_l2(rhs, fx["scalar_rhs"]) < 5e-7
    assert rel_l2(be.spmv(box, off, A, fx["T_out"]), fx["scalar_rhs"]) < 2e-6      # C_s T_out = rhs up to the solver tolerance
    src = be.buoyancy_source(box, fx["T_out"], 1.0)
    assert np.array_equal(src.reshape(3, -1), fx["vsrc"])
    offv, Av, _ = be.assemble(box, u, bvel, dt)
    rhsv, Sb = be.adv_rhs(box, u, bvel, dt, src)
    assert rel_l2(Av, fx["A"]) < 5e-7 and rel_l2(rhsv, fx["rhs"]) < 5e-7
    hb = be.hbya(box, u, fx["ustar"], offv, Av, Sb, dt, src)
    assert rel_l2(hb, fx["hbya0"]) < 1e-6
    assert rel_l2(be.divergence(box, hb, bvel), fx["div0"]) < 1e-6
    assert rel_l2(be.correct(box, hb, fx["p0"], Av), fx["u0"]) < 1e-6


def test_channel_initial_domain_glue_without_a_gpu(tmp_path):
    """InitialDomains3D for a box WITHOUT a passive scalar (turbulent channel): save -> directory layout -> pool -> per-environment
    assignment, on a stand-in solver."""
    import torch
    from fluidgym_b200.box3d import Box3DDomain
    from fluidgym_b200.envs.tcf import TCF3DEnv
    from fluidgym_b200.grids import channel_vertex_grid
    e = object.__new__(TCF3DEnv)
    e.L, e.re_wall, e.x, e.grid_refinement_strength = float(np.pi), 180.0, 8, 2
    e.dom = Box3DDomain(channel_vertex_grid(2.0, np.pi, np.pi / 2, 8, 4, 2, 8), closed=(False, True, False), viscosity=1e-3)
    N, NB = e.dom.N, e.dom.NB
    rng = np.random.default_rng(0)

    class S:
        u, p, bvel = torch.from_numpy(rng.standard_normal((2, 3, N)).astype(np.float32)), torch.zeros(2, N), torch.zeros(2, 3, NB)
    e.solver, e.n_envs, e.device = S, 2, torch.device("cpu")
    e.initial_domains_path = str(tmp_path)
    e._np_rng = np.random.default_rng(1)
    assert e.initial_domain_id == "channel_flow3D_L3.14_Re180_Res8_Ref2"
    for idx in range(10):
        S.u[0, 0, 0] = float(idx)
        e.save_initial_domain(idx, env_index=0)
    S.u.zero_()
    idxs = e._load_initial_domains_on_reset(True)                  # one random index per environment
    assert len(idxs) == 2 and all(0 <= i < 10 for i in idxs)
    assert [float(S.u[k, 0, 0]) for k in range(2)] == [float(i) for i in idxs]
    e.test()
    with pytest.raises(RuntimeError, match="Initial domain not found"):
        e._load_initial_domains_on_reset(False)                    # no "test" split was written


def test_specification_with_per_cell_viscosity_matches_reference(tcf, golden):
    """Smagorinsky model on (C_smag = 0.1, van Driest): the reference's per-cell viscosity of the first substep (golden
    tcf32_sgs_substep0.npz) fed into the specification reproduces its predictor matrix diagonal, right-hand side and HbyA -- the diffusive
    face coefficient (alpha_P nu_P + alpha_N nu_N) / 2 and the wall term 2 alpha_b nu_P (K.cu:3697-3750, 3845, 4342, 5204)."""
    box = tcf[0]
    fx = golden("tcf32_sgs_substep0.npz")
    bvel = {2: fx["bvel2"].reshape(3, 32, 1, 32), 3: fx["bvel3"].reshape(3, 32, 1, 32)}
    dt, u, visc = float(fx["dt"][0]), fx["u_in"], fx["visc"]
    off, A, fl = be.assemble(box, u, bvel, dt, visc=visc)
    assert rel_l2(A, fx["A"]) < 5e-7
    rhs, Sb = be.adv_rhs(box, u, bvel, dt, fx["src"], visc=visc)
    assert rel_l2(rhs, fx["rhs"]) < 5e-7
    hb = be.hbya(box, u, fx["ustar"].reshape(3, *box.shape), off, A, Sb, dt, fx["src"])
    assert rel_l2(hb, fx["hbya0"]) < 5e-7
    # with the constant viscosity the same inputs are far from this golden: the test sees the term
    off0, A0, _ = be.assemble(box, u, bvel, dt)
    assert rel_l2(A0, fx["A"]) > 1e-3
