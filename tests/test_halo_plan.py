"""Static communication plan of the pushed-halo CG (fluidgym_b200.solver.halo_plan -> tables.cg_slot / cg_exp / cg_cnt):
a CPU emulation of what k_cg_cluster_mb<PUSH> does with it -- every CTA keeps its cells in slots [0, per) of its shared-memory
vector, exports listed cells into its neighbours' halo slots, and gathers all four stencil neighbours from LOCAL slots --
must reproduce x[nbr] for every cluster size the launcher picks (2: RBC, 4: cylinder-24, 8: cylinder-32, 16: airfoil)."""
import numpy as np
import pytest

from fluidgym_b200.solver import cluster_shape, halo_plan


def _domains():
    from fluidgym_b200.envs.airfoil_domain import make_airfoil_domain
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
    from fluidgym_b200.envs.rbc_domain import make_rbc_domain
    return {"rbc": lambda: make_rbc_domain()[0], "cyl24": lambda: make_cylinder_domain(24), "cyl32": lambda: make_cylinder_domain(32),
            "airfoil": make_airfoil_domain}


@pytest.mark.parametrize("name, cs_expected", [("rbc", 2), ("cyl24", 4), ("cyl32", 8), ("airfoil", 16)])
def test_pushed_halo_gather_reproduces_the_neighbour_table(name, cs_expected):
    cd = _domains()[name]().prepare()
    N, nbr = cd.N, np.asarray(cd.nbr)
    plan = halo_plan(nbr, N, cg_impl=6)
    assert plan is not None and plan["cs"] == cs_expected == cluster_shape(N)[0]
    cs, pad, per = plan["cs"], plan["pad"], -(-N // plan["cs"])
    assert per <= pad and plan["hmax"] % 2 == 0
    smem_floats = 2 * pad + plan["hmax"] + 2 * 16 * cs
    assert smem_floats * 4 + plan["emax"] * 8 + 64 <= 227 * 1024, "halo plan must fit in shared memory"
    rng = np.random.default_rng(0)
    x = rng.standard_normal(N).astype(np.float32)
    vs = np.zeros((cs, pad + plan["hmax"]), dtype=np.float32)          # per-CTA shared-memory vector: owned cells + halo slots
    for r in range(cs):
        own = x[r * per:min((r + 1) * per, N)]
        vs[r, :own.size] = own
    received = np.zeros(cs, dtype=np.int64)
    for r in range(cs):                                                  # publish(): st.async of the export list
        n_exp, n_halo = plan["cnt"][r]
        for e in range(n_exp):
            a, d = int(plan["exp"][r, e, 0]), int(plan["exp"][r, e, 1])
            dst, slot = (a >> 24) & 0xff, a & 0xffffff
            assert dst != r and slot < per and pad <= d < pad + plan["hmax"]
            vs[dst, d] = vs[r, slot]
            received[dst] += 1
    assert np.array_equal(received, plan["cnt"][:, 1]), "every CTA expects exactly the bytes its neighbours push"
    for f in range(4):                                                   # apply(): plain local gathers
        inner = nbr[f] >= 0
        cells = np.nonzero(inner)[0]
        owner = cells // per
        got = vs[owner, plan["slot"][f][cells]]
        assert np.array_equal(got, x[nbr[f][cells]]), f"face {f}"
