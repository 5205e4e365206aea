"""Grid generators against vertex fixtures produced by the reference's own generators
(tests/golden/make_vertex_fixtures.py).  Bar: bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from fluidgym_b200.grids import cylinder_vertex_grids, weights_exp


@pytest.mark.parametrize("res", [8, 24, 32])
def test_cylinder_vertices_bit_exact(res):
    ref = np.load(os.path.join(GOLDEN, f"cylinder_vertices_res{res}.npz"))
    grids = cylinder_vertex_grids(res)
    assert len(grids) == 5
    for i, g in enumerate(grids):
        assert g.dtype == np.float32
        assert np.array_equal(g, ref[f"b{i}"]), f"block {i}"


def test_cell_counts_match_survey():
    # 4 x (24 x 37) + 445 x 24 = 14 232 cells; res 32 -> 23 424
    for res, n in ((24, 14232), (32, 23424)):
        g = cylinder_vertex_grids(res)
        assert sum((a.shape[1] - 1) * (a.shape[2] - 1) for a in g) == n


def test_weights_exp_properties():
    for ref in ("START", "END", "BOTH"):
        w = weights_exp(12, 0.95, ref)
        assert len(w) == 13 and w[0] == 0 and abs(w[-1] - 1) < 1e-12
        assert all(w[i] < w[i + 1] for i in range(12))
