"""Generates tests/golden/cylinder_vertices_res{R}.npz from the REFERENCE's own grid generators.

Run in the build container (needs /root/reference, which is pure python on this path):
    python tests/golden/make_vertex_fixtures.py
It imports ``simulation/pict/data/shapes.py`` standalone (only torch/numpy/scipy needed) and replays
the vertex construction of ``envs/cylinder/grid.py:85-289`` by executing that function body up to the
``grids`` list with a stub for the native ``PISOtorch`` module (no GPU, no Domain objects created).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/src/fluidgym"


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    shapes = load("ref_shapes", os.path.join(REF, "simulation/pict/data/shapes.py"))
    src = open(os.path.join(REF, "envs/cylinder/grid.py")).read()
    start = src.index("    res_z = circle_resolution_angular")
    end = src.index("    if ndims == 3:")
    body = "def build(shapes, np, torch, domain_height, domain_length, cylinder_radius, cylinder_offset_y, " \
           "circle_thickness, quad_thickness_x, circle_resolution_angular, vortex_street_refinement_base, " \
           "vortex_street_refinement_axes, dtype, debug=False):\n" + src[start:end] + "    return grids\n"
    ns = {}
    exec(body, ns)
    here = os.path.dirname(os.path.abspath(__file__))
    for res in (8, 24, 32):
        grids = ns["build"](shapes, np, torch, 4.1, 22.0, 0.5, 0.05, 0.5, 1.0, res, 0.95, ["+y", "-y"], torch.float32)
        out = {f"b{i}": g[0].numpy() for i, g in enumerate(grids)}
        np.savez_compressed(os.path.join(here, f"cylinder_vertices_res{res}.npz"), **out)
        print(res, [tuple(g.shape) for g in grids])


if __name__ == "__main__":
    main()
