import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope="session")
def cyl24():
    """Compiled cylinder domain (res 24) using the REFERENCE's transforms (bit-identical inputs for the
    op-level parity tests) + the spec."""
    from fluidgym_b200.domain import CompiledDomain, FIXED
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain

    spec = make_cylinder_domain(24)
    g = np.load(os.path.join(GOLDEN, "cyl24_geometry.npz"))
    T, bT = g["T"], g["bT"]
    transforms, btr = [], {}
    o = 0
    for b in spec.blocks:
        n = b.nx * b.ny
        transforms.append(T[o:o + n].reshape(b.ny, b.nx, 9))
        o += n
    o = 0
    for bi, b in enumerate(spec.blocks):
        for f in range(4):
            if b.bounds[f].type == FIXED:
                n = b.size(1 - (f >> 1))
                btr[(bi, f)] = bT[o:o + n]
                o += n
    cd = CompiledDomain(spec, transforms=transforms, btransforms=btr)
    return spec, cd


@pytest.fixture(scope="session")
def cyl24_own():
    """Same domain with transforms computed by the product itself."""
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain

    spec = make_cylinder_domain(24)
    return spec, spec.prepare()


def _with_reference_transforms(spec, geometry_file):
    from fluidgym_b200.domain import CompiledDomain, FIXED
    g = np.load(os.path.join(GOLDEN, geometry_file))
    T, bT = g["T"], g["bT"]
    transforms, btr = [], {}
    o = 0
    for b in spec.blocks:
        n = b.nx * b.ny
        transforms.append(T[o:o + n].reshape(b.ny, b.nx, 9))
        o += n
    o = 0
    for bi, b in enumerate(spec.blocks):
        for f in range(4):
            if b.bounds[f].type == FIXED:
                n = b.size(1 - (f >> 1))
                btr[(bi, f)] = bT[o:o + n]
                o += n
    return CompiledDomain(spec, transforms=transforms, btransforms=btr)


@pytest.fixture(scope="session")
def airfoil():
    """Compiled Airfoil2D-medium domain (6 blocks, 46 806 cells) with the REFERENCE's transforms."""
    from fluidgym_b200.envs.airfoil_domain import make_airfoil_domain
    spec = make_airfoil_domain()
    return spec, _with_reference_transforms(spec, "airfoil_geometry.npz")
