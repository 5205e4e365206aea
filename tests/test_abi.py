"""The C-ABI shared library loads and exports every symbol include/fluidgym_b200.h declares
(no compute calls: there is no GPU in the CPU test environment)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "fluidgym_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fgb_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = _declared()
    for n in ("fgb_piso_substep", "fgb_sim_step", "fgb_solve_pressure", "fgb_setup_advection", "fgb_batch_create"):
        assert n in names


def test_library_exports_every_declared_symbol():
    from fluidgym_b200 import build, native
    path = build.build()
    lib = ctypes.CDLL(path)
    for n in _declared():
        assert hasattr(lib, n), f"{n} declared in the header but not exported by {path}"
    assert sorted(native.EXPORTS) == _declared()
    assert lib.fgb_version() >= 1


def test_no_cpu_fallback_without_device():
    """Creating a batch without a CUDA device must fail loudly (never silently compute on the CPU)."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fluidgym_b200 import native
    from fluidgym_b200.domain import DomainSpec
    from fluidgym_b200.grids import uniform_box_grid
    from fluidgym_b200.solver import BatchedPISO
    spec = DomainSpec(0.01)
    b = spec.create_block(uniform_box_grid(4, 4, (0, 0), (1, 1)))
    spec.make_periodic(b, 0)
    spec.close_boundary(b, "-y")
    spec.close_boundary(b, "+y")
    with pytest.raises(native.FGBError):
        BatchedPISO(spec.prepare(), 2)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fluidgym_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(d, f)).read()
                assert "piso_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f
