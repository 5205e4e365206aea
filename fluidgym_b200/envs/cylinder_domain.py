"""Domain description of the 2-D cylinder vortex-street environments.

Mirrors ``make_vortex_street_domain`` (``envs/cylinder/grid.py:18-418``): four O-grid blocks around the
cylinder (left, top, right, bottom) plus the wake block, parabolic inflow on ``left:-x``
(``envs/util/profiles.py:35-86``), advective outflow on ``wake:+x`` and no-slip walls elsewhere.
"""
from __future__ import annotations

import numpy as np
import torch

from ..domain import DomainSpec
from ..grids import cylinder_vertex_grids

LEFT, TOP, RIGHT, BOTTOM, WAKE = range(5)


def inflow_profile(h: float, res_y: int) -> np.ndarray:
    """Parabolic profile with unit mean (profiles.py:64-71), float32 like the reference."""
    y = torch.linspace(-h / 2, h / 2, res_y, dtype=torch.float32)
    profile = 6 * (h / 2 - y) * (h / 2 + y) / h ** 2
    profile = profile / profile.mean()
    return profile.numpy()


def jet_profile(h: int) -> np.ndarray:
    """Parabolic jet profile with unit maximum (profiles.py:6-32)."""
    y = torch.linspace(-h / 2, h / 2, h, dtype=torch.float32)
    profile = 6 * (h / 2 - y) * (h / 2 + y) / h ** 2
    profile = profile / torch.max(profile)
    return profile.numpy()


def make_cylinder_domain(resolution: int = 24, reynolds_number: float = 100.0, u_mean: float = 1.0,
                         domain_height: float = 4.1, domain_length: float = 22.0,
                         cylinder_offset_y: float = 0.05) -> DomainSpec:
    viscosity = float(torch.tensor([u_mean / reynolds_number], dtype=torch.float32)[0])
    left, bottom, top, right, wake = cylinder_vertex_grids(resolution, domain_height, domain_length,
                                                          cylinder_offset_y=cylinder_offset_y)
    dom = DomainSpec(viscosity, name="CylinderDomain")
    b_left = dom.create_block(left, "BlockCylinderLeft")
    b_top = dom.create_block(top, "BlockCylinderTop")
    b_right = dom.create_block(right, "BlockCylinderRight")
    b_bottom = dom.create_block(bottom, "BlockCylinderBottom")
    b_wake = dom.create_block(wake, "BlockVortexStreet")
    inflow = np.zeros((2, resolution), dtype=np.float32)
    inflow[0] = inflow_profile(domain_height - 2 * cylinder_offset_y, resolution)
    dom.close_boundary(b_left, "-x", inflow)   # inflow
    dom.close_boundary(b_left, "+x")           # cylinder
    dom.close_boundary(b_top, "+y")            # wall
    dom.close_boundary(b_top, "-y")            # cylinder
    dom.close_boundary(b_right, "-x")          # cylinder
    dom.close_boundary(b_bottom, "-y")         # wall
    dom.close_boundary(b_bottom, "+y")         # cylinder
    dom.close_boundary(b_wake, "+y")
    dom.close_boundary(b_wake, "-y")
    dom.close_boundary(b_wake, "+x", inflow)   # advective outflow, initialised with the inflow profile
    dom.connect(b_left, "+y", b_top, "-x", "+y")
    dom.connect(b_left, "-y", b_bottom, "-x", "-y")
    dom.connect(b_right, "+y", b_top, "+x", "-y")
    dom.connect(b_right, "-y", b_bottom, "+x", "+y")
    dom.connect(b_right, "+x", b_wake, "-x", "-y")
    return dom
