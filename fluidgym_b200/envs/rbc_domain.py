"""Domain description of the 2-D Rayleigh-Benard environments (``envs/rbc/rbc_env_base.py:190-278``):
one block, periodic in x, no-slip plates at -y/+y with Dirichlet temperatures (hot bottom = actuated
heaters, cold top), wall-refined orthogonal grid, temperature as passive scalar."""
from __future__ import annotations

import numpy as np
import torch

from ..domain import DomainSpec
from ..grids import wall_refined_ortho_grid


def make_rbc_domain(rayleigh_number=8e4, prandtl_number=0.7, n_heaters=12, heater_width=8, aspect_ratio=1.0,
                    uniform_grid=False, H=1.0, T_hot=1.0, T_cold=0.0, resolution_scale_y=2.0, grid_base=1.02):
    ar = aspect_ratio * torch.pi
    nx = int(heater_width * n_heaters)
    ny = round(resolution_scale_y * nx / ar)
    L = H * ar
    nu = float(torch.tensor([(prandtl_number / rayleigh_number) ** 0.5], dtype=torch.float32)[0])
    kappa = float(torch.tensor([(rayleigh_number * prandtl_number) ** -0.5], dtype=torch.float32)[0])
    grid = wall_refined_ortho_grid(nx, ny, (0, -H / 2), (L, H / 2), ["-y", "+y"], 1.0 if uniform_grid else grid_base)
    dom = DomainSpec(nu, name="RBCDomain", scalar_viscosity=kappa)
    b = dom.create_block(grid, "RBCBlock")
    dom.make_periodic(b, 0)
    dom.close_boundary(b, "-y", scalar=np.full(nx, T_hot, np.float32))
    dom.close_boundary(b, "+y", scalar=np.full(nx, T_cold, np.float32))
    return dom, dict(nx=nx, ny=ny, L=L, H=H, nu=nu, kappa=kappa)
