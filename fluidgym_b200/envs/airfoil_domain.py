"""Domain description of the 2-D NACA 0012 airfoil environments.

Mirrors ``make_airfoil_domain`` (``envs/airfoil/grid.py:243-716``, resolution_div = 1): a C-type mesh of six
blocks -- inflow box (left), three body-fitted blocks wrapped around the airfoil (front, top, bottom; wall-normal
resolution 96 with geometric refinement 0.97) and two wake blocks whose streamwise spacing grows by
``tail_grow_mul`` from the trailing-edge cell size -- parabolic inflow on ``left:-x``, advective outflow on both
wake ``+x`` faces, no-slip elsewhere.  The float32 steps of the surface construction (rotation, normals,
nearest-ray search) are done in float32 torch like the reference so that block sizes and vertices come out
identical; the transfinite interpolation itself is ``grids.transfinite_grid``.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from ..domain import DomainSpec
from ..grids import transfinite_grid, weights_exp
from .cylinder_domain import inflow_profile

LEFT, FRONT, TOP, BOT, TAIL_UPPER, TAIL_LOWER = range(6)
JET_CENTERS = (0.2, 0.4, 0.6)      # envs/airfoil/grid.py:14-15
JET_WIDTH = 0.08
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "naca0012_sharp.npy")


def airfoil_polyline(attack_angle_deg: float) -> torch.Tensor:
    """Surface points [2, 160] float32, rotated by the angle of attack (grid.py:49-99)."""
    pts = torch.tensor(np.load(_DATA), dtype=torch.float32)          # [n, 2]
    if attack_angle_deg != 0.0:
        a = torch.tensor(-attack_angle_deg * np.pi / 180.0, dtype=torch.float32)
        rot = torch.tensor([[torch.cos(a), -torch.sin(a)], [torch.sin(a), torch.cos(a)]])
        pts = torch.matmul(pts, rot.T)
    return pts.T.contiguous()


def _point_line_distance(origin, direction, point):
    """Distance of ``point`` [2] from the lines origin_i + s direction_i ([2, n] each), float32."""
    d1, d2 = direction[0], direction[1]
    num = torch.abs(d1 * (origin[1] - point[1]) - (origin[0] - point[0]) * d2)
    return num / torch.sqrt(d1 * d1 + d2 * d2)


def _front_outline(normals, half_height, width_left, attack_angle_deg):
    """Where the surface normals of the nose section meet the rectangle (-width_left, +-half_height): the
    reference does not intersect rays but distributes points uniformly along the three rectangle sides, switching
    sides at the normal that points closest to the rectangle's corners (grid.py:149-240)."""
    ang = 180.0 - torch.atan2(normals[1], normals[0]) * 180.0 / np.pi
    ang = torch.where(ang < 180, ang, -360 + ang)
    ang = ang - attack_angle_deg
    corner = np.rad2deg(np.arctan2(half_height, width_left))
    upper = ang > 0
    a_up, a_lo = ang[upper], ang[~upper]
    n_up, n_lo = int(a_up.numel()), int(a_lo.numel())
    i_top = int(torch.argmin(torch.abs(a_up - corner)))
    i_bot = int(torch.argmin(torch.abs(a_lo + corner)))
    fill_top = n_up - i_top - 1            # points of the upper half that sit on the vertical side
    x_up = torch.cat([torch.linspace(0, -width_left, n_up - fill_top + 1)[1:-1], torch.full((fill_top + 1,), -width_left)])
    x_lo = torch.cat([torch.full((i_bot + 1,), -width_left), torch.linspace(-width_left, 0, n_lo - i_bot + 1)[1:-1]])
    on_top = n_up - fill_top
    y_up = torch.cat([torch.full((on_top,), half_height), torch.linspace(half_height, 0, n_up - on_top + 2)[1:-1]])
    y_lo = torch.cat([torch.linspace(0, -half_height, i_bot + 1)[:-1], torch.full((n_lo - i_bot,), -half_height)])
    outline = torch.stack([torch.cat([x_up, x_lo]), torch.cat([y_up, y_lo])])
    return outline, i_top, n_up + i_bot


def airfoil_vertex_grids(H: float = 1.4, L: float = 4.5, attack_angle_deg: float = 10.0, tail_grow_mul: float = 1.01):
    """The six vertex grids [left, front, top, bottom, tail_upper, tail_lower], each [2, Y+1, X+1] float32."""
    offset_left, front_width, hh = 1.5, 0.5, H / 2
    normal_res = 96
    w_start = weights_exp(normal_res - 1, 0.97, "START")
    w_end = weights_exp(normal_res - 1, 0.97, "END")
    f32 = torch.float32
    c = airfoil_polyline(attack_angle_deg)                           # [2, n]
    n = c.shape[1]
    len_x = torch.max(c[0])
    te = c[:, :1]
    te_spacing = torch.linalg.vector_norm(c[:, 1] - c[:, 0])
    te_ext = te + torch.stack([te_spacing, torch.zeros((), dtype=f32)]).reshape(2, 1)
    ext = torch.cat([te_ext, c, te_ext], dim=1)
    # outward normals from central differences of the closed polyline
    tang = ext[:, 2:] - ext[:, :-2]
    normals = torch.flip(tang, dims=(0,)) * torch.tensor([1.0, -1.0]).reshape(2, 1)
    normals = normals / torch.linalg.vector_norm(normals, dim=0)
    seg = torch.linalg.vector_norm(ext[:, 1:] - ext[:, :-1], dim=0)
    min_size = torch.min(seg).numpy().tolist()
    # wake spacing: geometric growth from the smallest surface segment until half the height is covered
    sizes, dist = [min_size], min_size
    while dist < hh:
        sizes.append(sizes[-1] * tail_grow_mul)
        dist = dist + sizes[-1]
    tail_w = [0] + (np.cumsum(sizes) / dist).tolist()
    p_top0 = torch.tensor([0.0, hh], dtype=f32)
    p_top1 = torch.stack([len_x, torch.tensor(hh, dtype=f32)])
    p_bot0 = torch.tensor([0.0, -hh], dtype=f32)
    p_bot1 = torch.stack([len_x, torch.tensor(-hh, dtype=f32)])
    half = n // 2
    i_top = int(torch.argmin(_point_line_distance(c[:, :half], normals[:, :half], p_top0)))
    i_bot = int(torch.argmin(_point_line_distance(c[:, half:], normals[:, half:], p_bot0))) + half
    # outer boundary, walked like the surface: top side (right to left), front, bottom side (left to right)
    step = (p_top0 - p_top1) / i_top
    outer_top = torch.stack([p_top1 + step * i for i in range(i_top + 1)], dim=1)
    n_bot = (n - 1) - i_bot
    step = (p_bot1 - p_bot0) / n_bot
    outer_bot = torch.stack([p_bot0 + step * i for i in range(n_bot + 1)], dim=1)
    outline, k_up, k_lo = _front_outline(normals[:, i_top + 1:i_bot], hh, front_width, attack_angle_deg)
    outer = torch.cat([outer_top, outline, outer_bot], dim=1)
    k_up, k_lo = k_up + 7, k_lo + 7                                  # resolution_div == 1 (grid.py:437-439)
    nb = outer_bot.shape[1]
    s_top = slice(0, nb + k_up + 3)
    s_front = slice(nb + k_up + 2, nb + k_lo + 3)
    s_bot = slice(nb + k_lo + 2, None)
    surf_top = torch.flip(c[:, s_top], dims=(1,))                     # left to right
    surf_front = torch.flip(c[:, s_front], dims=(1,))                 # lower junction to upper junction (along +y)
    surf_bot = c[:, s_bot]
    res_top, res_front, res_bot = surf_top.shape[1], surf_front.shape[1], surf_bot.shape[1]
    assert outer[:, s_top].shape[1] == res_top

    def pt(t, i):
        return (t[0, i].item(), t[1, i].item())

    def border(t):
        return t.T.clone().numpy().tolist()

    t0, t1, b0, b1 = pt(surf_top, 0), pt(surf_top, -1), pt(surf_bot, 0), pt(surf_bot, -1)
    left = transfinite_grid([res_front, int(0.75 * normal_res)],
                            [(-offset_left, -hh), (-front_width, -hh), (-offset_left, hh), (-front_width, hh)])
    top = transfinite_grid([normal_res, res_top], [t0, t1, (-front_width, hh), (t1[0], hh)],
                           [None, None, border(surf_top), None], x_weights=w_end)
    front = transfinite_grid([res_front, normal_res], [(-front_width, -hh), b0, (-front_width, hh), t0],
                             [None, border(surf_front), None, None], y_weights=w_start)
    bot = transfinite_grid([normal_res, res_bot], [(-front_width, -hh), (b1[0], -hh), b0, b1],
                           [None, None, None, border(surf_bot)], x_weights=w_start)
    tail_up = transfinite_grid([normal_res, len(tail_w)], [t1, (L, t1[1]), (t1[0], hh), (L, hh)], None,
                               x_weights=w_end, y_weights=tail_w)
    tail_lo = transfinite_grid([normal_res, len(tail_w)], [(b1[0], -hh), (L, -hh), b1, (L, b1[1])], None,
                               x_weights=w_start, y_weights=tail_w)
    return [np.ascontiguousarray(g, dtype=np.float32) for g in (left, front, top, bot, tail_up, tail_lo)]


def make_airfoil_domain(reynolds_number: float = 3e3, u_mean: float = 0.3, airfoil_length: float = 1.0, H: float = 1.4,
                        L: float = 4.5, attack_angle_deg: float = 10.0, tail_grow_mul: float = 1.01) -> DomainSpec:
    viscosity = float(torch.tensor([(u_mean * airfoil_length) / reynolds_number], dtype=torch.float32)[0])
    grids = airfoil_vertex_grids(H, L, attack_angle_deg, tail_grow_mul)
    dom = DomainSpec(viscosity, name="AirfoilDomain")
    names = ["LeftBlock", "AirfoilFront", "AirfoilTop", "AirfoilBot", "TailUpper", "TailLower"]
    left, front, top, bot, tail_up, tail_lo = [dom.create_block(g, nm) for g, nm in zip(grids, names)]
    ny_left = grids[LEFT].shape[1] - 1
    inflow = np.zeros((2, ny_left), dtype=np.float32)
    inflow[0] = inflow_profile(H, ny_left) * np.float32(u_mean)
    dom.close_boundary(left, "-x", inflow)
    for blk, face in ((left, "+y"), (left, "-y"), (top, "+y"), (tail_up, "+y"), (tail_lo, "-y"),     # tunnel walls
                      (front, "+x"), (top, "-y"), (bot, "+y")):                                     # airfoil surface
        dom.close_boundary(blk, face)
    for blk in (tail_up, tail_lo):                                                                  # advective outflow
        ny = grids[blk].shape[1] - 1
        out = np.zeros((2, ny), dtype=np.float32)
        out[0] = u_mean
        dom.close_boundary(blk, "+x", out)
    dom.connect(left, "+x", front, "-x", "-y")
    dom.connect(front, "+y", top, "-x", "+y")
    dom.connect(front, "-y", bot, "-x", "-y")
    dom.connect(top, "+x", tail_up, "-x", "-y")
    dom.connect(bot, "+x", tail_lo, "-x", "-y")
    dom.connect(tail_up, "-y", tail_lo, "+y", "-x")
    return dom
