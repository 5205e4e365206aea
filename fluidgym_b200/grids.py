"""Vertex-grid generators for the flow-control domains (host side, setup time only).

Own restatement of the grid construction the reference performs with
``simulation/pict/data/shapes.py`` (``make_torus_2D`` :679-766, ``generate_grid_vertices_2D``
:450-507 + ``interpolate_vertices_from_borders_2D`` :266-355, ``make_weights_exp`` :398-411,
``make_wall_refined_ortho_grid`` :585-638) and ``envs/cylinder/grid.py:85-289``.  All arithmetic is
float64 numpy, rounded once to float32 at the end, exactly as the reference stores its vertices, so
the metric tensors derived from them agree bit-for-bit (checked in tests/test_grids.py against
vertex fixtures generated from the reference's own generators).

Layout of every returned grid: ``[2, Y+1, X+1]`` float32, channel 0 = x, channel 1 = y.
"""
from __future__ import annotations

import numpy as np


def weights_exp(res: int, base: float, refinement: str) -> list:
    """Cumulative cell-size weights in [0,1] with geometric growth (shapes.py:398-411)."""
    exponents = list(range(res))
    if refinement == "END":
        exponents.reverse()
    elif refinement == "BOTH":
        exponents = exponents[: res // 2] + list(reversed(exponents))[res // 2:]
    sizes = [base ** e for e in exponents]
    total = np.sum(sizes)
    return [0] + [w / total for w in np.cumsum(sizes)]


def _lerp_border(lo, hi, weights):
    return [(lo[0] * (1 - w) + hi[0] * w, lo[1] * (1 - w) + hi[1] * w) for w in weights]


def transfinite_grid(res, corners, borders=None, x_weights=None, y_weights=None) -> np.ndarray:
    """Vertices of a quad patch from its 4 borders (shapes.py:266-355, 450-507).

    res = [ny, nx] vertex counts; corners = [(-x,-y), (+x,-y), (-x,+y), (+x,+y)];
    borders = [-x, +x, -y, +y] vertex lists or None (straight line between the corners).
    ``x_weights`` parametrises the -x/+x borders (runs along y), ``y_weights`` the -y/+y ones.
    """
    ny, nx = res
    if borders is None:
        borders = [None] * 4
    borders = list(borders)
    if x_weights is None:
        x_weights = [i / (ny - 1) for i in range(ny)]
    if y_weights is None:
        y_weights = [i / (nx - 1) for i in range(nx)]
    c_of = {0: (0, 2), 1: (1, 3), 2: (0, 1), 3: (2, 3)}
    for b in range(4):
        if borders[b] is None:
            w = x_weights if b < 2 else y_weights
            borders[b] = _lerp_border(corners[c_of[b][0]], corners[c_of[b][1]], w)
    bx0, bx1, by0, by1 = [np.asarray(b, dtype=np.float64) for b in borders]
    assert len(bx0) == ny and len(bx1) == ny and len(by0) == nx and len(by1) == nx
    grid = np.zeros((2, ny, nx), dtype=np.float64)
    for y in range(ny):
        wu = x_weights[y]
        wl = 1 - x_weights[y]
        row = by0 * wl + by1 * wu  # [nx, 2]
        start = by0[0] * wl + by1[0] * wu
        end = by0[-1] * wl + by1[-1] * wu
        size = end - start
        target = bx1[y] - bx0[y]
        if np.any(np.isclose(size, 0)):
            diff = target - size
            frac = (np.arange(nx) / (nx - 1))[:, None]
            val = row - start + diff * frac + bx0[y]
        else:
            val = (row - start) * (target / size) + bx0[y]
        grid[0, y] = val[:, 0]
        grid[1, y] = val[:, 1]
    return grid.astype(np.float32)


def torus_2d(res: int, r1: float, r2: float, start_angle: float, angle: float) -> np.ndarray:
    """Annulus sector, x along the angle, y along the radius with ~square cells (shapes.py:679-766)."""
    start_angle = start_angle % 360
    x = res + 1
    rad_step = np.deg2rad(angle / (x - 1))
    start_rad = np.deg2rad(start_angle)
    end_rad = start_rad + np.deg2rad(angle)
    corners = [
        (np.cos(start_rad) * r1, np.sin(start_rad) * r1),
        (np.cos(end_rad) * r1, np.sin(end_rad) * r1),
        (np.cos(start_rad) * r2, np.sin(start_rad) * r2),
        (np.cos(end_rad) * r2, np.sin(end_rad) * r2),
    ]
    lower = [(np.cos(start_rad + rad_step * i) * r1, np.sin(start_rad + rad_step * i) * r1) for i in range(x)]
    upper = [(np.cos(start_rad + rad_step * i) * r2, np.sin(start_rad + rad_step * i) * r2) for i in range(x)]
    r = r2 - r1
    sizes = []
    d = r1
    y = 1
    width_scale = 2 * np.pi / x * (abs(angle) / 360)
    while d < r2:
        width = d * width_scale
        sizes.append(width)
        d += width
        y += 1
    scale = (d - r1) / r
    sizes = [w / scale for w in sizes]
    x_weights = [0] + [w / r for w in np.cumsum(sizes)]
    return transfinite_grid([y, x], corners, [None, None, lower, upper], x_weights=x_weights)


def wall_refined_ortho_grid(res_x, res_y, corner_lower, corner_upper, wall_refinement, base) -> np.ndarray:
    """Axis-aligned box with geometric refinement towards the named walls (shapes.py:585-638)."""
    corners = [tuple(corner_lower), (corner_upper[0], corner_lower[1]),
               (corner_lower[0], corner_upper[1]), tuple(corner_upper)]

    def pick(lo, hi, res):
        if lo in wall_refinement:
            return weights_exp(res, base, "BOTH" if hi in wall_refinement else "START")
        if hi in wall_refinement:
            return weights_exp(res, base, "END")
        return None

    y_w = pick("-x", "+x", res_x)
    x_w = pick("-y", "+y", res_y)
    return transfinite_grid([res_y + 1, res_x + 1], corners, None, x_weights=x_w, y_weights=y_w)


def cylinder_vertex_grids(resolution: int = 24, domain_height: float = 4.1, domain_length: float = 22.0,
                          cylinder_radius: float = 0.5, cylinder_offset_y: float = 0.05,
                          circle_thickness: float = 0.5, quad_thickness_x: float = 1.0,
                          refinement_base: float = 0.95, refinement_axes=("+y", "-y")) -> list:
    """The five vertex grids [left, bottom, top, right, wake] of the vortex-street domain.

    Follows ``envs/cylinder/grid.py:85-289`` with the parameters of ``CylinderEnvBase._get_domain``
    (``cylinder_env_base.py:247-263``).  All grids are oriented x right / y up.
    """
    quad_thickness_y = quad_thickness_x + cylinder_offset_y
    x_min = -(cylinder_radius + circle_thickness + quad_thickness_x)
    x_max = domain_length + x_min
    if domain_height != 2 * cylinder_radius + 2 * circle_thickness + 2 * quad_thickness_y:
        raise ValueError("domain_height does not match radius/thickness parameters")
    r1 = cylinder_radius
    r2 = r1 + circle_thickness
    c_top = torus_2d(resolution, r1, r2, 135, -90)
    c_right = np.flip(np.swapaxes(torus_2d(resolution, r1, r2, 45, -90), -1, -2), -2)
    c_bot = np.flip(torus_2d(resolution, r1, r2, -45, -90), (-2, -1))
    c_left = np.flip(np.swapaxes(torus_2d(resolution, r1, r2, -135, -90), -1, -2), -1)

    qx = cylinder_radius + circle_thickness + quad_thickness_x
    qy = cylinder_radius + circle_thickness + quad_thickness_y
    qy_top = qy + cylinder_offset_y
    qy_bot = qy - cylinder_offset_y
    qi = np.sin(np.deg2rad(45)) * r2
    res_radial_circle = c_top.shape[-2] - 1
    q_ang = resolution + 1
    q_rad = int(np.ceil(quad_thickness_y / circle_thickness * res_radial_circle))

    def border(t):  # [2, n] -> list of (x, y) in float32 precision like the reference (.tolist())
        return np.moveaxis(np.ascontiguousarray(t), 0, 1).tolist()

    q_top = transfinite_grid([q_rad, q_ang], [(-qi, qi), (qi, qi), (-qx, qy_top), (qx, qy_top)],
                             [None, None, border(c_top[:, -1, :]), None])
    q_bot = transfinite_grid([q_rad, q_ang], [(-qx, -qy_bot), (qx, -qy_bot), (-qi, -qi), (qi, -qi)],
                             [None, None, None, border(c_bot[:, 0, :])])
    xw = weights_exp(q_ang - 1, refinement_base, "BOTH")
    q_right = transfinite_grid([q_ang, q_rad], [(qi, -qi), (qx, -qy_bot), (qi, qi), (qx, qy_top)],
                               [border(c_right[:, :, -1]), None, None, None], x_weights=xw)
    q_left = transfinite_grid([q_ang, q_rad], [(-qx, -qy_bot), (-qi, -qi), (-qx, qy_top), (-qi, qi)],
                              [None, border(c_left[:, :, 0]), None, None])

    left = np.concatenate([q_left[:, :, :-1], c_left], axis=-1)
    top = np.concatenate([c_top[:, :-1, :], q_top], axis=-2)
    right = np.concatenate([c_right[:, :, :-1], q_right], axis=-1)
    bottom = np.concatenate([q_bot[:, :-1, :], c_bot], axis=-2)
    res_wake = int(q_rad / quad_thickness_y * 18)
    wake = wall_refined_ortho_grid(res_wake, resolution, (-1 * x_min, -qy_bot), (x_max, qy_top),
                                   list(refinement_axes), refinement_base)
    return [np.ascontiguousarray(g, dtype=np.float32) for g in (left, bottom, top, right, wake)]


def uniform_box_grid(res_x: int, res_y: int, lower, upper) -> np.ndarray:
    """Uniform axis-aligned box (Rayleigh-Benard style single block)."""
    xs = np.linspace(lower[0], upper[0], res_x + 1, dtype=np.float64)
    ys = np.linspace(lower[1], upper[1], res_y + 1, dtype=np.float64)
    g = np.zeros((2, res_y + 1, res_x + 1))
    g[0] = xs[None, :]
    g[1] = ys[:, None]
    return g.astype(np.float32)


def channel_y_weights(N: int = 1, ny_half: int = 48) -> list:
    """Wall-normal vertex weights of the channel grid: geometric growth 1.2^(N/2) from both walls
    (envs/tcf/grid.py:15-31)."""
    ny = 2 * (ny_half // N)
    r = 1.2 ** (N / 2)
    h0 = 0.5 * (1 - r) / (1 - r ** (ny / 2))
    h = 0
    y = [0.0] * ny
    for i in range((ny - 2) // 2):
        h += h0 * (r ** i)
        y[i] = h
        y[ny - i - 2] = 1 - h
    y[ny // 2 - 1] = 0.5
    y[ny - 1] = 1.0
    return [0] + y


def channel_vertex_grid(H: float, L: float, D: float, x: int, y_half: int, yN: int, z: int) -> np.ndarray:
    """[3, z+1, y+1, x+1] float32 vertices of the channel block: uniform in x and z, refined towards both walls in y
    (envs/tcf/grid.py:34-72 + shapes.extrude_grid_z, shapes.py:641-680)."""
    delta = H / 2
    yw = channel_y_weights(N=yN, ny_half=y_half * yN)
    ny = len(yw) - 1
    g2 = transfinite_grid([ny + 1, x + 1], [(-L / 2, -delta), (L / 2, -delta), (-L / 2, delta), (L / 2, delta)], None, x_weights=yw)
    zs = np.asarray([(-D / 2) * (1 - w) + (D / 2) * w for w in (k / z for k in range(z + 1))], dtype=np.float32)
    out = np.zeros((3, z + 1, ny + 1, x + 1), dtype=np.float32)
    out[0] = g2[0][None]
    out[1] = g2[1][None]
    out[2] = zs[:, None, None]
    return out
