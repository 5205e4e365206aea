"""Helpers shared by the environments whose agents are lined up along the span (z) of an extruded domain (CylinderJet3D,
Airfoil3D; ``envs/util/obs_extraction.py:60-205`` of the reference), batched over a leading environment dimension.  Host-side
pieces, used by ``envs/cylinder3d.py``."""
from __future__ import annotations

import torch


def spanwise_sensor_voxels(xy_physical: torch.Tensor, n_sensors_z: int, H: float, L: float, render_shape) -> torch.Tensor:
    """[3, n_sensors_z * n_xy] integer voxel coordinates (x, y, z), z-major (jet_cylinder_env_3d.py:271-301,
    cylinder_env_base.py:436-449: the z coordinate is scaled with render_shape[1], as in the reference)."""
    sz = torch.linspace(-H / 2, H / 2, n_sensors_z + 1)[:-1] + H / (2 * n_sensors_z)
    n_xy = xy_physical.shape[1]
    pc = torch.stack([xy_physical[0].unsqueeze(0).expand(n_sensors_z, -1).T, xy_physical[1].unsqueeze(0).expand(n_sensors_z, -1).T,
                      sz.unsqueeze(1).expand(-1, n_xy).T]).clone()
    pc[0] = (pc[0] + 2.0) * ((render_shape[0] - 1) / (L - 2.0))
    pc[1] = (pc[1] + H / 2) * ((render_shape[1] - 1) / H)
    pc[2] = (pc[2] + H / 2) * ((render_shape[1] - 1) / H)
    gc = torch.round(pc).to(torch.int64)
    return torch.stack([gc[c].reshape(-1, n_sensors_z).T for c in range(3)]).flatten(start_dim=1)


def global_obs_from_samples(u_s: torch.Tensor, p_s: torch.Tensor, n_agents: int, n_sensors_per_agent: int, local_2d_obs: bool = False):
    """u_s [B, n_sensors, 3], p_s [B, n_sensors] sampled at ``spanwise_sensor_voxels`` -> the reference's global observation
    {"velocity": [B, n_agents, per_agent, 3, n_xy], "pressure": [B, n_agents, per_agent, n_xy]}.  NB the reference reshapes the
    [sensor, component] axes with a raw ``view`` (obs_extraction.py:134-135), i.e. the axis labelled "component" does not hold the
    components; reproduced."""
    B = u_s.shape[0]
    nsz = n_agents * n_sensors_per_agent
    nd = 2 if local_2d_obs else 3                                     # local_2d_obs: x / y velocity only (obs_extraction.py:121-139)
    v = u_s[:, :, :nd].contiguous().view(B, nsz, nd, -1).view(B, n_agents, n_sensors_per_agent, nd, -1)
    if local_2d_obs:
        v = v.permute(0, 1, 2, 4, 3)
    p = p_s.contiguous().view(B, nsz, -1).view(B, n_agents, n_sensors_per_agent, -1)
    return {"velocity": v, "pressure": p}


def local_obs_windows(global_obs: dict, local_obs_window: int) -> dict:
    """{k: [B, n_agents, ...]} -> {k: [B, n_agents, window, ...]}: circular windows of neighbouring agents centred on every agent
    (transform_global_to_local_obs_3d, obs_extraction.py:153-205)."""
    out = {}
    for k, v in global_obs.items():
        n = v.shape[1]
        idx = (torch.arange(n, device=v.device)[:, None] + torch.arange(local_obs_window, device=v.device)[None, :] - local_obs_window // 2) % n
        out[k] = v[:, idx]
    return out
