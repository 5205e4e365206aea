"""Sensor map (static sparse restatement of splat + normalise + hole filling, resampling.cu:191-364)
against observations returned by the reference's env.reset()/env.step() for the recorded states."""
import numpy as np

from fluidgym_b200.sensors import fill_levels, sensor_tables


def _sensor_px(res=24):
    import torch
    from fluidgym_b200.envs.cylinder import CylinderJet2DEnv
    env = CylinderJet2DEnv.__new__(CylinderJet2DEnv)
    env.resolution = res
    pc = CylinderJet2DEnv.sensor_locations_physical(env)
    rs = env.render_shape
    pc[0, :] += 2.0
    pc[0, :] *= (rs[0] - 1) / (22.0 - 2.0)
    pc[1, :] += 4.1 / 2
    pc[1, :] *= (rs[1] - 1) / 4.1
    return rs, torch.round(pc).to(torch.int32).numpy()


def test_sensor_map_reproduces_reference_observations(cyl24_own, golden):
    spec, cd = cyl24_own
    rs, px = _sensor_px()
    assert rs == (515, 96) and px.shape == (2, 151)
    idx, w = sensor_tables([b.vertex for b in spec.blocks], rs, px, 16)
    assert np.abs(w.sum(0) - 1).max() < 1e-5          # convex combinations
    for fx, u, p, ov, op in ((golden("cyl24_reset.npz"), "u", "p", "obs_velocity", "obs_pressure"),
                             (golden("cyl24_steps.npz"), "env0_u", "env0_p", "step0_obs_velocity", "step0_obs_pressure")):
        vel = np.stack([(w * fx[u][c][idx]).sum(0) for c in range(2)], axis=1)
        prs = (w * fx[p][idx]).sum(0)
        assert np.abs(vel - fx[ov]).max() < 2e-5
        assert np.abs(prs - fx[op]).max() < 2e-5


def test_fill_levels():
    v = np.zeros((5, 7), bool)
    v[2, 3] = True
    lv = fill_levels(v, 3)
    assert lv[2, 3] == 0 and lv[2, 4] == 1 and lv[1, 3] == 1 and lv[1, 4] == 2 and lv[0, 5] == -1
