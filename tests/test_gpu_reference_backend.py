"""The drop-in boundary executed: the UNMODIFIED reference environment (baseline/_ref, its own reset,
actuation, force integration and sensor rendering) steps on this library through ``fluidgym_b200.reference_backend.B200Simulation``
patched in at ``CylinderEnvBase._get_simulation`` (envs/cylinder/cylinder_env_base.py:303-332), and its ``env.step`` reproduces the
rewards / drag / observations of its stock run (tests/golden/cyl24_steps.npz, written by oracle/ref_harness.py)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "baseline", "_ref", "fluidgym")


@pytest.mark.skipif(not os.path.isdir(REF), reason="the unmodified reference is not installed under baseline/_ref (oracle/build_ref.sh)")
def test_reference_env_steps_on_the_b200_backend(golden):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shims
    ref_shims.install()
    import fluidgym
    from fluidgym.envs.cylinder.cylinder_env_base import CylinderEnvBase
    from fluidgym.simulation.extensions import PISOtorch
    from fluidgym_b200.reference_backend import B200Simulation, patch_reference_env

    st = golden("cyl24_steps.npz")
    rs = golden("cyl24_reset.npz")
    calls = {"solve": 0}
    orig_solve = PISOtorch.SolveLinear

    def counted(*a, **k):
        calls["solve"] += 1
        return orig_solve(*a, **k)

    orig = patch_reference_env(CylinderEnvBase, outflow=[(4, 1)])             # wake block, +x face (cylinder_env_base.py:277-300)
    PISOtorch.SolveLinear = counted
    try:
        env = fluidgym.make("CylinderJet2D-easy-v0", load_initial_domain=False, load_domain_statistics=False, randomize_initial_state=False)
        obs0, _ = env.reset(seed=42)
        assert isinstance(env._sim, B200Simulation)
        # reset is the reference's own code (projection by its solver): the golden reset state
        u0 = torch.cat([b.velocity.reshape(2, -1) for b in env._domain.getBlocks()], dim=1).cpu().numpy()
        assert np.abs(u0 - rs["u"]).max() < 1e-5
        n_reset_solves = calls["solve"]
        assert n_reset_solves > 0
        # compiled tables from the live Domain == tables from our own grid generator
        from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
        own = make_cylinder_domain(24).prepare()
        assert np.array_equal(env._sim.cd.nbr, own.nbr)
        assert np.abs(env._sim.cd.Wp - own.Wp).max() < 1e-5 * np.abs(own.Wp).max()
        errs = []
        for k in range(2):
            act = torch.full_like(env._zero_action, float(st["actions"][k].ravel()[0]))
            obs, reward, term, trunc, info = env.step(act)
            errs.append(dict(reward=abs(float(reward) - float(st[f"step{k}_reward"])) / abs(float(st[f"step{k}_reward"])),
                             drag=abs(float(info["drag"]) - float(st[f"step{k}_info_drag"])) / abs(float(st[f"step{k}_info_drag"])),
                             obs_velocity=float(np.abs(obs["velocity"].cpu().numpy() - st[f"step{k}_obs_velocity"]).max()),
                             obs_pressure=float(np.abs(obs["pressure"].cpu().numpy() - st[f"step{k}_obs_pressure"]).max())))
        print("reference env on the B200 backend vs its stock run:", errs, "substeps", env._sim.substeps)
        assert calls["solve"] == n_reset_solves            # no stock Krylov solve during env.step: the fields came from this library
        assert env._sim.substeps >= 50
        for e in errs:
            assert e["reward"] < 1e-4 and e["drag"] < 1e-4 and e["obs_velocity"] < 1e-3 and e["obs_pressure"] < 5e-3
    finally:
        CylinderEnvBase._get_simulation = orig
        PISOtorch.SolveLinear = orig_solve
