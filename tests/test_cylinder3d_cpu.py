"""CylinderJet3D environment (fluidgym_b200/envs/cylinder3d.py) on the CPU: the environment class runs unchanged on top of a
stand-in solver (tests/cpu_harness/extruded_standin.py) that executes the per-cell code of the CUDA kernels on the host with numpy
Krylov solvers, and is compared with the UNMODIFIED reference's CylinderJet3D-easy run at resolution 8 (tests/golden/cyl3d_env.npz,
cyl3d_substep*.npz; generated on a B200 by oracle/ref_harness.py).  Only the CUDA launch path is not covered here."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, rel_l2

torch = pytest.importorskip("torch")
sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_harness"))


@pytest.fixture(scope="module")
def compiled():
    from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
    spec = make_cylinder_domain(8)
    return spec, spec.prepare()


def _env(compiled, **kw):
    from extruded_standin import HostExtrudedPISO3D
    from fluidgym_b200.envs.cylinder3d import CylinderJet3DEnv
    return CylinderJet3DEnv(resolution=8, n_jets=8, device="cpu", compiled=compiled, solver_cls=HostExtrudedPISO3D, **kw)


def test_reset_reproduces_the_reference_reset_state(compiled, golden):
    """zero field + inflow -> outflow update with time step 1 -> projection (A = 1, 4 deferred iterations) -> zero action"""
    fx = golden("cyl3d_env.npz")
    env = _env(compiled)
    obs, info = env.reset(seed=42)
    s = env.solver
    assert np.abs(s.bvel[0].numpy() - fx["reset_bvel"]).max() < 2e-6                                 # observed 2.4e-7
    # the reference's projection stops unconverged after 1000 CG iterations per solve (residual 4.6e-6 > 5e-7): two correct
    # implementations agree to the truncation level only -- observed u 9.7e-5, p 4.3e-5, observations 3.8e-4 / 5.3e-4 abs
    assert rel_l2(s.u[0].numpy().reshape(3, 8, -1), fx["reset_u"]) < 3e-4
    assert rel_l2(s.p[0].numpy().reshape(8, -1), fx["reset_p"]) < 3e-4
    assert obs["velocity"].shape == (1, 8, 2, 3, 151) and obs["pressure"].shape == (1, 8, 2, 151)
    assert np.abs(obs["velocity"][0].numpy() - fx["reset_obs_velocity"]).max() < 1e-3
    assert np.abs(obs["pressure"][0].numpy() - fx["reset_obs_pressure"]).max() < 2e-3


def test_first_solver_steps_follow_the_reference_trace(compiled, golden):
    """From the reference's reset state with its first action: boundary values, time step and state of the first two solver
    steps (action smoothing, jets, both flux balances, outflow update, CFL plan, substep)."""
    fx, s0, s1 = golden("cyl3d_env.npz"), golden("cyl3d_substep0.npz"), golden("cyl3d_substep1.npz")
    env = _env(compiled, step_length=0.02)                           # two solver steps per env.step
    env.seed(0)
    env.set_state(fx["reset_u"], fx["reset_p"], fx["reset_bvel"])
    seen = []
    orig = env.solver.piso_substep

    def spy(dt):
        seen.append((env.solver.u[0].numpy().copy(), env.solver.bvel[0].numpy().copy(), dt.numpy().copy()))
        return orig(dt)

    env.solver.piso_substep = spy
    env.step(torch.from_numpy(fx["actions"][0])[None])
    assert len(seen) == 2 and env.last_substeps == 2
    for (u_in, bvel, dt), ref in zip(seen, (s0, s1)):
        assert abs(float(dt[0]) - float(ref["dt"][0])) < 1e-9
        assert np.abs(bvel - ref["bvel"]).max() < 2e-6
        assert rel_l2(u_in.reshape(3, 8, -1), ref["u_in"]) < 1e-5
    assert rel_l2(env.solver.u[0].numpy().reshape(3, 8, -1), s1["u1"]) < 2e-5
    assert rel_l2(env.solver.p[0].numpy().reshape(8, -1), s1["p1"]) < 5e-4


def test_env_step_reproduces_the_reference(compiled, golden):
    """One full env.step (25 solver steps) from the reference's reset state with its action: state, reward, drag / lift (total
    and per plane) and observations of the unmodified reference."""
    fx = golden("cyl3d_env.npz")
    env = _env(compiled)
    env.seed(0)
    env.set_state(fx["reset_u"], fx["reset_p"], fx["reset_bvel"])
    obs, reward, terminated, truncated, info = env.step(torch.from_numpy(fx["actions"][0])[None])
    assert env.last_substeps == 25 and terminated is False and truncated is False
    s = env.solver
    eu, ep = rel_l2(s.u[0].numpy().reshape(3, 8, -1), fx["env0_u"]), rel_l2(s.p[0].numpy().reshape(8, -1), fx["env0_p"])
    print("CylinderJet3D env.step on the CPU stand-in: u", eu, "p", ep, "reward", float(reward[0]), "ref", float(fx["step0_reward"]),
          "drag", float(info["drag"][0]), float(fx["step0_info_drag"]), "lift", float(info["lift"][0]), float(fx["step0_info_lift"]))
    assert eu < 1e-4 and ep < 2e-3
    assert np.abs(s.bvel[0].numpy() - fx["env0_bvel"]).max() < 1e-4
    assert abs(float(reward[0]) - float(fx["step0_reward"])) < 1e-4 * max(1.0, abs(float(fx["step0_reward"])))
    assert abs(float(info["drag"][0]) - float(fx["step0_info_drag"])) < 1e-4 * max(1.0, abs(float(fx["step0_info_drag"])))
    assert abs(float(info["lift"][0]) - float(fx["step0_info_lift"])) < 1e-4
    assert np.abs(info["all_cds"][0].numpy() - fx["step0_info_all_cds"]).max() < 1e-4 * max(1.0, np.abs(fx["step0_info_all_cds"]).max())
    assert np.abs(info["all_cls"][0].numpy() - fx["step0_info_all_cls"]).max() < 1e-4
    assert np.abs(obs["velocity"][0].numpy() - fx["step0_obs_velocity"]).max() < 2e-4
    assert np.abs(obs["pressure"][0].numpy() - fx["step0_obs_pressure"]).max() < 2e-3 * max(1.0, np.abs(fx["step0_obs_pressure"]).max())


def _reference_local_obs(global_obs, n_jets, window, local_2d):
    """literal restatement of CylinderJetEnv3D._get_local_obs (jet_cylinder_env_3d.py:303-326) for one environment"""
    offset = window // 2
    out = {}
    for k, v in global_obs.items():
        shifted = torch.roll(v, shifts=offset, dims=0)
        lst = []
        for _ in range(n_jets):
            w = shifted[:window]
            if local_2d:
                w = w.squeeze()
            lst.append(w)
            shifted = torch.roll(shifted, shifts=-1, dims=0)
        out[k] = torch.stack(lst, dim=0)
    return out


@pytest.mark.parametrize("local_2d", [False, True])
def test_multi_agent_interface(compiled, golden, local_2d):
    """use_marl: one agent per jet; local observation windows (incl. local_2d_obs), local / global reward mix, spaces and the
    reference's error strings; two environments with different actions in one batch."""
    from fluidgym_b200 import spaces
    fx = golden("cyl3d_env.npz")
    env = _env(compiled, n_envs=2, use_marl=True, local_2d_obs=local_2d, step_length=0.01, local_reward_weight=0.8, cd_ref=3.0)
    assert env.n_agents == 8 and env.action_space.shape == (1,)
    with pytest.raises(RuntimeError, match="must be reset before stepping"):
        env.step(torch.zeros(2, 8, 1))
    env.seed(1)
    env.set_state(fx["reset_u"], fx["reset_p"], fx["reset_bvel"])
    with pytest.raises(ValueError, match="does not match expected shape"):
        env.step(torch.zeros(2, 8))
    a = env.sample_action()
    assert a.shape == (2, 8, 1) and float(a.abs().max()) <= 1.0
    a[1] = torch.from_numpy(fx["actions"][0])
    obs, reward, terminated, truncated, info = env.step(a)
    space = env.observation_space
    assert isinstance(space, spaces.Dict)
    for k, sub in space.spaces.items():
        assert tuple(obs[k].shape) == (2, 8) + tuple(sub.shape), (k, obs[k].shape, sub.shape)
    if local_2d:
        assert tuple(obs["velocity"].shape[2:]) == (151, 2)
    # windows = the reference's roll loop applied to our global observation, environment by environment
    g = env._get_global_obs()
    for e in range(2):
        ref = _reference_local_obs({k: v[e] for k, v in g.items()}, 8, env.local_obs_window, local_2d)
        for k in ref:
            assert torch.equal(obs[k][e], ref[k]), k
    # rewards (jet_cylinder_env_3d.py:454-486)
    cds, cls_ = env._drag_and_lift()
    glob = 3.0 - cds.sum(1) / 4.0 - (cls_.sum(1) / 4.0).abs()
    assert torch.allclose(info["global_reward"], glob, atol=1e-6) and "all_cds" not in info
    loc = 3.0 - cds.view(2, 8, -1).sum(2) / 0.5 - (cls_.view(2, 8, -1).sum(2) / 0.5).abs()
    assert reward.shape == (2, 8) and torch.allclose(reward, 0.8 * loc + 0.2 * glob[:, None], atol=1e-5)
    assert not torch.allclose(env.solver.bvel[0], env.solver.bvel[1])          # different actions -> different jets
    assert torch.allclose(info["drag"], cds.sum(1) / 4.0, atol=1e-6)


def test_constructor_checks_and_spaces(compiled):
    from fluidgym_b200.envs.cylinder3d import CylinderJet3DEnv
    with pytest.raises(ValueError, match="evenly divides"):
        CylinderJet3DEnv(resolution=8, n_jets=3, device="cpu", compiled=compiled)
    with pytest.raises(ValueError, match="only supported in multi-agent mode"):
        CylinderJet3DEnv(resolution=8, n_jets=8, local_2d_obs=True, device="cpu", compiled=compiled)
    from fluidgym_b200 import native
    with pytest.raises(native.FGBError):                                         # the product path has no CPU solver
        if torch.cuda.is_available():
            raise native.FGBError("n/a")
        CylinderJet3DEnv(resolution=8, n_jets=8, device="cpu", compiled=compiled)
    env = _env(compiled)
    assert env.n_agents == 1 and env.action_space.shape == (8, 1)
    assert env.observation_space.spaces["velocity"].shape == (8, 2, 3, 151) and env.observation_space.spaces["pressure"].shape == (8, 2, 151)
    assert env.render_shape == (171, 32, 32) and env.n_sim_steps == 25 and env.nz_per_agent == 1
    with pytest.raises(ValueError, match="Seed must be provided"):
        env.reset()
    with pytest.raises(RuntimeError, match="seeded before sampling"):
        env.sample_action()


def test_adaptive_plan_splits_per_environment(compiled, golden):
    """single_step with a CFL limit that binds for one environment only: that environment takes several substeps whose sizes add
    up to dt, the other one takes a single step and keeps its state while it waits."""
    fx = golden("cyl3d_env.npz")
    env = _env(compiled, n_envs=2)
    env.set_state(fx["reset_u"], fx["reset_p"], fx["reset_bvel"])
    s = env.solver
    s.u[1] *= 3.0
    mv = s.max_velocity()
    assert float(mv[1]) > 2.5 * float(mv[0]) > 0
    cfl = float(mv[0]) * 0.0101                                               # env 0: one step of 0.01, env 1: three substeps
    dts = []
    orig = s.piso_substep

    def spy(dt):
        dts.append(dt.numpy().copy())
        s.u += 0.0                                                             # keep the stand-in cheap: no solve
        s.p[:] = float(len(dts))

    s.piso_substep = spy
    u0 = s.u[0].clone()
    rounds = s.single_step(0.01, cfl)
    assert rounds == 3 and len(dts) == 3
    assert abs(sum(float(d[1]) for d in dts) - 0.01) < 1e-8 and abs(float(dts[0][0]) - 0.01) < 1e-9
    assert torch.equal(s.u[0], u0) and float(s.p[0, 0]) == 1.0 and float(s.p[1, 0]) == 3.0
    s.piso_substep = orig


def test_unsupported_options_fail_loudly(compiled):
    import fluidgym_b200
    from extruded_standin import HostExtrudedPISO3D
    kw = dict(resolution=8, device="cpu", compiled=compiled, solver_cls=HostExtrudedPISO3D)
    env = fluidgym_b200.make("CylinderJet3D-easy-v0", load_initial_domain=True, initial_domains_path="/nonexistent", **kw)
    with pytest.raises(RuntimeError, match="Initial domain not found"):          # the reference's message (fluid_env.py:1076-1079)
        env.reset(seed=0)
    # differentiable=True is a supported option since round 2 (reverse mode of the extruded substep, GPU-tested against the reference's
    # gradients in tests/test_gpu_extruded.py); the constructor only records it
    assert fluidgym_b200.make("CylinderJet3D-easy-v0", differentiable=True, **kw).differentiable is True
    env = fluidgym_b200.make("CylinderJet3D-medium-v0", load_initial_domain=False, load_domain_statistics=False, **kw)
    assert env.reynolds_number == 250.0 and env.initial_domain_id == "cylinder_3D_Re250_Res8"


def test_rl_library_adapters_accept_the_spanwise_environment(compiled):
    """integration.py on CylinderJet3D (host logic only, the solver calls are stubbed out): SB3 VecEnv slots = environments x
    agents, PettingZoo per-agent dictionaries, gymnasium single-environment view."""
    from fluidgym_b200.integration import GymFluidEnv, PettingZooFluidEnv, VecFluidEnv

    def mk(**kw):
        e = _env(compiled, step_length=0.01, **kw)
        e.solver.piso_substep = lambda dt: None
        e.solver.make_divergence_free = lambda max_iter=1000: None
        return e

    for marl, slots, lead in ((False, 2, (8, 2)), (True, 16, (3, 2))):
        v = VecFluidEnv(mk(n_envs=2, use_marl=marl))
        v.seed(1)
        obs = v.reset()
        assert v.num_envs == slots and obs["velocity"].shape == (slots,) + lead + (3, 151) and obs["pressure"].shape == (slots,) + lead + (151,)
        o, r, d, info = v.step(np.zeros((slots,) + v.action_space.shape, np.float32))
        assert r.shape == (slots,) and d.shape == (slots,) and len(info) == slots
    pz = PettingZooFluidEnv(mk(n_envs=1, use_marl=True))
    o, _ = pz.reset(seed=3)
    assert len(pz.agents) == 8 and o[pz.agents[0]]["velocity"].shape == (3, 2, 3, 151)
    o, r, t, tr, _ = pz.step({a: np.zeros(1, np.float32) for a in pz.agents})
    assert set(r) == set(pz.agents)
    g = GymFluidEnv(mk(n_envs=1))
    o, _ = g.reset(seed=1)
    assert o["velocity"].shape == (8, 2, 3, 151)
    o, r, t, tr, _ = g.step(np.zeros((8, 1), np.float32))
    assert np.isfinite(float(r))


def test_on_disk_initial_domains_of_the_extruded_grid(compiled, golden, tmp_path):
    """load_initial_domain=True for CylinderJet3D: the reference's domain format with one more spatial axis, six boundaries per
    block (z pair PERIODIC) and the axes triple of the block connections (fluidgym_b200/domain_io.py::load_extruded_domain /
    save_extruded_domain).  Round trip of the reference's traced 3-D state through a file, and reset() loading it per environment.
    (The 3-D specifics of the format follow the reference's writer source; no reference-written 3-D multi-block file exists yet.)"""
    import json
    from fluidgym_b200.domain_io import load_extruded_domain
    fx = golden("cyl3d_env.npz")
    env = _env(compiled, n_envs=2, load_initial_domain=True, initial_domains_path=str(tmp_path))
    env.solver.make_divergence_free = lambda max_iter=1000: None                    # keep the loaded state recognisable
    env.set_state(fx["env0_u"], fx["env0_p"], fx["env0_bvel"])
    env.solver.u[1] *= 0.5
    p0 = env.save_initial_domain(0, env_index=0)
    p3 = env.save_initial_domain(3, env_index=1)
    assert p0.endswith(os.path.join("cylinder_3D_Re100_Res8", "0", "train"))
    d = json.load(open(p0 + ".json"))
    assert d["spatialDims"] == 3 and len(d["blocks"]) == 5 and all(len(b["boundaries"]) == 6 for b in d["blocks"])
    conn = [bd for b in d["blocks"] for bd in b["boundaries"] if bd["type"] == "CONNECTED"]
    assert len(conn) == 10 and all(len(c["axes"]) == 3 and 4 in c["axes"][1:] for c in conn)
    spec, zv, st = load_extruded_domain(p0)
    assert np.array_equal(st["u"], fx["env0_u"]) and np.array_equal(st["p"], fx["env0_p"]) and np.array_equal(st["bvel"], fx["env0_bvel"])
    assert np.allclose(zv, np.linspace(-2, 2, 9)) and all(np.array_equal(a.vertex, b.vertex) for a, b in zip(spec.blocks, env.spec.blocks))
    cd2 = spec.prepare()                                                            # the file compiles into the same tables
    assert np.array_equal(cd2.nbr, env.cd.nbr) and np.array_equal(cd2.fl_comp, env.cd.fl_comp)
    env.solver.u.zero_()
    env.reset(seed=1, randomize=False)                                              # index 0 for every environment
    u4 = env.solver.u.view(2, 3, 8, -1)
    assert torch.equal(u4[0], torch.from_numpy(fx["env0_u"])) and torch.equal(u4[1], u4[0])
    env.load_initial_domain(3, env_index=[1])
    assert torch.allclose(env.solver.u.view(2, 3, 8, -1)[1], 0.5 * torch.from_numpy(fx["env0_u"]))
    with pytest.raises(RuntimeError, match="Initial domain not found"):
        env.load_initial_domain(7)


def test_first_gpu_run_tool_dry_run_on_the_stand_in(tmp_path):
    """tools/extruded_check.py (the first GPU run of the extruded path, tests/test_gpu_extruded.py) executed here
    on the CPU stand-in: the tool's own code paths and bars are exercised before it ever meets a GPU."""
    import json
    import subprocess
    out = tmp_path / "check.json"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "extruded_check.py"), "--standin", "--skip-env-step", "--json", str(out)],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    v = json.load(open(out))
    assert v["ok"] and [s["stage"] for s in v["stages"]] == ["substep", "substep", "reset"]
