"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).  Every call goes through the C ABI
(include/fluidgym_b200.h via ctypes); the checker is the CPU oracle and the reference golden trace."""
import os
import numpy as np
import pytest

from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _solver(cd, B, out_mask=None, **kw):
    from fluidgym_b200.solver import BatchedPISO
    return BatchedPISO(cd, B, out_mask=out_mask, **kw)


def _load_state(sol, fx, noise=None):
    u = torch.from_numpy(fx["u_in"]).cuda()
    p = torch.from_numpy(fx["presres_in"]).cuda()
    bv = torch.from_numpy(fx["bvel_in"]).cuda()
    sol.u.copy_(u.unsqueeze(0).expand_as(sol.u))
    sol.p.copy_(p.unsqueeze(0).expand_as(sol.p))
    sol.bvel.copy_(bv.unsqueeze(0).expand_as(sol.bvel))
    if noise is not None:
        gen = torch.Generator(device="cuda").manual_seed(1234)
        sol.u[1:] += noise * torch.randn(sol.u[1:].shape, device="cuda", generator=gen)


def test_library_loaded_and_versioned():
    from fluidgym_b200 import native
    L = native.load()
    assert L.fgb_version() >= 1


def test_ops_match_reference_trace(cyl24, golden):
    """Each native op on the reference's own inputs (substep 1 of the golden trace); tolerance 2e-6
    relative L2 per op (fp32 round-off; the reference itself is built with --use_fast_math)."""
    spec, cd = cyl24
    fx = golden("cyl24_substep1.npz")
    sol = _solver(cd, 2, cg_impl=0)
    _load_state(sol, fx)
    dt = float(fx["dt"][0])
    sol.setup_advection(dt)
    torch.cuda.synchronize()
    assert rel_l2(sol.buffer("A")[0].cpu().numpy(), fx["A"]) < 2e-6
    assert rel_l2(sol.buffer("rhs")[1].cpu().numpy(), fx["rhs"]) < 2e-6
    # C off-diagonals against the reference CSR (sorted by column)
    from oracle import ell_to_csr
    co = sol.buffer("Coff")[0].cpu().numpy()
    A = sol.buffer("A")[0].cpu().numpy()
    val = np.concatenate([A[:, None], co.T], axis=1)
    idx = np.concatenate([np.arange(cd.N)[:, None], np.where(cd.nbr >= 0, cd.nbr, -1).T], axis=1).astype(np.int32)
    cv, ci, cr = ell_to_csr(val, idx)
    assert np.array_equal(ci, fx["C_index"]) and np.array_equal(cr, fx["C_row"])
    assert rel_l2(cv, fx["C_value"]) < 2e-6
    sol.solve_advection(zero_init=True)
    torch.cuda.synchronize()
    its = sol.buffer("iters")[0].cpu().numpy()
    assert list(its[:2]) == list(fx["bicg_iters"])
    assert rel_l2(sol.buffer("ures")[0].cpu().numpy(), fx["ustar"]) < 5e-6
    sol.setup_pressure_matrix()
    po = sol.buffer("Poff")[1].cpu().numpy()
    pd = sol.buffer("Pdiag")[1].cpu().numpy()
    pv, pi, pr = ell_to_csr(np.concatenate([pd[:, None], po.T], axis=1), idx)
    assert np.array_equal(pi, fx["P_index"])
    assert rel_l2(pv, fx["P_value"]) < 2e-6
    sol.setup_pressure_rhs(dt)
    torch.cuda.synchronize()
    assert rel_l2(sol.buffer("hbya")[0].cpu().numpy(), fx["hbya0"]) < 5e-6
    assert np.abs(sol.buffer("div")[0].cpu().numpy() - fx["div0"]).max() < 2e-6 * np.abs(fx["div0"]).max() + 1e-6
    # corrector with the reference's pressure
    pref = torch.from_numpy(fx["p0"]).cuda().unsqueeze(0).repeat(2, 1).contiguous()
    sol.correct_velocity(p=pref)
    torch.cuda.synchronize()
    assert rel_l2(sol.buffer("ures")[1].cpu().numpy(), fx["u0"]) < 2e-6


@pytest.mark.parametrize("cg_impl", [0, 1, 2, 3, 4, 5, 6, 7, 8, 11, 12])
def test_pressure_solve_matches_oracle(cyl24, golden, cg_impl):
    """CG: same algorithm, same stopping rule -> iteration count within 2% of the oracle's and of the
    reference's, residual below tolerance, solution within the tolerance ball (5e-4 relative; the
    reference's own run-to-run reproducibility of this 2500-iteration unpreconditioned solve)."""
    from oracle import Oracle
    spec, cd = cyl24
    fx = golden("cyl24_substep1.npz")
    sol = _solver(cd, 3, cg_impl=cg_impl)
    _load_state(sol, fx)
    dt = float(fx["dt"][0])
    sol.setup_advection(dt)
    sol.solve_advection()
    sol.setup_pressure_matrix()
    sol.setup_pressure_rhs(dt)
    div = sol.buffer("div")[0].cpu().numpy().copy()
    sol.solve_pressure(zero_init=True)
    torch.cuda.synchronize()
    its = sol.buffer("iters").cpu().numpy()
    res = sol.buffer("resid").cpu().numpy()
    assert (res[:, 2] < 1e-5).all()
    assert abs(int(its[0, 2]) - int(fx["cg_iters"][0])) <= 0.02 * fx["cg_iters"][0] + 2
    assert (its[:, 2] == its[0, 2]).all()          # identical environments -> identical iteration counts
    p = sol.p.cpu().numpy()
    assert np.array_equal(p[0], p[1]) and np.array_equal(p[0], p[2])   # deterministic
    assert abs(p[0].mean()) < 1e-4 * np.abs(p[0]).max()
    assert rel_l2(p[0], fx["p0"]) < (5e-4 if cg_impl != 2 else 1e-3)   # variant 2 reorders the A*p recurrence
    # (the mean is removed after the solve, SIM.py:1922-1925; P has rows that do not sum to zero at
    #  one-sided non-orthogonal corners, so the shifted iterate is not re-checked against P x = rhs)
    del div


@pytest.mark.parametrize("cg_impl", [0, 1, 2, 3, 4, 5, 6, 7, 8, 11, 12])
def test_substep_matches_reference(cyl24, golden, cg_impl):
    """One full PISO substep from the reference's state.  u: 2e-4, p: 1e-3 relative L2 (bounded by the
    CG tolerance ball, see DESIGN.md 'parity'); batch entries with perturbed states are checked against
    the CPU oracle run on the same perturbed input."""
    from oracle import Oracle
    spec, cd = cyl24
    for name, tol_u, tol_p in (("cyl24_substep1.npz", 2e-4, 1e-3), ("cyl24_substep0.npz", 2e-4, 1e-3)):
        fx = golden(name)
        sol = _solver(cd, 3, cg_impl=cg_impl)
        _load_state(sol, fx, noise=1e-3)
        u_in = sol.u.cpu().numpy().copy()
        dt = float(fx["dt"][0])
        sol.piso_substep(dt)
        torch.cuda.synchronize()
        u = sol.u.cpu().numpy()
        p = sol.p.cpu().numpy()
        assert rel_l2(u[0], fx["u1"]) < tol_u
        assert rel_l2(p[0], fx["p1"]) < tol_p
        orc = Oracle(cd.sizes, cd.btype, cd.bconn, cd.T, cd.bT, fx["bvel_in"], float(cd.visc))
        uo = u_in[2].copy()
        po = fx["presres_in"].astype(np.float32).copy()
        orc.substep(uo, po, dt)
        assert rel_l2(u[2], uo) < tol_u
        assert rel_l2(p[2], po) < tol_p


def test_sim_step_and_outflow(cyl24, golden):
    """Adaptive single_step incl. the advective-outflow hook against the reference's first sim step of
    env.step (jets at control = 0.05)."""
    spec, cd = cyl24
    from fluidgym_b200.envs.cylinder import CylinderJet2DEnv
    env = CylinderJet2DEnv(n_envs=2, compiled=(spec, cd))
    rs = golden("cyl24_reset.npz")
    st = golden("cyl24_steps.npz")
    env.set_state(rs["u"], rs["presres"], rs["bvel"])
    action = torch.tensor(st["actions"][0], device="cuda").reshape(1, 1).repeat(2, 1)
    env._apply_action(action)
    n = env.solver.single_step(env.dt, env.cfl, char_vel=(1.0, 0.0))
    torch.cuda.synchronize()
    assert n == 1
    assert rel_l2(env.solver.u[0].cpu().numpy(), st["sim0_u"]) < 2e-4
    assert rel_l2(env.solver.p[1].cpu().numpy(), st["sim0_p"]) < 1e-3


def test_reset_matches_reference(cyl24, golden):
    """reset(): initial field + make_divergence_free (CG capped at 1000 iterations -> best iterate)."""
    spec, cd = cyl24
    from fluidgym_b200.envs.cylinder import CylinderJet2DEnv
    env = CylinderJet2DEnv(n_envs=2, compiled=(spec, cd))
    env.reset(seed=42)
    rs = golden("cyl24_reset.npz")
    torch.cuda.synchronize()
    assert rel_l2(env.solver.bvel[0].cpu().numpy(), rs["bvel"]) < 1e-5
    assert rel_l2(env.solver.u[0].cpu().numpy(), rs["u"]) < 5e-3
    assert rel_l2(env.solver.p[0].cpu().numpy(), rs["presres"]) < 5e-3


@pytest.mark.parametrize("cg_impl", [6, 11, 12])
def test_env_step_matches_reference(cyl24, golden, cg_impl):
    """env.step from the reference's reset state: drag/lift, reward, sensors after 25 sim steps."""
    spec, cd = cyl24
    from fluidgym_b200.envs.cylinder import CylinderJet2DEnv
    env = CylinderJet2DEnv(n_envs=2, compiled=(spec, cd), cg_impl=cg_impl)
    rs = golden("cyl24_reset.npz")
    st = golden("cyl24_steps.npz")
    env.reset(seed=42)
    env.set_state(rs["u"], rs["presres"], rs["bvel"])
    action = torch.tensor(st["actions"][0], device="cuda").reshape(1, 1).repeat(2, 1)
    obs, reward, term, trunc, info = env.step(action)
    torch.cuda.synchronize()
    assert rel_l2(env.solver.u[0].cpu().numpy(), st["env0_u"]) < 1e-3
    assert abs(float(info["drag"][0]) - float(st["step0_info_drag"])) < 1e-3 * abs(float(st["step0_info_drag"])) + 1e-4
    assert abs(float(info["lift"][0]) - float(st["step0_info_lift"])) < 1e-3 * abs(float(st["step0_info_drag"])) + 1e-4
    assert abs(float(reward[0]) - float(st["step0_reward"])) < 1e-3 * abs(float(st["step0_reward"])) + 1e-4
    assert obs["velocity"].shape == (2, 151, 2) and obs["pressure"].shape == (2, 151)
    assert np.abs(obs["velocity"][0].cpu().numpy() - st["step0_obs_velocity"]).max() < 2e-3
    assert np.abs(obs["pressure"][0].cpu().numpy() - st["step0_obs_pressure"]).max() < 2e-3


@pytest.mark.parametrize("cg_impl", [6, 11, 12])
def test_100_solver_step_horizon_matches_reference(cyl24, golden, cg_impl):
    """north_star: relative L2 <= 1e-3 on u over a 100-step horizon.  Four env.step calls (4 x 25 solver steps) with the
    reference's recorded actions from the reference's reset state, compared with the reference's state, rewards and
    sensors after every env step."""
    spec, cd = cyl24
    from fluidgym_b200.envs.cylinder import CylinderJet2DEnv
    env = CylinderJet2DEnv(n_envs=2, compiled=(spec, cd), cg_impl=cg_impl)
    rs = golden("cyl24_reset.npz")
    st = golden("cyl24_steps.npz")
    env.reset(seed=42)
    env.set_state(rs["u"], rs["presres"], rs["bvel"])
    errs = []
    for k in range(4):
        action = torch.tensor(st["actions"][k], device="cuda").reshape(1, 1).repeat(2, 1)
        obs, reward, term, trunc, info = env.step(action)
        drag_ref = float(st[f"step{k}_info_drag"])
        errs.append((abs(float(reward[0]) - float(st[f"step{k}_reward"])), abs(float(info["drag"][0]) - drag_ref) / abs(drag_ref),
                     float(np.abs(obs["velocity"][0].cpu().numpy() - st[f"step{k}_obs_velocity"]).max())))
    e_u = rel_l2(env.solver.u[0].cpu().numpy(), st["env3_u"])
    e_p = rel_l2(env.solver.p[0].cpu().numpy(), st["env3_p"])
    print("100-step horizon: u", e_u, "p", e_p, "per env step (|d reward|, rel drag, max |d sensor velocity|)", errs)
    # north_star bars: u 1e-3 over 100 steps, rewards 1e-4.  Observed on B200: u 7.4e-6, p 2.6e-5, rewards <= 6.2e-5 absolute
    # (cg_impl 6) / <= 1.2e-4 absolute = 1.2e-5 of |reward| ~ 10 (cg_impl 11, another summation order), drag <= 1.2e-5 relative,
    # sensors <= 1.0e-4 -- the bars below leave one order of magnitude on u; the reward bar is 1e-4 RELATIVE and 2e-4 absolute
    assert e_u < 1e-4
    assert e_p < 5e-4                  # p carries the tolerance ball of its last CG solve
    for k, (d_reward, d_drag, d_obs) in enumerate(errs):
        assert d_reward < 1e-4 * abs(float(st[f"step{k}_reward"])) and d_reward < 2e-4 and d_drag < 1e-4 and d_obs < 1e-3


def test_parallel_fluid_env_api():
    """ParallelFluidEnv(env_id, cuda_ids): one environment per entry, results stacked on the CPU
    (envs/parallel_env.py:162-287 of the reference)."""
    from fluidgym_b200.envs.parallel_env import ParallelFluidEnv
    env = ParallelFluidEnv("CylinderJet2D-easy-v0", cuda_ids=[0, 0, 0])
    obs, info = env.reset(seed=11)
    assert obs["velocity"].shape == (3, 151, 2) and obs["velocity"].device.type == "cpu"
    a = env.sample_action()
    assert a.shape == (3, 1)
    obs, reward, term, trunc, info = env.step(a)
    assert reward.shape == (3,) and set(info) == {"drag", "lift"} and isinstance(term, bool) and isinstance(trunc, bool)
    with pytest.raises(NotImplementedError):
        env.get_state()
    env.close()


def test_batch_entries_are_independent(cyl24):
    """Environment e of a batch evolves exactly as it would alone (no cross-talk through shared workspaces)."""
    spec, cd = cyl24
    from fluidgym_b200.envs.cylinder import CylinderJet2DEnv
    big = CylinderJet2DEnv(n_envs=5, compiled=(spec, cd))
    one = CylinderJet2DEnv(n_envs=1, compiled=(spec, cd))
    big.reset(seed=3)
    one.reset(seed=3)
    gen = torch.Generator(device="cuda").manual_seed(0)
    big.solver.u += 0.02 * torch.randn(big.solver.u.shape, device="cuda", generator=gen)
    one.solver.u.copy_(big.solver.u[3:4])
    act = torch.linspace(-1, 1, 5, device="cuda").reshape(5, 1)
    big.n_sim_steps_override = None
    ob, rb, *_ = big.step(act)
    oo, ro, *_ = one.step(act[3:4])
    assert torch.equal(big.solver.u[3], one.solver.u[0]) and torch.equal(big.solver.p[3], one.solver.p[0])
    assert torch.equal(rb[3], ro[0]) and torch.equal(ob["pressure"][3], oo["pressure"][0])


def test_pushed_halo_cg_is_bit_identical_to_dsmem_gather_cg(cyl24, golden):
    """cg_impl 6 (halo cells pushed into the consumer's shared memory, plain ld.shared gathers) performs exactly the
    arithmetic of cg_impl 3 (gathers through the cluster window): same iterates, bit for bit."""
    spec, cd = cyl24
    fx = golden("cyl24_substep1.npz")
    out = []
    for impl in (3, 6):
        sol = _solver(cd, 2, cg_impl=impl)
        for dst, src in ((sol.u, fx["u_in"]), (sol.p, fx["presres_in"]), (sol.bvel, fx["bvel_in"])):
            dst.copy_(torch.from_numpy(np.ascontiguousarray(src)).cuda().unsqueeze(0).expand_as(dst))
        sol.piso_substep(float(fx["dt"][0]))
        torch.cuda.synchronize()
        out.append((sol.u.clone(), sol.p.clone(), sol.buffer("iters").clone()))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])


def test_full_size_properties_256_envs(cyl24, golden):
    """BASELINE configs[1] size (256 environments x 14 232 cells): properties that do not need the oracle.
    (i) the pressure solve meets its own criterion when re-evaluated with independent kernels (P p = div within the tolerance),
    (ii) the corrector reduces the face-flux divergence (colocated grid without Rhie-Chow: not to zero), (iii) environments that start
    from the same state stay bit-identical, and distinct states stay distinct."""
    spec, cd = cyl24
    fx = golden("cyl24_substep1.npz")
    B = 256
    sol = _solver(cd, B)
    for dst, src in ((sol.u, fx["u_in"]), (sol.p, fx["presres_in"]), (sol.bvel, fx["bvel_in"])):
        dst.copy_(torch.from_numpy(np.ascontiguousarray(src)).cuda().unsqueeze(0).expand_as(dst))
    gen = torch.Generator(device="cuda").manual_seed(0)
    sol.u[128:] += 0.01 * torch.randn(sol.u[128:].shape, device="cuda", generator=gen)      # first half identical, second half distinct
    dt = float(fx["dt"][0])
    sol.piso_substep(dt)
    torch.cuda.synchronize()
    assert torch.equal(sol.u[0], sol.u[127]) and not torch.equal(sol.u[128], sol.u[129])
    res = sol.buffer("resid").cpu().numpy()
    assert (res[:, 2:4] < 1e-5).all() and np.isfinite(res).all()
    # independent residual check of the last pressure system: r = div - P x, evaluated with torch on the ELL arrays
    Poff, Pdiag, div, p = sol.buffer("Poff"), sol.buffer("Pdiag"), sol.buffer("div"), sol.p
    nbr = torch.from_numpy(cd.nbr.astype(np.int64)).cuda()
    x = p + sol.buffer("pmean")[1][:, None]       # the solver's iterate = returned pressure + the mean it removed (P is not singular-exact)
    Pp = Pdiag * x
    for f in range(4):
        inner = nbr[f] >= 0
        Pp = Pp + torch.where(inner, Poff[:, f] * x[:, nbr[f].clamp(min=0)], torch.zeros_like(x))
    r = (div - Pp).pow(2).mean(dim=1).sqrt()
    assert float(r.max()) < 5e-5
    # divergence of the corrected velocity: recompute with the library's own divergence kernel on u
    sol.buffer("hbya").copy_(sol.u)
    native_div = sol.buffer("div").clone()
    from fluidgym_b200 import native
    from fluidgym_b200.solver import _ptr
    native.check(sol.lib.fgb_setup_pressure_rhs(sol.handle, _ptr(sol.u), _ptr(sol.bvel), None, None, _ptr(sol._dt(dt)), 0, None, sol.stream), "div")
    torch.cuda.synchronize()
    d_after = sol.buffer("div").pow(2).mean(dim=1).sqrt()
    d_before = native_div.pow(2).mean(dim=1).sqrt()           # divergence of HbyA (+ deferred term) the last solve removed
    print("divergence rms: HbyA", float(d_before.max()), "corrected velocity", float(d_after.max()))
    assert torch.isfinite(d_after).all() and float(d_after.max()) < float(d_before.max())    # the projection reduces it


def test_wrappers_on_the_cuda_environment():
    import fluidgym_b200 as fg
    from fluidgym_b200 import wrappers
    env = wrappers.FlattenObservation(wrappers.SensorNoise(wrappers.ActionNoise(fg.make("RBC2D-easy-v0", n_envs=3), 0.1, seed=1), 0.01, seed=2))
    obs, _ = env.reset(seed=5)
    assert obs.shape == (3, 8 * 48 + 2 * 8 * 48) and obs.is_cuda
    obs, r, term, trunc, info = env.step(env.sample_action())
    assert obs.shape == (3, 1152) and r.shape == (3,) and torch.isfinite(obs).all()


def test_initial_domain_files_written_by_the_reference_drive_reset(tmp_path):
    """load_initial_domain=True: reset() draws the state of every environment from files in the reference's on-disk format
    (here: the domain the reference itself saved after one env.step, tests/golden/cyl24_domain.*), and a state written by
    save_initial_domain() reads back bit-identically."""
    import shutil
    import fluidgym_b200 as fg
    from conftest import GOLDEN
    root = tmp_path / "initial_domains"
    env = fg.make("CylinderJet2D-easy-v0", n_envs=3, load_initial_domain=True, initial_domains_path=str(root))
    assert env.initial_domain_id == "cylinder_2D_Re100_Res24"
    with pytest.raises(RuntimeError, match="Initial domain not found"):
        env.reset(seed=1)
    for idx in range(10):
        d = root / env.initial_domain_id / str(idx)
        d.mkdir(parents=True)
        for ext in (".json", ".npz"):
            shutil.copy(os.path.join(GOLDEN, "cyl24_domain" + ext), d / ("train" + ext))
    obs, _ = env.reset(seed=1, randomize=False)
    ref = np.load(os.path.join(GOLDEN, "cyl24_domain.npz"))
    u_left = env.solver.u[0, :, :24 * 37].cpu().numpy().reshape(2, 24, 37)
    # reset projects the loaded field again (make_divergence_free, as the reference's _get_simulation does): close, not equal
    print("after reset (re-projected) vs stored", rel_l2(u_left, ref["1"][0]))
    assert rel_l2(u_left, ref["1"][0]) < 0.5 and float(env.solver.u.abs().max()) > 0.5
    env.load_initial_domain(0)
    assert np.array_equal(env.solver.u[2, :, :24 * 37].cpu().numpy().reshape(2, 24, 37), ref["1"][0])
    env.solver.u.mul_(0.5)
    path = env.save_initial_domain(3, mode="val", env_index=1)
    env.val()
    env.solver.u.zero_()
    env._domain_pool.clear()
    env.load_initial_domain(3)
    assert torch.equal(env.solver.u[0], env.solver.u[1]) and float(env.solver.u.abs().max()) > 0.2 and path.endswith("val")


def test_vec_env_adapter_on_the_cuda_environments():
    """integration.VecFluidEnv (SB3 VecEnv protocol) over the batched CUDA environments: single-agent cylinder batch and the
    multi-agent RBC batch (slots = environments x agents)."""
    import fluidgym_b200
    from fluidgym_b200.integration import VecFluidEnv
    v = VecFluidEnv(fluidgym_b200.make("CylinderJet2D-easy-v0", n_envs=3, step_length=0.02))
    obs = v.reset(seed=0)
    assert v.num_envs == 3 and obs["velocity"].shape == (3, 151, 2) and obs["pressure"].dtype == np.float32
    obs, rew, dones, infos = v.step(np.zeros((3, 1), dtype=np.float32))
    assert rew.shape == (3,) and np.isfinite(rew).all() and not dones.any() and "drag" in infos[0]
    m = VecFluidEnv(fluidgym_b200.make("RBC2D-easy-v0", n_envs=2, use_marl=True, step_length=0.1))
    obs = m.reset(seed=0)
    assert m.num_envs == 24 and obs["temperature"].shape == (24, 8, 44)
    obs, rew, dones, infos = m.step(np.zeros((24, 1), dtype=np.float32))
    assert rew.shape == (24,) and np.isfinite(rew).all() and np.isfinite(infos[13]["nusselt"])


@pytest.mark.parametrize("groups", [2, 3])
def test_environment_groups_are_bit_identical(cyl24, golden, groups):
    """fgb_batch_set_groups: contiguous groups of the batch on streams of their own (overlapping solves and assembly of
    independent environments) give bit-identical states, iteration counts and substep counts."""
    spec, cd = cyl24
    fx = golden("cyl24_substep1.npz")
    out = []
    for g in (1, groups):
        sol = _solver(cd, 7, cg_impl=6)
        _load_state(sol, fx, noise=2e-3)
        sol.set_groups(g)
        n = 0
        for _ in range(2):
            n += sol.single_step(0.02, 0.8, char_vel=(1.0, 0.0))
        sol.piso_substep(0.004)
        torch.cuda.synchronize()
        out.append((sol.u.clone(), sol.p.clone(), sol.bvel.clone(), sol.buffer("iters").clone(), n))
    assert out[0][4] == out[1][4]
    for a, b in zip(out[0][:4], out[1][:4]):
        assert torch.equal(a, b)


def test_velocity_gradients_match_reference(cyl24, golden):
    """PISOtorch.ComputeSpatialVelocityGradients of the reference's state after an env.step (5 connected blocks with flipped axes,
    prescribed faces with the one-sided 1.5-cell difference); golden tests/golden/cyl24_velocity_gradients.npz."""
    spec, cd = cyl24
    fx = golden("cyl24_velocity_gradients.npz")
    sol = _solver(cd, 2)
    sol.u.copy_(torch.from_numpy(fx["u"]).cuda().unsqueeze(0).expand_as(sol.u))
    sol.bvel.copy_(torch.from_numpy(fx["bvel"]).cuda().unsqueeze(0).expand_as(sol.bvel))
    g = sol.velocity_gradients()
    torch.cuda.synchronize()
    assert g.shape == (2, 2, 2, cd.N) and torch.equal(g[0], g[1])
    for c in range(2):
        for d in range(2):
            e = rel_l2(g[0, c, d].cpu().numpy(), fx["grad"][c, d])
            print(f"d u_{c} / d x_{d}: rel L2 {e:.2e}")
            assert e < 5e-6
    w = sol.vorticity()[0].cpu().numpy()
    assert rel_l2(w, fx["grad"][1, 0] - fx["grad"][0, 1]) < 1e-5
