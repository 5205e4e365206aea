#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- golden-vector generator; runs the UNMODIFIED reference on a GPU box.

Drives the reference package installed under ``baseline/_ref`` (see ``oracle/build_ref.sh``)
through its own public API (``fluidgym.make`` -> ``reset`` -> ``step``) and records

* the geometry the reference built (vertex coordinates, cell/boundary transforms),
* the state after ``reset`` and after selected sim steps (u, p per block, boundary velocities),
* every native op of the first PISO substeps (``SetupAdvectionMatrix`` ... ``CorrectVelocity``,
  ``SolveLinear`` with iteration counts / residuals) by wrapping the pybind11 module functions
  that ``PISOtorch_simulation.py`` calls (``SIM.py:1431-2002``),
* observations / reward / drag / lift of ``env.step`` and the wall-clock of the reference
  CUDA path (B=1) for BASELINE.md section 3a.

Output: ``<out>/<tag>_*.npz`` + ``<out>/<tag>_meta.json``.  The small files are committed under
``tests/golden/`` and pin ``oracle/`` (parity is otherwise unpinned: SURVEY.md section 8c).
Needs a GPU; never reads /root/reference (the box does not have it).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402


def t2n(t):
    return t.detach().cpu().numpy().copy()


def dump_geometry(env, out, tag):
    dom = env._domain
    geo = {}
    blocks = dom.getBlocks()
    for bi, blk in enumerate(blocks):
        geo[f"b{bi}_vertex"] = t2n(blk.vertexCoordinates)
        geo[f"b{bi}_transform"] = t2n(blk.transform)
        for f in range(2 * dom.getSpatialDims()):
            bnd = blk.getBoundary(f)
            if type(bnd).__name__ == "FixedBoundary":
                geo[f"b{bi}_f{f}_transform"] = t2n(bnd.transform)
                geo[f"b{bi}_f{f}_velocity"] = t2n(bnd.velocity)
    np.savez_compressed(os.path.join(out, f"{tag}_geometry.npz"), **geo)


def snapshot_state(env):
    dom = env._domain
    st = {}
    for bi, blk in enumerate(dom.getBlocks()):
        st[f"b{bi}_u"] = t2n(blk.velocity)
        st[f"b{bi}_p"] = t2n(blk.pressure)
        if dom.hasPassiveScalar():
            st[f"b{bi}_s"] = t2n(blk.passiveScalar)
        for f in range(2 * dom.getSpatialDims()):
            bnd = blk.getBoundary(f)
            if type(bnd).__name__ == "FixedBoundary":
                st[f"b{bi}_f{f}_velocity"] = t2n(bnd.velocity)
                if dom.hasPassiveScalar() and bnd.passiveScalar is not None:
                    st[f"b{bi}_f{f}_scalar"] = t2n(bnd.passiveScalar)
    st["pressureResult"] = t2n(dom.pressureResult)
    st["velocityResult"] = t2n(dom.velocityResult)
    return st


class OpTracer:
    """Wraps the native ops the python driver calls and records their outputs."""

    OPS = ["SetupAdvectionMatrix", "SetupAdvectionVelocity", "SetupPressureMatrix", "SetupPressureRHS",
           "SetupPressureRHSdiv", "CorrectVelocity", "SolveLinear", "CopyVelocityResultToBlocks",
           "SetupAdvectionScalar", "SetupPressureCorrection"]

    def __init__(self, PISOtorch, max_substeps):
        self.P = PISOtorch
        self.max_substeps = max_substeps
        self.records = {}
        self.meta = []
        self.substep = -1
        self.n_substeps_total = 0
        self.sim_step_states = []
        self.keep_states_every = None
        self.env = None
        self.solver_log = []
        self._orig = {}

    def install(self):
        for name in self.OPS:
            if not hasattr(self.P, name):
                continue
            self._orig[name] = getattr(self.P, name)
            setattr(self.P, name, self._make(name))

    def uninstall(self):
        for name, fn in self._orig.items():
            setattr(self.P, name, fn)

    lean = False

    def _rec(self, key, t):
        if self.lean and key.split("_")[0] in ("C", "P", "Cs"):
            return
        self.records[f"s{self.substep}_{key}"] = t2n(t)

    def _make(self, name):
        orig = self._orig[name]

        def wrapper(*a, **k):
            is_scalar_mat = name == "SetupAdvectionMatrix" and bool(k.get("forPassiveScalar", False) or (len(a) > 3 and a[3]))
            starts = name == "SetupAdvectionMatrix" and (is_scalar_mat or not a[0].hasPassiveScalar())
            if starts:
                self.substep += 1
                self.n_substeps_total += 1
                self.count = {}
                if self.substep < self.max_substeps:
                    dom = a[0]
                    self.records[f"s{self.substep}_dt"] = t2n(a[1])
                    for bi, blk in enumerate(dom.getBlocks()):
                        self._rec(f"in_b{bi}_u", blk.velocity)
                        self._rec(f"in_b{bi}_p", blk.pressure)
                        if blk.hasViscosity():                       # per-cell viscosity set by a "PRE" prep function (SGS model)
                            self._rec(f"in_b{bi}_viscosity", blk.viscosity)
                        if dom.hasPassiveScalar():
                            self._rec(f"in_b{bi}_s", blk.passiveScalar)
                        for f in range(2 * dom.getSpatialDims()):
                            bnd = blk.getBoundary(f)
                            if type(bnd).__name__ == "FixedBoundary":
                                self._rec(f"in_b{bi}_f{f}_velocity", bnd.velocity)
                                if dom.hasPassiveScalar() and bnd.passiveScalar is not None:
                                    self._rec(f"in_b{bi}_f{f}_scalar", bnd.passiveScalar)
                    self._rec("in_velocityResult", dom.velocityResult)
                    self._rec("in_pressureResult", dom.pressureResult)
            res = orig(*a, **k)
            rec = self.substep < self.max_substeps and self.substep >= 0
            c = self.count.get(name, 0) if hasattr(self, "count") else 0
            if hasattr(self, "count"):
                self.count[name] = c + 1
            if name == "SolveLinear":
                infos = [(float(i.finalResidual), int(i.usedIterations), bool(i.converged)) for i in res]
                use_bicg = bool(a[6]) if len(a) > 6 else bool(k.get("useBiCG", False))
                self.solver_log.append((self.substep, "bicg" if use_bicg else "cg", infos))
                if rec:
                    self._rec(f"solve{c}_x", a[2])
                    self._rec(f"solve{c}_rhs", a[1])
                    self.meta.append({"substep": self.substep, "solve": c, "bicg": use_bicg, "infos": infos})
            elif rec:
                dom = a[0]
                if name == "SetupAdvectionMatrix" and is_scalar_mat:
                    self._rec("Cs_value", dom.C.value)
                    self._rec("Cs_index", dom.C.index)
                    self._rec("Cs_row", dom.C.row)
                elif name == "SetupAdvectionScalar":
                    self._rec(f"scalarRHS{c}", dom.scalarRHS)
                elif name == "SetupPressureCorrection":
                    self._rec(f"P_value{c}", dom.P.value)
                    if c == 0:
                        self._rec("P_index", dom.P.index)
                        self._rec("P_row", dom.P.row)
                    self._rec(f"pressureRHS{c}", dom.pressureRHS)
                    self._rec(f"pressureRHSdiv{c}", dom.pressureRHSdiv)
                elif name == "SetupAdvectionMatrix":
                    self._rec("C_value", dom.C.value)
                    self._rec("C_index", dom.C.index)
                    self._rec("C_row", dom.C.row)
                    self._rec("A", dom.A)
                elif name == "SetupAdvectionVelocity":
                    self._rec(f"velocityRHS{c}", dom.velocityRHS)
                    for bi, blk in enumerate(dom.getBlocks()):
                        if blk.velocitySource is not None:
                            self._rec(f"b{bi}_velocitySource", blk.velocitySource)
                elif name == "SetupPressureMatrix":
                    self._rec(f"P_value{c}", dom.P.value)
                    if c == 0:
                        self._rec("P_index", dom.P.index)
                        self._rec("P_row", dom.P.row)
                elif name in ("SetupPressureRHS", "SetupPressureRHSdiv"):
                    cc = self.count.get("SetupPressureRHS", 0) + self.count.get("SetupPressureRHSdiv", 0) - 1
                    self._rec(f"pressureRHS{cc}", dom.pressureRHS)
                    self._rec(f"pressureRHSdiv{cc}", dom.pressureRHSdiv)
                elif name == "CorrectVelocity":
                    self._rec(f"velocityResult{c}", dom.velocityResult)
                    self._rec(f"pressureResult{c}", dom.pressureResult)
            return res

        return wrapper


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/golden")
    ap.add_argument("--env", default="CylinderJet2D-easy-v0")
    ap.add_argument("--tag", default=None)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--trace-substeps", type=int, default=2)
    ap.add_argument("--env-steps", type=int, default=4)
    ap.add_argument("--time-steps", type=int, default=8)
    ap.add_argument("--action", type=float, default=0.5)
    ap.add_argument("--kw", default="{}", help="JSON dict of extra fluidgym.make keyword arguments")
    ap.add_argument("--perturb", type=float, default=0.0, help="std of Gaussian noise added to the block velocities after reset")
    ap.add_argument("--lean", action="store_true", help="do not record the CSR matrices (large 3-D grids)")
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"], help="fluidgym.make(dtype=...)")
    ap.add_argument("--pressure-tol", type=float, default=None, help="override Simulation.pressure_tol after reset (tight-tolerance goldens)")
    ap.add_argument("--advection-tol", type=float, default=None, help="override Simulation.advection_tol after reset")
    ap.add_argument("--res-z", type=int, default=None,
                    help="Airfoil3D: spanwise resolution (class attribute AirfoilEnvBase._res_z, 96 in the reference) -- a smaller value "
                         "makes the golden run affordable; the reference's code is untouched")
    ap.add_argument("--gradients", action="store_true", help="also record ComputeSpatialVelocityGradients of the state after the env steps")
    ap.add_argument("--save-domain-only", action="store_true",
                    help="reset, advance --env-steps steps, write the domain with the reference's own save_domain() and exit")
    args = ap.parse_args()
    tag = args.tag or args.env.replace("-", "_")
    os.makedirs(args.out, exist_ok=True)

    ref_shims.install()
    import torch
    import fluidgym
    from fluidgym.simulation.extensions import PISOtorch

    assert torch.cuda.is_available(), "the reference needs a GPU"
    meta = {"env": args.env, "seed": args.seed, "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0)}

    kw = json.loads(args.kw)
    if args.res_z is not None:
        from fluidgym.envs.airfoil.airfoil_env_base import AirfoilEnvBase
        AirfoilEnvBase._res_z = int(args.res_z)
        meta["res_z"] = int(args.res_z)
    if args.dtype == "float64":
        kw["dtype"] = torch.float64
    env = fluidgym.make(args.env, load_initial_domain=False, load_domain_statistics=False,
                        randomize_initial_state=False, **kw)
    meta["dtype"] = args.dtype
    t0 = time.time()
    obs0, _ = env.reset(seed=args.seed)
    torch.cuda.synchronize()
    meta["reset_seconds"] = time.time() - t0
    for name, val in (("pressure_tol", args.pressure_tol), ("advection_tol", args.advection_tol)):
        if val is not None:
            assert hasattr(env._sim, name), name
            setattr(env._sim, name, val)
            meta[name] = val
    if args.perturb > 0:
        g = torch.Generator(device="cuda").manual_seed(args.seed)
        for blk in env._domain.getBlocks():
            u = blk.velocity
            blk.setVelocity((u + args.perturb * torch.randn(u.shape, device=u.device, generator=g)).contiguous())
        env._domain.UpdateDomainData()
        meta["perturb"] = args.perturb
    if args.save_domain_only:
        from fluidgym.simulation.pict.util.domain_io import save_domain
        for i in range(args.env_steps):
            env.step(torch.full_like(env._zero_action, args.action))
        save_domain(env._domain, os.path.join(args.out, f"{tag}_domain"))
        np.savez_compressed(os.path.join(args.out, f"{tag}_domain_state.npz"), **snapshot_state(env))
        print("saved", os.path.join(args.out, f"{tag}_domain"))
        return
    dump_geometry(env, args.out, tag)
    st = snapshot_state(env)
    st.update({f"obs_{k}": t2n(v) for k, v in obs0.items()})
    np.savez_compressed(os.path.join(args.out, f"{tag}_state_reset.npz"), **st)
    meta["n_sim_steps"] = int(env._n_sim_steps)
    meta["dt"] = float(env._dt)
    visc = getattr(env, "_viscosity", None)
    if visc is None:
        visc = getattr(env, "_kinematic_viscosity", None)
    meta["viscosity"] = float(visc.cpu().item()) if visc is not None else None
    for k2 in ("_thermal_diffusivity",):
        if hasattr(env, k2):
            meta[k2.strip("_")] = float(getattr(env, k2).cpu().item())

    tracer = OpTracer(PISOtorch, args.trace_substeps)
    tracer.lean = args.lean
    tracer.install()

    # record the block state after every sim step of the first env.step
    sim = env._sim
    orig_single = sim.single_step
    per_sim = []

    def single_step_logged(*a, **k):
        r = orig_single(*a, **k)
        if len(per_sim) < 200:
            s = {}
            for bi, blk in enumerate(env._domain.getBlocks()):
                s[f"b{bi}_u"] = t2n(blk.velocity)
                s[f"b{bi}_p"] = t2n(blk.pressure)
                if env._domain.hasPassiveScalar():
                    s[f"b{bi}_s"] = t2n(blk.passiveScalar)
            s["substeps_so_far"] = np.array(tracer.n_substeps_total)
            per_sim.append(s)
        return r

    sim.single_step = single_step_logged

    step_out = {}
    actions = []
    for i in range(args.env_steps):
        act = torch.full_like(env._zero_action, args.action * float(np.cos(0.7 * i)))
        if act.numel() > 1:
            ramp = torch.linspace(-1.0, 1.0, act.numel(), device=act.device).reshape(act.shape)
            act = act * torch.sin(3.0 * ramp + 0.5 * i)
        actions.append(t2n(act))
        obs, reward, term, trunc, info = env.step(act)
        step_out[f"step{i}_reward"] = t2n(reward)
        for k, v in obs.items():
            step_out[f"step{i}_obs_{k}"] = t2n(v)
        for k, v in info.items():
            step_out[f"step{i}_info_{k}"] = t2n(v)
        sst = snapshot_state(env)
        np.savez_compressed(os.path.join(args.out, f"{tag}_state_step{i}.npz"), **sst)
    step_out["actions"] = np.stack(actions)
    if args.gradients:                                   # PISOtorch.ComputeSpatialVelocityGradients of the final state
        env._domain.UpdateDomainData()
        grads = PISOtorch.ComputeSpatialVelocityGradients(env._domain)
        np.savez_compressed(os.path.join(args.out, f"{tag}_gradients.npz"),
                            **{f"b{bi}_d{d}": t2n(g) for bi, gb in enumerate(grads) for d, g in enumerate(gb)})
    np.savez_compressed(os.path.join(args.out, f"{tag}_steps.npz"), **step_out)
    np.savez_compressed(os.path.join(args.out, f"{tag}_trace.npz"), **tracer.records)
    for j in (0, 1, len(per_sim) - 1):
        if 0 <= j < len(per_sim):
            np.savez_compressed(os.path.join(args.out, f"{tag}_simstep{j}.npz"), **per_sim[j])
    meta["trace_meta"] = tracer.meta
    meta["substeps_in_env_steps"] = tracer.n_substeps_total
    its = {"cg": [], "bicg": []}
    for _, kind, infos in tracer.solver_log:
        for (res, it, conv) in infos:
            its[kind].append(it)
    meta["mean_iters"] = {k: (float(np.mean(v)) if v else None) for k, v in its.items()}
    meta["max_iters"] = {k: (int(np.max(v)) if v else None) for k, v in its.items()}
    meta["n_solves"] = {k: len(v) for k, v in its.items()}

    if args.time_steps <= 0:
        with open(os.path.join(args.out, f"{tag}_meta.json"), "w") as f:
            json.dump(meta, f, indent=1)
        print(json.dumps({k: meta[k] for k in ("mean_iters", "max_iters", "n_solves", "substeps_in_env_steps")}))
        return
    # timing of the reference CUDA path (B=1): env.step wall clock
    tracer.uninstall()
    sim.single_step = orig_single
    n_sub0 = 0
    cnt = {"n": 0}
    orig_mat = PISOtorch.SetupAdvectionMatrix

    def count_mat(*a, **k):
        cnt["n"] += 1
        return orig_mat(*a, **k)

    PISOtorch.SetupAdvectionMatrix = count_mat
    act = torch.zeros_like(env._zero_action)
    for _ in range(2):
        env.step(act)
    torch.cuda.synchronize()
    cnt["n"] = 0
    t0 = time.time()
    for i in range(args.time_steps):
        env.step(torch.full_like(env._zero_action, 0.3 * float(np.sin(0.3 * i))))
    torch.cuda.synchronize()
    el = time.time() - t0
    meta["timing"] = {"env_steps": args.time_steps, "seconds": el, "substeps": cnt["n"],
                      "env_steps_per_s": args.time_steps / el, "substeps_per_s": cnt["n"] / el}
    with open(os.path.join(args.out, f"{tag}_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps({k: meta[k] for k in ("timing", "mean_iters", "max_iters", "n_solves", "substeps_in_env_steps")}))


if __name__ == "__main__":
    main()
