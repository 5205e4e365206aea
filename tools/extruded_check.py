"""FIRST GPU RUN of the extruded D = 3 path (fluidgym_b200/extruded3d.py, envs/cylinder3d.py) against the UNMODIFIED reference's
CylinderJet3D-easy run at resolution 8 (tests/golden/cyl3d_substep*.npz, cyl3d_env.npz):
  1. one PISO substep from the reference's traced state (twice, two environments per launch),
  2. environment reset (projection with A = 1) against the reference's reset state,
  3. one env.step (25 solver steps) from the reference's reset state with its action: state, reward, drag / lift, observations,
  4. timing of env.step next to the reference's (tests/golden/cyl3d_meta.json: 1.08 substeps/s).
    python tools/extruded_check.py [--json out.json]
Prints one JSON line per stage and a final {"ok": bool, ...} verdict; exit code 0 iff every bar holds.  The bars are the ones of
the CPU stand-in tests (tests/test_cylinder3d_cpu.py), where the same kernel cell code runs on the host."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain  # noqa: E402
from fluidgym_b200.extruded3d import ExtrudedPISO3D  # noqa: E402

DEV = "cuda:0"
SOLVER = ExtrudedPISO3D          # --standin swaps in the CPU stand-in of the tests (a dry run of THIS TOOL without a GPU, not a product path)


def sync():
    if DEV != "cpu":
        torch.cuda.synchronize()


class _Counters:
    """iteration counters: the Krylov handle's on the GPU, the stand-in's lists in a dry run"""

    def __init__(self, s):
        self.s = s
        self.c0 = self.read()

    def read(self):
        if hasattr(self.s, "cg_iters"):
            return float(sum(self.s.cg_iters)) / max(self.s.B, 1)
        return float(self.s.buffer("iter_total")[0, 0])

    def delta(self):
        return self.read() - self.c0


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(np.asarray(b, np.float64)))


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name))


def check_substeps(cd, out):
    ok = True
    for s in (0, 1):
        fx = golden(f"cyl3d_substep{s}.npz")
        nz, N2 = fx["A"].shape
        sol = SOLVER(cd, nz, float(fx["hz"][0]), n_envs=2, device=DEV)
        sol.u.copy_(torch.from_numpy(fx["u_in"]).reshape(1, 3, -1).to(DEV).expand_as(sol.u))
        sol.p.copy_(torch.from_numpy(fx["p_in"]).reshape(1, -1).to(DEV).expand_as(sol.p))
        sol.bvel.copy_(torch.from_numpy(fx["bvel"]).to(DEV).unsqueeze(0).expand_as(sol.bvel))
        sol.piso_substep(torch.full((2,), float(fx["dt"][0])))
        sync()
        u, p = sol.u[0].cpu().numpy().reshape(3, nz, N2), sol.p[0].cpu().numpy().reshape(nz, N2)
        line = {"stage": "substep", "substep": s, "rel_l2_u": rel(u, fx["u1"]), "rel_l2_p": rel(p, fx["p1"]),
                "ref_bicg_iters": fx["bicg_iters"].tolist(), "ref_cg_iters": fx["cg_iters"].tolist(),
                "envs_equal": bool(torch.equal(sol.u[0], sol.u[1]))}
        line["ok"] = bool(np.isfinite(u).all() and line["rel_l2_u"] < 2e-5 and line["rel_l2_p"] < 5e-4 and line["envs_equal"])
        ok &= line["ok"]
        out.append(line)
        print(json.dumps(line), flush=True)
        del sol
    return ok


def check_env(compiled, out, with_step=True):
    from fluidgym_b200.envs.cylinder3d import CylinderJet3DEnv
    fx = golden("cyl3d_env.npz")
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "cyl3d_meta.json")))
    env = CylinderJet3DEnv(n_envs=2, resolution=8, n_jets=8, compiled=compiled, device=DEV, solver_cls=None if SOLVER is ExtrudedPISO3D else SOLVER)
    t0 = time.perf_counter()
    obs, _ = env.reset(seed=42)
    sync()
    t_reset = time.perf_counter() - t0
    s = env.solver
    line = {"stage": "reset", "seconds": t_reset, "bvel_abs": float(np.abs(s.bvel[0].cpu().numpy() - fx["reset_bvel"]).max()),
            "rel_l2_u": rel(s.u[0].cpu().numpy().reshape(3, 8, -1), fx["reset_u"]), "rel_l2_p": rel(s.p[0].cpu().numpy().reshape(8, -1), fx["reset_p"]),
            "obs_velocity_abs": float(np.abs(obs["velocity"][0].cpu().numpy() - fx["reset_obs_velocity"]).max()),
            "obs_pressure_abs": float(np.abs(obs["pressure"][0].cpu().numpy() - fx["reset_obs_pressure"]).max())}
    line["ok"] = bool(line["bvel_abs"] < 2e-6 and line["rel_l2_u"] < 3e-4 and line["rel_l2_p"] < 3e-4 and line["obs_velocity_abs"] < 1e-3
                      and line["obs_pressure_abs"] < 2e-3)
    ok = line["ok"]
    out.append(line)
    print(json.dumps(line), flush=True)
    if not with_step:
        return ok

    env.set_state(fx["reset_u"], fx["reset_p"], fx["reset_bvel"], last_control=0.0)
    a = torch.from_numpy(fx["actions"][0])[None].expand(2, -1, -1).to(DEV)
    counters = _Counters(s)
    sync()
    t0 = time.perf_counter()
    obs, reward, _, _, info = env.step(a)
    sync()
    t_step = time.perf_counter() - t0
    line = {"stage": "env.step", "seconds": t_step, "substeps": env.last_substeps, "substeps_per_s_x_envs": 2 * env.last_substeps / t_step,
            "reference_substeps_per_s": meta["timing"]["substeps_per_s"],
            "rel_l2_u": rel(s.u[0].cpu().numpy().reshape(3, 8, -1), fx["env0_u"]), "rel_l2_p": rel(s.p[0].cpu().numpy().reshape(8, -1), fx["env0_p"]),
            "reward": float(reward[0]), "ref_reward": float(fx["step0_reward"]), "drag": float(info["drag"][0]), "ref_drag": float(fx["step0_info_drag"]),
            "lift": float(info["lift"][0]), "ref_lift": float(fx["step0_info_lift"]),
            "all_cds_abs": float(np.abs(info["all_cds"][0].cpu().numpy() - fx["step0_info_all_cds"]).max()),
            "obs_velocity_abs": float(np.abs(obs["velocity"][0].cpu().numpy() - fx["step0_obs_velocity"]).max()),
            "obs_pressure_abs": float(np.abs(obs["pressure"][0].cpu().numpy() - fx["step0_obs_pressure"]).max()),
            "envs_equal": bool(torch.allclose(s.u[0], s.u[1], atol=1e-6)),
            # the reference reports the index of the last iteration (one less than the count); CPU stand-in: 1016.6 vs 1014.7 + 1
            "cg_iters_per_solve": counters.delta() / (8 * max(env.last_substeps, 1)),
            "ref_cg_iters_per_solve": meta["mean_iters"]["cg"] + 1.0}
    line["ok"] = bool(env.last_substeps == 25 and line["rel_l2_u"] < 1e-4 and line["rel_l2_p"] < 2e-3 and abs(line["reward"] - line["ref_reward"]) < 1e-3
                      and abs(line["drag"] - line["ref_drag"]) < 1e-3 and abs(line["lift"] - line["ref_lift"]) < 1e-4 and line["all_cds_abs"] < 1e-3
                      and line["obs_velocity_abs"] < 2e-4 and line["obs_pressure_abs"] < 2e-3
                      and abs(line["cg_iters_per_solve"] - line["ref_cg_iters_per_solve"]) < 0.03 * line["ref_cg_iters_per_solve"])
    ok &= line["ok"]
    out.append(line)
    print(json.dumps(line), flush=True)
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--skip-env-step", action="store_true", help="substep and reset stages only")
    ap.add_argument("--standin", action="store_true", help="dry run of this tool on the CPU stand-in solver of the tests (no GPU)")
    args = ap.parse_args()
    if args.standin:
        global DEV, SOLVER
        sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_harness"))
        from extruded_standin import HostExtrudedPISO3D
        DEV, SOLVER = "cpu", HostExtrudedPISO3D
    spec = make_cylinder_domain(8)
    cd = spec.prepare()
    out, ok = [], True
    try:
        ok &= check_substeps(cd, out)
        ok &= check_env((spec, cd), out, with_step=not args.skip_env_step)
    except Exception as e:                                              # report, do not hide: this is a bring-up tool
        out.append({"stage": "exception", "error": f"{type(e).__name__}: {e}"})
        print(json.dumps(out[-1]), flush=True)
        ok = False
    verdict = {"ok": bool(ok), "stages": out}
    print(json.dumps({"ok": bool(ok)}))
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        json.dump(verdict, open(args.json, "w"), indent=1)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
