"""CylinderJet3D timing: env.step throughput of the extruded D = 3 path on one GPU, next to the
unmodified reference's figure at resolution 8 (tests/golden/cyl3d_meta.json: 1.08 substeps/s on a B200).
    python tools/cyl3d_bench.py --resolutions 8 24 --steps 2 [--envs 1] [--out file.json]
State = reset (projection of the inflow field) + `--settle` uncontrolled env steps; actions 0.3 on every jet.  Device-timed
between synchronisations; CG / BiCGStab iteration counts from the solver's own counters."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluidgym_b200 as fg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--resolutions", nargs="+", type=int, default=[8, 24])
    ap.add_argument("--envs", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--settle", type=int, default=1)
    ap.add_argument("--out", default=None)
    ap.add_argument("--airfoil3d", action="store_true", help="also time Airfoil3D-easy-v0 (4.5 M cells; first run: finite-ness + timing only)")
    a = ap.parse_args()
    ref = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cyl3d_meta.json")))
    rows = []
    cases = [("CylinderJet3D-easy-v0", dict(resolution=res), res) for res in a.resolutions]
    if a.airfoil3d:
        cases.append(("Airfoil3D-easy-v0", {}, 96))
    for env_id, kw, res in cases:
        env = fg.make(env_id, n_envs=a.envs, **kw)
        t0 = time.time()
        env.reset(seed=42)
        torch.cuda.synchronize()
        t_reset = time.time() - t0
        for _ in range(a.settle):
            env.step(torch.zeros_like(env._zero_action))
        s = env.solver
        it0, l0 = s.buffer("iter_total").clone(), s.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        nsub = 0
        for _ in range(a.steps):
            env.step(torch.full_like(env._zero_action, 0.3))
            nsub += env.last_substeps
        e1.record()
        torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) * 1e-3
        it = (s.buffer("iter_total") - it0)[0].tolist()
        row = dict(env=env_id, finite=bool(torch.isfinite(s.u).all() and torch.isfinite(s.p).all()), resolution=res, cells=s.N, n_envs=a.envs, reset_seconds=t_reset,
                   env_steps_per_s=a.envs * a.steps / dt, substeps_per_s=a.envs * nsub / dt, ms_per_substep=1e3 * dt / max(nsub, 1),
                   cg_iters_per_solve=it[0] / max(8 * nsub, 1), bicg_iters_per_rhs=it[1] / max(3 * nsub, 1),
                   us_per_cg_iteration_upper_bound=1e6 * dt / max(it[0], 1),
                   launches_per_substep=(s.launch_count() - l0) / max(nsub, 1), seconds=dt, gpu=torch.cuda.get_device_name(0),
                   reference_substeps_per_s_res8=ref["timing"]["substeps_per_s"], reference_cg_iters_per_solve_res8=ref["mean_iters"]["cg"])
        print(json.dumps(row), flush=True)
        rows.append(row)
        del env, s
        torch.cuda.empty_cache()
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
