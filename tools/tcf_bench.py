"""Turbulent channel flow (BASELINE config 5, single GPU part) timing: env.step throughput of the D = 3 path.
    python tools/tcf_bench.py --ids TCFSmall3D-both-easy-v0 TCFLarge3D-both-easy-v0 --steps 3
State = Reichardt profile + 5 % Gaussian noise (the same perturbation oracle/ref_harness.py --perturb 0.05 applies to the reference).
"""
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
    ap.add_argument("--ids", nargs="+", default=["TCFSmall3D-both-easy-v0", "TCFLarge3D-both-easy-v0"])
    ap.add_argument("--envs", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rows = []
    for env_id in a.ids:
        env = fg.make(env_id, n_envs=a.envs)
        env.reset(seed=42)
        g = torch.Generator(device="cuda").manual_seed(42)
        env.solver.u += 0.05 * torch.randn(env.solver.u.shape, device="cuda", generator=g)
        act = torch.zeros_like(env._zero_action)
        env.step(act)
        it0 = env.solver.buffer("iter_total").clone()
        l0 = env.lib.fgb_ortho3_launch_count(env.solver.handle)
        torch.cuda.synchronize()
        t0 = time.time()
        nsub = 0
        for i in range(a.steps):
            env.step(torch.full_like(env._zero_action, 0.3))
            nsub += env.last_substeps
        torch.cuda.synchronize()
        dt = time.time() - t0
        it = (env.solver.buffer("iter_total") - it0)[0].tolist()
        row = dict(env=env_id, cells=env.dom.N, n_envs=a.envs, env_steps_per_s=a.envs * a.steps / dt, substeps_per_s=a.envs * nsub / dt,
                   ms_per_substep=1e3 * dt / max(nsub, 1), substeps_per_env_step=nsub / a.steps, cg_iters_per_solve=it[0] / max(2 * nsub, 1),
                   bicg_iters_per_rhs=it[1] / max(3 * nsub, 1), launches_per_substep=(env.lib.fgb_ortho3_launch_count(env.solver.handle) - l0) / max(nsub, 1),
                   coop_blocks=None, seconds=dt, gpu=torch.cuda.get_device_name(0))
        print(json.dumps(row), flush=True)
        rows.append(row)
        del env
        torch.cuda.empty_cache()
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
