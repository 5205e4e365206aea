"""Airfoil2D-medium (BASELINE config 4) timing: forward env.step throughput for a few batch sizes and one
differentiable rollout (forward + backward through `--diff-steps` env steps) on one GPU.
    python tools/airfoil_bench.py --envs 1 8 --steps 2 --diff-steps 2
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
    ap.add_argument("--envs", type=int, nargs="+", default=[1, 8])
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--diff-steps", type=int, default=2)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = {"env": "Airfoil2D-medium-v0", "gpu": torch.cuda.get_device_name(0), "forward": [], "differentiable": None}
    compiled = None
    for B in a.envs:
        env = fg.make("Airfoil2D-medium-v0", n_envs=B, compiled=compiled)
        compiled = (env.spec, env.cd)
        env.reset(seed=42)
        act = torch.linspace(-1, 1, 3, device="cuda").repeat(B, 1) * 0.5
        env.step(act)                                   # warm-up
        it0 = env.solver.buffer("iter_total").clone()
        torch.cuda.synchronize()
        t0 = time.time()
        nsub = 0
        for _ in range(a.steps):
            env.step(act)
            nsub += env.last_substeps
        torch.cuda.synchronize()
        dt = time.time() - t0
        its = (env.solver.buffer("iter_total") - it0)[:, 0].float().mean().item()
        row = dict(n_envs=B, env_steps_per_s=B * a.steps / dt, substeps_per_s=B * nsub / dt, substeps_per_env_step=nsub / a.steps,
                   cg_iterations_per_substep=its / max(nsub, 1), seconds=dt)
        print(json.dumps(row), flush=True)
        res["forward"].append(row)
        del env
    if a.diff_steps > 0:
        env = fg.make("Airfoil2D-medium-v0", n_envs=1, compiled=compiled, differentiable=True)
        env.reset(seed=42)
        actions = [(torch.linspace(-1, 1, 3, device="cuda").reshape(1, 3) * 0.5).requires_grad_(True) for _ in range(a.diff_steps)]
        torch.cuda.synchronize()
        t0 = time.time()
        total, nsub = 0.0, 0
        for act in actions:
            obs, r, *_ = env.step(act)
            total = total + r.sum()
            nsub += env.last_substeps
        torch.cuda.synchronize()
        t1 = time.time()
        total.backward()
        torch.cuda.synchronize()
        t2 = time.time()
        res["differentiable"] = dict(env_steps=a.diff_steps, substeps=nsub, forward_s=t1 - t0, backward_s=t2 - t1,
                                     grad_first_action=actions[0].grad.flatten().tolist(), grad_last_action=actions[-1].grad.flatten().tolist(),
                                     peak_mem_GB=torch.cuda.max_memory_allocated() / 2 ** 30)
        print(json.dumps(res["differentiable"]), flush=True)
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
