#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- gradient goldens; runs the UNMODIFIED reference (baseline/_ref) with
``differentiable=True`` on a GPU box.

What the reference itself documents as its gradient interface (examples/interfaces/gradient_based_methods.py:
``reward.backward()`` -> ``action.grad``; examples/advanced/compute_state_vjp.py + envs/util/diff_tools.py:8-60:
``mark_state_differentiable`` -> ``torch.autograd.grad(outputs, inputs, cotangent)``) is recorded for one
``env.step`` from the reset state (optionally after ``--develop`` undifferentiated env steps):

* the state before the step (block velocities / pressures / boundary velocities, passive scalar),
* the action, the reward, d reward / d action, d reward / d (block velocity [, passive scalar]) of the incoming state,
* the vector-Jacobian product of the outgoing flat state with a fixed deterministic cotangent w.r.t. the incoming
  state and the action (the reference's state_vjp example with a non-trivial cotangent),
* the state after the step and the Krylov iteration counts of the forward pass.

Output: ``<out>/<tag>_grad.npz`` + ``<out>/<tag>_grad_meta.json``; reduced by tests/golden/extract_grad_fixtures.py.
Never reads /root/reference.
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
from ref_harness import snapshot_state, t2n  # noqa: E402


def cotangent_like(t, phase):
    """Deterministic, sign-changing cotangent: sin(0.37 i + phase) over the flat index."""
    import torch
    i = torch.arange(t.numel(), device=t.device, dtype=torch.float64)
    return torch.sin(0.37 * i + phase).to(t.dtype).reshape(t.shape)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/golden_grad")
    ap.add_argument("--env", default="CylinderJet2D-easy-v0")
    ap.add_argument("--tag", default=None)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--develop", type=int, default=0, help="env steps (zero action, detached) before the differentiated step")
    ap.add_argument("--action", type=float, default=0.5)
    ap.add_argument("--kw", default="{}")
    ap.add_argument("--pressure-tol", type=float, default=None)
    ap.add_argument("--advection-tol", type=float, default=None)
    ap.add_argument("--res-z", type=int, default=None, help="Airfoil3D: spanwise resolution (class attribute AirfoilEnvBase._res_z), as ref_harness.py")
    ap.add_argument("--perturb", type=float, default=0.0, help="std of Gaussian noise added to the block velocities after reset (as ref_harness.py)")
    args = ap.parse_args()
    tag = args.tag or args.env.replace("-", "_")
    os.makedirs(args.out, exist_ok=True)

    ref_shims.install()
    import torch
    import fluidgym
    from fluidgym.simulation.extensions import PISOtorch

    assert torch.cuda.is_available(), "the reference needs a GPU"
    meta = {"env": args.env, "seed": args.seed, "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
            "develop": args.develop, "action": args.action}
    if args.res_z is not None:
        from fluidgym.envs.airfoil.airfoil_env_base import AirfoilEnvBase
        AirfoilEnvBase._res_z = int(args.res_z)
        meta["res_z"] = int(args.res_z)
    env = fluidgym.make(args.env, differentiable=True, load_initial_domain=False, load_domain_statistics=False,
                        randomize_initial_state=False, **json.loads(args.kw))
    env.reset(seed=args.seed)
    for name, val in (("pressure_tol", args.pressure_tol), ("advection_tol", args.advection_tol)):
        if val is not None:
            setattr(env._sim, name, val)
            meta[name] = val
    if args.perturb > 0:
        gen = torch.Generator(device="cuda").manual_seed(args.seed)
        for blk in env._domain.getBlocks():
            u = blk.velocity
            blk.setVelocity((u + args.perturb * torch.randn(u.shape, device=u.device, generator=gen)).contiguous())
        env._domain.UpdateDomainData()
        meta["perturb"] = args.perturb
    for _ in range(args.develop):
        with torch.no_grad():
            env.step(torch.zeros_like(env._zero_action))
        env.detach()
    out = {f"pre_{k}": v for k, v in snapshot_state(env).items()}

    # Krylov iteration log of the differentiated forward pass
    log = []
    orig = PISOtorch.SolveLinear

    def solve_logged(*a, **k):
        res = orig(*a, **k)
        use_bicg = bool(a[6]) if len(a) > 6 else bool(k.get("useBiCG", False))
        log.append(("bicg" if use_bicg else "cg", [int(i.usedIterations) for i in res]))
        return res

    PISOtorch.SolveLinear = solve_logged

    dom = env._domain
    has_s = dom.hasPassiveScalar()
    inputs, names = [], []
    for bi, blk in enumerate(dom.getBlocks()):      # = envs/util/diff_tools.py::mark_state_differentiable
        inputs.append(blk.velocity.requires_grad_(True)); names.append(f"b{bi}_u")
        if has_s:
            inputs.append(blk.passiveScalar.requires_grad_(True)); names.append(f"b{bi}_s")
    act = torch.full_like(env._zero_action, args.action)
    if act.numel() > 1:
        ramp = torch.linspace(-1.0, 1.0, act.numel(), device=act.device).reshape(act.shape)
        act = act * torch.sin(3.0 * ramp + 0.5)
    act = act.clone().requires_grad_(True)
    t0 = time.time()
    obs, reward, term, trunc, info = env.step(act)
    torch.cuda.synchronize()
    meta["forward_seconds"] = time.time() - t0
    n_forward = len(log)
    PISOtorch.SolveLinear = orig
    out["action"] = t2n(act)
    out["reward"] = t2n(reward)
    for k, v in info.items():
        if hasattr(v, "detach"):
            out[f"info_{k}"] = t2n(v)
    for k, v in obs.items():
        out[f"obs_{k}"] = t2n(v)
    out.update({f"post_{k}": v for k, v in snapshot_state(env).items()})

    # (1) d reward / d (action, incoming state)
    t0 = time.time()
    g = torch.autograd.grad(reward.sum(), [act] + inputs, retain_graph=True, allow_unused=True)
    torch.cuda.synchronize()
    meta["backward_seconds"] = time.time() - t0
    out["dreward_daction"] = t2n(g[0]) if g[0] is not None else np.zeros(act.shape, np.float32)
    for n, gi in zip(names, g[1:]):
        out[f"dreward_d{n}"] = t2n(gi) if gi is not None else np.zeros(0, np.float32)
    # (2) vjp of the outgoing state with a fixed cotangent (diff_tools.get_flat_state order: velocity [, scalar] per block)
    outs, cots = [], []
    for bi, blk in enumerate(dom.getBlocks()):
        outs.append(blk.velocity); cots.append(cotangent_like(blk.velocity, 0.1 * bi))
        if has_s:
            outs.append(blk.passiveScalar); cots.append(cotangent_like(blk.passiveScalar, 0.3 + 0.1 * bi))
    for bi, c in enumerate(cots):
        out[f"cotangent{bi}"] = t2n(c)
    g2 = torch.autograd.grad(outs, [act] + inputs, grad_outputs=cots, allow_unused=True)
    out["vjp_daction"] = t2n(g2[0]) if g2[0] is not None else np.zeros(act.shape, np.float32)
    for n, gi in zip(names, g2[1:]):
        out[f"vjp_d{n}"] = t2n(gi) if gi is not None else np.zeros(0, np.float32)
    env.detach()
    its = {"cg": [], "bicg": []}
    for kind, it in log[:n_forward]:
        its[kind].extend(it)
    meta["forward_iters"] = {k: {"n": len(v), "mean": float(np.mean(v)) if v else None, "max": int(np.max(v)) if v else None}
                             for k, v in its.items()}
    meta["n_sim_steps"] = int(env._n_sim_steps)
    meta["dt"] = float(env._dt)
    np.savez_compressed(os.path.join(args.out, f"{tag}_grad.npz"), **out)
    with open(os.path.join(args.out, f"{tag}_grad_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps({"reward": out["reward"].tolist(), "dreward_daction": out["dreward_daction"].tolist(),
                      "vjp_daction": out["vjp_daction"].tolist(), **meta}))


if __name__ == "__main__":
    main()
