"""Strong scaling of ONE large turbulent-channel domain slab-decomposed over the GPUs of a node (BASELINE config 5).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node=N --master-addr 127.0.0.1 tools/tcf_slab_bench.py [--large] [--steps K]
Every rank owns nz/N z-planes; halo planes / reduction partials travel over NVLink inside the persistent Krylov kernels.
Timing: CUDA events on every rank around K solver steps (adaptive CFL substeps, dynamic forcing), max over ranks.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fluidgym_b200.box3d import Box3DDomain, SlabPISO3D  # noqa: E402
from fluidgym_b200.envs.tcf import LARGE_TCF_3D_DEFAULT_CONFIG, SMALL_TCF_3D_DEFAULT_CONFIG, re_wall_to_cl  # noqa: E402
from fluidgym_b200.grids import channel_vertex_grid  # noqa: E402


def run(large=True, nz_mult=1, steps=20, warmup=5, plain=False, rank=0, world=1, local=0):
    """One strong-scaling point (needs an initialised NCCL process group when world > 1); returns the record on every rank."""
    a = argparse.Namespace(large=large, nz_mult=nz_mult, steps=steps, warmup=warmup, plain=plain)
    cfg = LARGE_TCF_3D_DEFAULT_CONFIG if a.large else SMALL_TCF_3D_DEFAULT_CONFIG
    x = z = cfg["resolution_x_z"]
    z *= a.nz_mult
    re_c = re_wall_to_cl(cfg["reynolds_number_wall"])
    visc = float(np.float32(1.0 / re_c))
    u_wall = cfg["reynolds_number_wall"] / re_c
    vertex = channel_vertex_grid(2.0, cfg["L"], cfg["D"] * a.nz_mult, x, cfg["resolution_y"] // 2, 1, z)
    dom = Box3DDomain(vertex, closed=(False, True, False), viscosity=visc)
    if a.plain:
        from fluidgym_b200.box3d import BatchedPISO3D

        class _Plain(BatchedPISO3D):              # same surface as SlabPISO3D for this script
            def load_global(self, u, p):
                self.u.copy_(torch.from_numpy(u).cuda().unsqueeze(0)); self.p.copy_(torch.from_numpy(p).cuda().unsqueeze(0))

            def owned(self, t):
                return t

            def error(self):
                return 0

            def close(self):
                pass
        assert world == 1
        slab = _Plain(dom, 1, device=f"cuda:{local}")
        slab.tabs = type("T", (), {"N": dom.N, "nzl": dom.nz})()
    else:
        slab = SlabPISO3D(dom, rank, world, f"cuda:{local}")
    tb = slab.tabs
    # state: Reichardt profile + 5 % noise, generated identically on every rank
    cc = dom.cell_centres()
    wd = (1 - np.abs(cc[1, 0, :, 0])) * u_wall / visc
    prof = ((1 / 0.41) * np.log(1 + 0.41 * wd) + 7.8 * (1 - np.exp(-wd / 11.0) - (wd / 11.0) * np.exp(-wd / 3))) * u_wall
    rng = np.random.default_rng(42)
    u = 0.05 * rng.standard_normal((3, dom.nz, dom.ny, dom.nx)).astype(np.float32)
    u[0] += prof[None, :, None].astype(np.float32)
    slab.load_global(u.reshape(3, -1), np.zeros(dom.N, np.float32))
    cells = np.arange(tb.N).reshape(tb.nzl, dom.ny, dom.nx)
    rows = torch.from_numpy(np.stack([cells[:, 0, :].ravel(), cells[:, -1, :].ravel()]).astype(np.int32)).cuda()
    pos_y = cc[1].mean(axis=(0, 2))
    d_lo, d_hi = float(1 + pos_y[0]), float(1 - pos_y[-1])
    step_length = cfg["step_length"] * (visc / u_wall ** 2)
    dt, cfl = step_length / 10, cfg["adaptive_cfl"]
    for _ in range(a.warmup):
        slab.single_step(dt, cfl, rows, d_lo, d_hi)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    it0 = slab.buffer("iter_total").clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nsub = 0
    for _ in range(a.steps):
        nsub += slab.single_step(dt, cfl, rows, d_lo, d_hi)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    it = (slab.buffer("iter_total") - it0)[0].tolist()
    err = slab.error()
    chk = float(slab.owned(slab.u).double().pow(2).sum())
    chk_t = torch.tensor([chk], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(chk_t)
    rec = {"workload": f"TCF {'Large' if a.large else 'Small'} {dom.nx}x{dom.ny}x{dom.nz} = {dom.N} cells, 1 environment", "n_gpus": world,
           "mode": "plain" if a.plain else "slab", "solver_steps": a.steps, "substeps": nsub, "ms_total": float(ms),
           "ms_per_substep": float(ms) / max(nsub, 1), "substeps_per_s": nsub / (float(ms) / 1e3),
           "cg_iters_per_solve": it[0] / max(2 * nsub, 1), "bicg_iters_per_rhs": it[1] / max(3 * nsub, 1), "slab_error": err,
           "checksum_u2": float(chk_t)}
    slab.close()
    if world > 1:
        dist.barrier()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--large", action="store_true")
    ap.add_argument("--nz-mult", type=int, default=1, help="repeat the domain along z (weak-scaling style sizes)")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--plain", action="store_true", help="world 1 only: the plain single-GPU solver on the same state (baseline of the protocol overhead)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rec = run(a.large, a.nz_mult, a.steps, a.warmup, a.plain, rank, world, local)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
