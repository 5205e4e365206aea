"""torchrun worker of tests/test_gpu_slab.py: the slab-decomposed D = 3 solver on `world` GPUs against the single-GPU solver on
the same state (the reference's traced 32^3 channel state, tests/golden/tcf32_substep0.npz).  Prints SLAB_OK on success."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fluidgym_b200.box3d import BatchedPISO3D, Box3DDomain, SlabPISO3D  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    sgs = len(sys.argv) > 2 and sys.argv[2] == "sgs"          # Smagorinsky model with wall damping on: the viscosity halo exchange
    G = os.path.join(ROOT, "tests", "golden")
    g, fx = np.load(os.path.join(G, "tcf32_geometry.npz")), np.load(os.path.join(G, "tcf32_substep0.npz"))
    visc = json.load(open(os.path.join(G, "tcf32_meta.json")))["viscosity"]
    dom = Box3DDomain(g["vertex"], closed=(False, True, False), viscosity=visc)
    dt = float(fx["dt"][0])
    slab = SlabPISO3D(dom, rank, world, f"cuda:{local}")
    slab.load_global(fx["u_in"], fx["p_in"], {2: fx["bvel2"], 3: fx["bvel3"]})
    damp = None
    if sgs:
        vd = 1 - np.exp(-(1 - np.abs(dom.cell_centres()[1].astype(np.float32))) * np.float32(180.0 / 25.0))
        damp = (vd * vd).astype(np.float32).reshape(-1)
        slab.set_sgs(0.1, damp)
    src = torch.zeros(1, 4, device="cuda")
    src[0, :3] = torch.from_numpy(fx["src"]).cuda()
    for _ in range(steps):
        slab.piso_substep(dt, src)
    torch.cuda.synchronize()
    err = slab.error()
    its = slab.buffer("iters")[0].tolist()
    # gather the owned parts on rank 0
    mine_u = slab.owned(slab.u)[0].contiguous()
    mine_p = slab.owned(slab.p)[0].contiguous()
    us = [torch.empty_like(mine_u) for _ in range(world)] if rank == 0 else None
    ps = [torch.empty_like(mine_p) for _ in range(world)] if rank == 0 else None
    dist.gather(mine_u, us, dst=0)
    dist.gather(mine_p, ps, dst=0)
    ok = err == 0
    if rank == 0:
        P = dom.nx * dom.ny
        u_all = torch.cat([x.reshape(3, -1, P) for x in us], dim=1).reshape(3, -1)
        p_all = torch.cat([x.reshape(-1, P) for x in ps], dim=0).reshape(-1)
        ref = BatchedPISO3D(dom, 1, device=f"cuda:{local}")
        ref.u.copy_(torch.from_numpy(fx["u_in"]).cuda().unsqueeze(0))
        ref.p.copy_(torch.from_numpy(fx["p_in"]).cuda().unsqueeze(0))
        ref.bvel.copy_(torch.from_numpy(np.concatenate([fx["bvel2"], fx["bvel3"]], axis=1)).cuda().unsqueeze(0))
        if sgs:
            ref.set_sgs(0.1, damp)
        for _ in range(steps):
            ref.piso_substep(dt, src)
        torch.cuda.synchronize()
        eu = float((u_all - ref.u[0]).norm() / ref.u[0].norm())
        ep = float((p_all - ref.p[0]).norm() / ref.p[0].norm())
        its_ref = ref.buffer("iters")[0].tolist()
        print(json.dumps({"world": world, "steps": steps, "sgs": sgs, "rel_err_u": eu, "rel_err_p": ep, "iters": its, "iters_single": its_ref, "slab_error": err}))
        ok = ok and eu < 1e-5 and ep < 1e-3 and its[:5] == its_ref[:5]
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    slab.close()
    dist.barrier()
    if rank == 0:
        print("SLAB_OK" if int(flag) == 1 else "SLAB_FAILED")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
