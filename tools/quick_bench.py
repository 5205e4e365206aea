"""Scratch timing helper (not the contract bench): times fgb_piso_substep for a batch of environments."""
import sys, os, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fluidgym_b200.envs.cylinder_domain import make_cylinder_domain
from fluidgym_b200.solver import BatchedPISO

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    impls = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1]
    fixture = sys.argv[3] if len(sys.argv) > 3 else "cyl24_steps.npz"
    spec = make_cylinder_domain(24)
    t0 = time.time(); cd = spec.prepare(); print("compile tables %.2fs" % (time.time() - t0))
    st = np.load(os.path.join(ROOT, "tests/golden", fixture))
    first = None
    for impl in impls:
        sol = BatchedPISO(cd, B, cg_impl=impl)
        u = torch.from_numpy(st["env3_u"]).cuda(); p = torch.from_numpy(st["env3_p"]).cuda()
        bv = torch.from_numpy(st["env0_bvel"]).cuda()
        gen = torch.Generator(device="cuda").manual_seed(0)
        sol.u.copy_(u.unsqueeze(0).expand_as(sol.u)); sol.p.copy_(p.unsqueeze(0).expand_as(sol.p)); sol.bvel.copy_(bv.unsqueeze(0).expand_as(sol.bvel))
        sol.u += 0.01 * torch.randn(sol.u.shape, device="cuda", generator=gen)
        for _ in range(3):
            sol.piso_substep(0.01)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 5
        e0.record()
        for _ in range(K):
            sol.piso_substep(0.01)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        its = sol.buffer("iters").cpu().numpy()
        print(json.dumps({"impl": impl, "B": B, "ms_per_substep": ms, "env_substeps_per_s": B / ms * 1e3,
                          "cg_iters_mean": float(its[:, 2:4].mean()) + 1, "cg_iters_max": int(its[:, 2:4].max()) + 1,
                          "bicg_iters_mean": float(its[:, :2].mean()) + 1}))
        res = (sol.u.clone(), sol.p.clone())
        if first is None:
            first = res
        else:   # same state, same number of substeps: distance to the first implementation of the list
            print(json.dumps({"impl": impl, "rel_l2_u_vs_first": float((res[0] - first[0]).norm() / first[0].norm()),
                              "rel_l2_p_vs_first": float((res[1] - first[1]).norm() / first[1].norm()),
                              "resid_max": float(sol.buffer("resid")[:, 2:4].max())}))
        del sol
main()
