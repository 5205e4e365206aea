"""Forward / backward wall time of one differentiable env.step of the D = 3 families (the configurations of the reference gradient goldens);
prints one JSON line per family.  The reference's own times are in the *_grad_meta.json files written by oracle/ref_grad_harness.py."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluidgym_b200 as fg
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def timed(env, act, leaves):
    for rep in range(2):                       # first pass = warm-up (allocator, first cooperative launches)
        state = env.get_state() if hasattr(env, "get_state") else None
        torch.cuda.synchronize(); t0 = time.time()
        obs, reward, *_ = env.step(act)
        torch.cuda.synchronize(); t1 = time.time()
        torch.autograd.grad(reward.sum(), [act] + leaves, allow_unused=True)
        torch.cuda.synchronize(); t2 = time.time()
        if rep == 0:
            return_first = (t1 - t0, t2 - t1)
            break
    return return_first


def main():
    out = []
    fx = np.load(os.path.join(G, "tcf32_grad.npz"))
    for rep in range(2):
        env = fg.make("TCFSmall3D-both-easy-v0", n_envs=1, resolution_x_z=32, resolution_y=33, differentiable=True)
        env.reset(seed=42); env.set_state(fx["pre_u"], np.zeros(32768, np.float32), np.zeros((3, 2048), np.float32))
        u0 = env.mark_state_differentiable()
        act = torch.from_numpy(fx["action"]).cuda().reshape(1, 512, 1).clone().requires_grad_(True)
        f, b = timed(env, act, [u0])
    out.append(dict(family="TCF 32x33x32, one env.step = 10 solver steps", forward_s=f, backward_s=b, reference_forward_s=0.32, reference_backward_s=2.73))
    fx = np.load(os.path.join(G, "cyl3d_grad.npz"))
    for rep in range(2):
        env = fg.make("CylinderJet3D-easy-v0", n_envs=1, resolution=8, n_jets=8, differentiable=True)
        env.seed(42); env.set_state(fx["pre_u"], fx["pre_p"], fx["pre_bvel"], last_control=np.zeros((1, 8), np.float32))
        u0 = env.mark_state_differentiable()
        act = torch.from_numpy(fx["action"]).cuda().reshape(1, 8, 1).clone().requires_grad_(True)
        f, b = timed(env, act, [u0])
    out.append(dict(family="CylinderJet3D res 8, one env.step = 25 solver steps", forward_s=f, backward_s=b, reference_forward_s=18.09, reference_backward_s=5.08))
    fx = np.load(os.path.join(G, "airfoil3d_grad.npz"))
    env = fg.make("Airfoil3D-easy-v0", n_envs=1, res_z=8, n_agents=4, init_from_2d=False, differentiable=True)
    env.seed(42); env.set_state(fx["pre_u"], fx["pre_p"], fx["pre_bvel"], last_control=0.0)
    u0 = env.mark_state_differentiable()
    act = torch.from_numpy(fx["action"]).cuda().reshape(env._zero_action.shape).clone().requires_grad_(True)
    f, b = timed(env, act, [u0])
    out.append(dict(family="Airfoil3D res_z 8, one env.step = 5 solver steps (first call, no warm-up)", forward_s=f, backward_s=b, reference_forward_s=165.7, reference_backward_s=221.2))
    for o in out:
        print(json.dumps(o))


main()
