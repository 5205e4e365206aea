"""Solvers on different GPUs inside ONE process: every native call is issued with the solver's own device
current (fluidgym_b200.native.DeviceLib), so the thread-local current device -- cuda:0 in fresh worker threads, the device of the
last constructed solver otherwise -- does not matter.  Needs two GPUs; skipped on a one-GPU box."""
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

needs_two = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")


@needs_two
def test_two_solvers_on_two_devices_in_one_thread():
    import fluidgym_b200 as fg
    e0 = fg.make("CylinderJet2D-easy-v0", n_envs=2, device="cuda:0", resolution=8)
    e1 = fg.make("CylinderJet2D-easy-v0", n_envs=2, device="cuda:1", resolution=8)      # leaves cuda:1 current
    ref = fg.make("CylinderJet2D-easy-v0", n_envs=2, device="cuda:0", resolution=8)
    for e in (e0, e1, ref):
        e.reset(seed=5)
    a = torch.tensor([[0.3], [-0.6]])
    torch.cuda.set_device(1)
    o0, r0, *_ = e0.step(a.to("cuda:0"))          # device 1 current, solver on device 0
    torch.cuda.set_device(0)
    o1, r1, *_ = e1.step(a.to("cuda:1"))          # device 0 current, solver on device 1
    oref, rref, *_ = ref.step(a.to("cuda:0"))
    assert torch.equal(r0.cpu(), rref.cpu()) and torch.equal(r1.cpu(), rref.cpu())
    assert torch.equal(o0["velocity"].cpu(), oref["velocity"].cpu()) and torch.equal(o1["pressure"].cpu(), oref["pressure"].cpu())


@needs_two
def test_parallel_fluid_env_over_two_devices():
    """ParallelFluidEnv(cuda_ids=[0, 1, 0, 1]): worker threads (current device 0) drive the device-1 environments."""
    import fluidgym_b200 as fg
    from fluidgym_b200.envs.parallel_env import ParallelFluidEnv
    env = ParallelFluidEnv("CylinderJet2D-easy-v0", cuda_ids=[0, 1, 0, 1], resolution=8)
    obs, _ = env.reset(seed=11)
    a = torch.tensor([[0.2], [0.4], [-0.2], [-0.4]])
    obs, reward, term, trunc, info = env.step(a)
    assert reward.shape == (4,) and torch.isfinite(reward).all() and obs["velocity"].shape[0] == 4
    # the device-1 environments (entries 1, 3; seed 11 + 1) equal a single-device run with the same seed and actions
    ref = fg.make("CylinderJet2D-easy-v0", n_envs=2, device="cuda:0", resolution=8)
    ref.reset(seed=12)
    o, r, *_ = ref.step(a[[1, 3]].to("cuda:0"))
    assert torch.equal(reward[[1, 3]], r.cpu())
    assert torch.equal(obs["velocity"][[1, 3]], o["velocity"].cpu())
    env.close()
