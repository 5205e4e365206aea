"""Host logic of envs/common.py::LinearSolveWatch on the CPU: the residual table of an env.step is examined one step later
(or on demand), a non-finite entry raises LinsolveError naming the environments, and an error is reported once
(reference behaviour: _check_solver_return_infos, PISOtorch_diff.py:280-300).  The GPU path with real solves is
tests/test_gpu_all_envs.py::test_non_finite_solve_raises_linsolve_error."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from fluidgym_b200.envs.common import LinearSolveWatch, LinsolveError  # noqa: E402


class _Solver:
    def __init__(self, B):
        self.resid = torch.zeros(B, 8)

    def buffer(self, name):
        assert name == "resid"
        return self.resid


class _Env(LinearSolveWatch):
    def __init__(self, B=3):
        self.solver, self._n_steps = _Solver(B), 0

    def step(self, residuals):
        self.solver.resid[:] = torch.as_tensor(residuals, dtype=torch.float32)
        self._n_steps += 1
        self._watch_linear_solves()


def test_healthy_steps_return_the_table():
    env = _Env()
    assert env.check_linear_solves() is None                   # nothing watched yet
    env.step(np.full((3, 8), 1e-6))
    t = env.check_linear_solves()
    assert t.shape == (3, 8) and np.allclose(t, 1e-6)
    assert env.check_linear_solves() is None                   # examined once


def test_non_finite_residual_raises_one_step_late_and_names_the_environment():
    env = _Env()
    env.step(np.full((3, 8), 1e-6))
    bad = np.full((3, 8), 1e-6)
    bad[2, 3] = np.nan
    bad[0, 4] = np.inf
    env.step(bad)                                              # examines step 1 (healthy), stores step 2
    with pytest.raises(LinsolveError) as err:
        env.step(np.full((3, 8), 1e-6))                        # examines step 2
    msg = str(err.value)
    assert "env.step 2" in msg and "[0, 2]" in msg and "of 3" in msg and "[3, 4]" in msg
    assert env.check_linear_solves() is None                   # reported once; the raising step stored nothing
    env.step(np.full((3, 8), 1e-6))
    assert env.check_linear_solves() is not None


def test_switch_and_stub_solvers():
    env = _Env()
    env.check_solves = False
    env.step(np.full((3, 8), np.nan))
    assert env.check_linear_solves() is None
    env = _Env()
    env.solver = object()                                      # host stand-in without a residual table: nothing to watch
    env._n_steps += 1
    env._watch_linear_solves()
    assert env.check_linear_solves() is None
