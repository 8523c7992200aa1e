"""CPU: the lockstep restart driver (gp/_lockstep.py) -- many scipy L-BFGS-B runs answered by one batched call per round
must return exactly what the same runs return one after the other."""
import threading

import numpy as np
import pytest
from scipy.optimize import fmin_l_bfgs_b

from inference_tools_b200.gp._lockstep import lockstep_lbfgs


def f_single(x):
    x = np.asarray(x, dtype=float)
    val = np.sin(3 * x).sum() + 0.1 * (x**2).sum() + 0.05 * x[0] * x[-1]
    grad = 3 * np.cos(3 * x) + 0.2 * x
    grad[0] += 0.05 * x[-1]
    grad[-1] += 0.05 * x[0]
    return val, grad


class Batched:
    def __init__(self):
        self.calls, self.points, self.lock = 0, 0, threading.Lock()

    def __call__(self, pts):
        with self.lock:
            self.calls += 1
            self.points += len(pts)
        out = [f_single(p) for p in pts]
        return np.array([o[0] for o in out]), np.stack([o[1] for o in out])


@pytest.mark.parametrize("n_starts,d,max_concurrent", [(1, 2, 256), (7, 2, 256), (40, 3, 256), (9, 2, 4)])
def test_lockstep_restarts_equal_sequential_restarts(n_starts, d, max_concurrent):
    rng = np.random.default_rng(n_starts)
    bounds = [(-2.0, 2.0)] * d
    x0s = rng.uniform(-2, 2, (n_starts, d))
    batched = Batched()
    got = lockstep_lbfgs(batched, x0s, bounds, max_concurrent=max_concurrent, pgtol=1e-10)
    want = [fmin_l_bfgs_b(f_single, x0, approx_grad=False, bounds=bounds, pgtol=1e-10) for x0 in x0s]
    assert len(got) == n_starts
    for g, w in zip(got, want):
        assert np.array_equal(g[0], w[0]) and g[1] == w[1]                       # the same trajectory, bit for bit
        assert g[2]["funcalls"] == w[2]["funcalls"] and g[2]["warnflag"] == w[2]["warnflag"]
    total = sum(w[2]["funcalls"] for w in want)
    assert batched.points == total                                              # nothing evaluated twice
    if n_starts >= 7 and max_concurrent >= n_starts:
        assert batched.calls <= max(w[2]["funcalls"] for w in want)            # one batched call per round of the slowest


def test_lockstep_propagates_an_error_of_the_batched_call_and_releases_its_workers():
    def broken(pts):
        raise NotImplementedError("no gradient terms for this kernel")

    before = threading.active_count()
    with pytest.raises(NotImplementedError):
        lockstep_lbfgs(broken, np.zeros((5, 2)), [(-1.0, 1.0)] * 2)
    assert threading.active_count() == before


def test_lockstep_error_in_a_later_round():
    calls = {"n": 0}

    def flaky(pts):
        calls["n"] += 1
        if calls["n"] == 3:
            raise RuntimeError("device lost")
        out = [f_single(p) for p in pts]
        return np.array([o[0] for o in out]), np.stack([o[1] for o in out])

    with pytest.raises(RuntimeError, match="device lost"):
        lockstep_lbfgs(flaky, np.random.default_rng(0).uniform(-1, 1, (6, 2)), [(-2.0, 2.0)] * 2)
