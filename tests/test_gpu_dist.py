"""Distributed regressor (csrc/dist.cu) on the GPU: the single-rank code path on one GPU, and -- when the box has at
least two GPUs -- two ranks over NCCL launched with torch.distributed.run (tests/dist_worker.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, make_kernel, make_mean, rel_err, synth
import inference_tools_b200.gp as gp
from inference_tools_b200 import _lib
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.mark.parametrize("n,block", [(700, 128), (1500, 256), (2048, 512), (4096, 1024)])
def test_block_cyclic_regressor_single_rank(n, block):
    """world = 1 runs the same sweep, back-substitution and panel-streamed predict as the multi-rank case (the NCCL calls
    become device copies): LML, alpha, mu / sigma against the dense single-GPU path and the oracle."""
    d = 2
    x, y, e = synth(77 + n, n, d)
    theta = np.array([0.2, 0.1, np.log(0.3), np.log(0.25)])
    q = np.random.default_rng(n).uniform(0, 1, (600, d))
    m = gp.GpRegressor(x, y, y_err=e, hyperpars=theta, distributed=(0, 1, None), dist_block=block)
    s = gp.GpRegressor(x, y, y_err=e, hyperpars=theta)
    ref = orc.Fit(x, y, ("SE",), "const", theta, e**2)
    lml_o = orc.marginal_likelihood(x, y, ("SE",), "const", theta, e**2)
    assert abs(m.marginal_likelihood(theta) - lml_o) <= TOL * abs(lml_o)
    assert rel_err(m.alpha, ref.alpha) < TOL and rel_err(m.alpha, s.alpha) < TOL
    mu, sig = m(q)
    mu_o, sig_o = ref.predict(q)
    assert rel_err(mu, mu_o) < TOL and np.abs(sig / sig_o - 1).max() < TOL
    mu_s, sig_s = s(q)
    assert rel_err(mu, mu_s) < 1e-11 and np.abs(sig / sig_s - 1).max() < 1e-10
    assert m(q[:0])[0].shape == (0,)                                  # an empty slab is legal (ranks without queries)
    # a different theta re-factors; the fitted state comes back for the next prediction
    th2 = theta + 0.1
    assert abs(m.marginal_likelihood(th2) - s.marginal_likelihood(th2)) <= TOL * abs(lml_o)
    mu2, _ = m(q)
    assert rel_err(mu2, mu) < 1e-13
    # gradient: rows of K^-1 by streamed solves + Y_a Y_b^T products, traces per row block (regression.py:544-567)
    lml_g, grad = m.marginal_likelihood_gradient(th2)
    lml_s, grad_s = s.marginal_likelihood_gradient(th2)
    lml_o2, grad_o = orc.marginal_likelihood_gradient(x, y, ("SE",), "const", th2, e**2)
    assert abs(lml_g - lml_o2) <= TOL * abs(lml_o2)
    assert np.abs(grad - grad_o).max() <= TOL * np.abs(grad_o).max()
    assert np.abs(grad - grad_s).max() <= 1e-10 * np.abs(grad_s).max()
    mu3, _ = m(q)                                                      # and the fit's own factor is restored afterwards
    assert rel_err(mu3, mu) < 1e-13


@pytest.mark.parametrize("comps,mean,d,n,block", [(("RQ", "WHITE"), "linear", 3, 1300, 256), (("SE", "HETERO"), "const", 1, 900, 128),
                                                  (("SE", "RQ"), "quadratic", 2, 2100, 512)])
def test_block_cyclic_gradient_other_models(comps, mean, d, n, block):
    """Every kernel / mean family the distributed path accepts through its gradient (single rank), against the dense
    single-GPU gradient and the oracle; sizes that are not a multiple of the block included."""
    x, y, e = synth(5 + n, n, d)
    tm = {"const": [0.3], "linear": [0.3] + [0.1] * d, "quadratic": [0.3] + [0.1] * d + [-0.05] * d}[mean]
    tc = []
    for c in comps:
        tc += {"SE": [0.1] + [np.log(0.35)] * d, "RQ": [-0.2, 0.8] + [np.log(0.3)] * d, "WHITE": [np.log(0.04)],
               "HETERO": list(np.log(0.03 + 0.02 * np.random.default_rng(1).uniform(size=n)))}[c]
    theta = np.array(tm + tc)
    noise = None if "HETERO" in comps else e
    s = gp.GpRegressor(x, y, y_err=noise, kernel=make_kernel(gp, comps), mean=make_mean(gp, mean), hyperpars=theta)
    m = gp.GpRegressor(x, y, y_err=noise, kernel=make_kernel(gp, comps), mean=make_mean(gp, mean), hyperpars=theta,
                       distributed=(0, 1, None), dist_block=block)
    lml_g, grad = m.marginal_likelihood_gradient(theta)
    lml_s, grad_s = s.marginal_likelihood_gradient(theta)
    assert grad.shape == grad_s.shape
    assert abs(lml_g - lml_s) <= TOL * abs(lml_s)
    assert np.abs(grad - grad_s).max() <= TOL * np.abs(grad_s).max()
    lml_o, grad_o = orc.marginal_likelihood_gradient(x, y, comps, mean, theta, None if noise is None else e**2)
    assert np.abs(grad - grad_o).max() <= TOL * np.abs(grad_o).max()


def test_block_cyclic_fit_without_hyperpars_runs_the_optimiser_in_lockstep():
    """distributed=... without hyperpars: L-BFGS restarts from start points every rank derives from the data, every
    evaluation a collective gradient call; the optimum must equal the single-GPU optimum from the same start points."""
    n, d = 900, 2
    x, y, e = synth(123, n, d)
    m = gp.GpRegressor(x, y, y_err=e, distributed=(0, 1, None), dist_block=256, n_starts=2)
    s = gp.GpRegressor(x, y, y_err=e, hyperpars=m.hyperpars)
    lwr, upr = (np.array([b[i] for b in s.hp_bounds]) for i in (0, 1))
    x0s = [lwr + (upr - lwr) * np.random.RandomState(m._lockstep_seed()).random_sample(size=len(lwr)), 0.5 * (lwr + upr)]
    best = sorted((s.launch_bfgs(x0) for x0 in x0s), key=lambda r: r[1])[0]
    assert abs(s.marginal_likelihood(m.hyperpars) + best[1]) <= 1e-7 * abs(best[1])
    with pytest.raises(ValueError):
        gp.GpRegressor(x, y, y_err=e, distributed=(0, 1, None), n_processes=2)


def test_block_cyclic_regressor_two_ranks_over_nccl():
    if _lib.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    port = 29500 + os.getpid() % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"), "3000", "2", "256", "900"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("DIST_RESULT ")][-1]
    out = json.loads(line[len("DIST_RESULT "):])
    assert abs(out["lml"] - out["lml_single"]) <= TOL * abs(out["lml_single"])
    assert abs(out["lml2"] - out["lml2_single"]) <= TOL * abs(out["lml2_single"])
    for k in ("alpha_err", "mu_err", "sig_err", "alpha_vs_oracle", "mu_vs_oracle", "sig_vs_oracle", "lml_vs_oracle", "grad_err",
              "lml_grad_err"):
        assert out[k] < TOL, (k, out)
    assert out["repeat_mu_err"] < 1e-13 and out["repeat_sig_err"] < 1e-13
