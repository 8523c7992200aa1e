"""Distributed regressor (csrc/dist.cu) on the GPU: the single-rank code path on one GPU, and -- when the box has at
least two GPUs -- two ranks over NCCL launched with torch.distributed.run (tests/dist_worker.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, rel_err, synth
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
    with pytest.raises(NotImplementedError):
        m.marginal_likelihood_gradient(theta)
    with pytest.raises(ValueError):
        gp.GpRegressor(x, y, y_err=e, distributed=(0, 1, None))       # hyper-parameters must be given
    # a different theta re-factors; the fitted state comes back for the next prediction
    th2 = theta + 0.1
    assert abs(m.marginal_likelihood(th2) - s.marginal_likelihood(th2)) <= TOL * abs(lml_o)
    mu2, _ = m(q)
    assert rel_err(mu2, mu) < 1e-13


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
    for k in ("alpha_err", "mu_err", "sig_err", "alpha_vs_oracle", "mu_vs_oracle", "sig_vs_oracle", "lml_vs_oracle"):
        assert out[k] < TOL, (k, out)
    assert out["repeat_mu_err"] < 1e-13 and out["repeat_sig_err"] < 1e-13
