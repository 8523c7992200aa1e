"""Worker of the multi-rank GPU test (tests/test_gpu_dist.py): one process per GPU under torch.distributed.run.

Builds the same GpRegressor on every rank with distributed=True (block-column-cyclic factor over NCCL), and checks on
rank 0 -- which also holds a single-GPU regressor of the same model -- that the log marginal likelihood, alpha and the
predictions of every rank's slab of query points agree with the single-GPU path and with the oracle."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import inference_tools_b200.gp as gp  # noqa: E402
from inference_tools_b200.sharding import shard_range  # noqa: E402


def main():
    n, d, block, m_total = (int(a) for a in sys.argv[1:5])
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 1, (n, d))
    y = np.sin(3 * x).sum(axis=1) + rng.normal(0, 0.05, n)
    e = np.full(n, 0.05)
    theta = np.array([0.2, 0.1] + [np.log(0.3)] * d)
    q_all = np.random.default_rng(6).uniform(0, 1, (m_total, d))
    lo, hi = shard_range(m_total, rank, world)
    if rank == world - 1:
        hi = lo                                   # the last rank passes an empty slab: it must still take part
    m = gp.GpRegressor(x, y, y_err=e, hyperpars=theta, device=local, distributed=True, dist_block=block)
    mu, sig = m(q_all[lo:hi])
    lml = m.marginal_likelihood(theta)
    alpha = m.alpha
    theta2 = theta + 0.1
    lml2 = m.marginal_likelihood(theta2)          # re-factors at theta2 ...
    lml_g, grad = m.marginal_likelihood_gradient(theta2)   # collective: row blocks of K^-1 per rank, traces, all-reduce
    mu_b, sig_b = m(q_all[lo:hi])                 # ... and predicting must bring the fit's own factor back
    out = {"rank": rank, "ok": True}
    if rank == 0:
        from oracle import gp_oracle as orc
        s = gp.GpRegressor(x, y, y_err=e, hyperpars=theta, device=local)
        mu_s, sig_s = s(q_all[lo:hi])
        rel = lambda a, b: float(np.abs(np.asarray(a) - b).max() / np.abs(b).max())
        out.update(lml=float(lml), lml_single=float(s.marginal_likelihood(theta)), lml2=float(lml2),
                   lml2_single=float(s.marginal_likelihood(theta2)), alpha_err=rel(alpha, s.alpha), mu_err=rel(mu, mu_s),
                   sig_err=float(np.abs(sig / sig_s - 1).max()), repeat_mu_err=rel(mu_b, mu), repeat_sig_err=rel(sig_b, sig))
        lml_gs, grad_s = s.marginal_likelihood_gradient(theta2)
        out.update(grad_err=rel(grad, grad_s), lml_grad_err=abs(float(lml_g) - float(lml_gs)) / abs(float(lml_gs)))
        if n <= 4096:
            ref = orc.Fit(x, y, ("SE",), "const", theta, e**2)
            mu_o, sig_o = ref.predict(q_all[lo:hi])
            out.update(alpha_vs_oracle=rel(alpha, ref.alpha), mu_vs_oracle=rel(mu, mu_o),
                       sig_vs_oracle=float(np.abs(sig / sig_o - 1).max()),
                       lml_vs_oracle=abs(lml - orc.marginal_likelihood(x, y, ("SE",), "const", theta, e**2)) / abs(lml))
        print("DIST_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    m.engine.dist_finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
