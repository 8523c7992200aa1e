"""Single-GPU gradient against the distributed gradient on one rank (which sits at 1.6e-11 of the CPU oracle on this data,
profiles/dist_grad_parity_r2.json) on dense SquaredExponential 2-D sets -- a GPU-only proxy for the oracle comparison
that takes seconds; also the lauum phase time.  gpurun; output gpurun_out/grad_chunk_check.json."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from inference_tools_b200 import _lib
out = []
for n in (8192, 16384):
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 1, (n, 2))
    y = np.sin(3 * x).sum(axis=1) + rng.normal(0, 0.05, n)
    theta = np.array([0.2, 0.1, np.log(0.3), np.log(0.3)])
    eng = _lib.Engine(0); eng.set_data(x, y, np.full(n, 0.05**2)); eng.set_model([_lib.COV_SE], _lib.MEAN_CONST)
    eng.dist_init(0, 1, None)
    lml_d, grad_d, info, t = eng.dist_lml_grad(theta, 1024)
    eng.lml_grad(theta)
    lml_s, grad_s, info_s = eng.lml_grad(theta)
    tm = eng.timers()
    with _lib.options(gemm_i8=0):
        lml_f, grad_f, _ = eng.lml_grad(theta)
    row = {"n": n, "single_vs_dist": float(np.abs(grad_s - grad_d).max() / np.abs(grad_d).max()),
           "dmma_vs_dist": float(np.abs(grad_f - grad_d).max() / np.abs(grad_d).max()), "lauum_ms": tm.get("lauum"), "trtri_ms": tm.get("trtri"),
           "guard_retries": eng.stat("grad_guard_retries")}
    out.append(row); print(json.dumps(row), flush=True)
    eng.dist_finalize(); eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/grad_chunk_check.json", "w"), indent=1)
