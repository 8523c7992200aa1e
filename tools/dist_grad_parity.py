"""Accuracy of the distributed gradient (gpb_dist_lml_grad, single rank) against the CPU oracle on a dense, ill-conditioned
data set (SquaredExponential 2-D: the case where the INT8 inverse chain is most exposed), beside the single-GPU gradient;
default dispatch and DMMA only.  gpurun; output gpurun_out/dist_grad_parity.json."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from inference_tools_b200 import _lib
from oracle import gp_oracle as orc

out = {"cases": []}
for n, d, block in [(2048, 2, 256), (8192, 2, 1024), (16384, 2, 1024), (16384, 3, 1024)]:
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 1, (n, d))
    y = np.sin(3 * x).sum(axis=1) + rng.normal(0, 0.05, n)
    e2 = np.full(n, 0.05**2)
    theta = np.array([0.2, 0.1] + [np.log(0.3)] * d)
    t0 = time.perf_counter()
    lml_o, grad_o = orc.marginal_likelihood_gradient_blocked(x, y, ("SE",), "const", theta, e2)
    case = {"n": n, "d": d, "block": block, "oracle_seconds": time.perf_counter() - t0}
    rel = lambda g: float(np.abs(g - grad_o).max() / np.abs(grad_o).max())
    for mode, opts in (("default", {}), ("dmma", {"gemm_i8": 0})):
        with _lib.options(**opts):
            eng = _lib.Engine(0)
            eng.set_data(x, y, e2)
            eng.set_model([_lib.COV_SE], _lib.MEAN_CONST)
            eng.dist_init(0, 1, None)
            lml_d, grad_d, info, t = eng.dist_lml_grad(theta, block)
            lml_d, grad_d, info, t = eng.dist_lml_grad(theta, block)
            single = {}
            for inv in (1, 0):      # K^-1 from the rows of L^-T (substitution) / from the recursive triangular inverse
                with _lib.options(grad_inverse=inv):
                    eng.lml_grad(theta)
                    lml_s, grad_s, info_s = eng.lml_grad(theta)
                    tm = eng.timers()
                    single[f"grad_inverse{inv}"] = {"grad_rel_err": rel(grad_s), "lml_rel_err": abs(lml_s - lml_o) / abs(lml_o),
                                                    "trtri_ms": tm.get("trtri"), "lauum_ms": tm.get("lauum"), "total_ms": sum(tm.values()),
                                                    "guard_retries": eng.stat("grad_guard_retries")}
            case[mode] = {"dist_grad_rel_err": rel(grad_d), "single": single, "dist_lml_rel_err": abs(lml_d - lml_o) / abs(lml_o),
                          "factor_s": t["factor_s"], "gradient_s": t["gradient_s"]}
            eng.dist_finalize()
            eng.close()
    out["cases"].append(case)
    print(json.dumps(case), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/dist_grad_parity.json", "w"), indent=1)
