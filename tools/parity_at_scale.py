"""Parity of the CUDA engine against the CPU oracle at the sizes the INT8 tensor-core GEMM path runs (gpurun only).

    python tools/parity_at_scale.py scale 8192 16384 32768      # alpha, mu/sigma (4096 points), LML, grad LML vs oracle
    python tools/parity_at_scale.py cond [N]                     # conditioning sweep: INT8 / DMMA / oracle, cond(K) 1e3 -> 1e12
    python tools/parity_at_scale.py cfg1                         # N = 200 latency of one LML-gradient evaluation

The oracle side is oracle/gp_oracle.py (memory-lean, row-blocked; the same LAPACK / BLAS calls as the reference:
numpy.linalg.cholesky, scipy.linalg.solve_triangular on the identity, iK.T @ iK).  Results are written as JSON under
gpurun_out/ and copied into profiles/ by the builder.  Test / measurement infrastructure: not imported by the product.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from inference_tools_b200 import _lib  # noqa: E402
import inference_tools_b200.gp as gp  # noqa: E402
from oracle import gp_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
MODES = {"default": {}, "dmma": {"gemm_i8": 0}, "int8_forced": {"gemm_i8": 2, "gemm_i8_min_k": 64}}


def synth(seed, n, d, sigma_n=0.05):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (n, d))
    y = np.sin(3 * x).sum(axis=1) + rng.normal(0, sigma_n, n)
    return x, y, np.full(n, sigma_n)


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def engine_results(x, y, e, kernel, theta, q, mode):
    """alpha, mu, sigma, LML (value-only and gradient path), grad, INT8 share of the GEMM flops, under one GEMM mode"""
    with _lib.options(**MODES[mode]):
        lib = _lib.load_library()
        f0, i0 = lib.gpb_gemm_flops(), lib.gpb_gemm_flops_int8()
        t0 = time.perf_counter()
        try:
            m = gp.GpRegressor(x, y, y_err=e, kernel=kernel, hyperpars=theta)
        except np.linalg.LinAlgError:
            return {"failed": "LinAlgError in set_hyperparameters"}
        mu, sig = m(q)
        lml_v = m.marginal_likelihood(theta)
        try:
            lml, grad = m.marginal_likelihood_gradient(theta)
        except np.linalg.LinAlgError:
            return {"failed": "LinAlgError in marginal_likelihood_gradient"}
        alpha = m.alpha
        secs = time.perf_counter() - t0
        f1, i1 = lib.gpb_gemm_flops(), lib.gpb_gemm_flops_int8()
        stats = {k: m.engine.stat(k) for k in ("dmma_retries", "grad_guard_retries", "grad_guard_est", "predict_block")}
        m.engine.close()
    return {"alpha": alpha, "mu": mu, "sig": sig, "lml_value_only": float(lml_v), "lml": float(lml), "grad": grad,
            "int8_share": (i1 - i0) / max(f1 - f0, 1.0), "seconds": secs, "stats": stats}


def oracle_results(x, y, e, comps, theta, q):
    t0 = time.perf_counter()
    lml, grad, alpha_inv = orc.marginal_likelihood_gradient_blocked(x, y, comps, "const", theta, e**2, want_parts=True)
    t1 = time.perf_counter()
    fit = orc.Fit(x, y, comps, "const", theta, e**2)
    mu, sig = fit.predict(q)
    lml_v = orc.marginal_likelihood(x, y, comps, "const", theta, e**2) if x.shape[0] <= 16384 else None
    t2 = time.perf_counter()
    return {"alpha": fit.alpha, "mu": mu, "sig": sig, "lml": float(lml), "lml_value_only": lml_v, "grad": grad,
            "seconds_grad": t1 - t0, "seconds_fit_predict": t2 - t1, "L_diag": np.diagonal(fit.L).copy()}


def compare(r, o):
    if "failed" in r:
        return r
    out = {"alpha": rel(r["alpha"], o["alpha"]), "mu": rel(r["mu"], o["mu"]), "sigma": float(np.abs(r["sig"] / o["sig"] - 1).max()),
           "lml": abs(r["lml"] - o["lml"]) / abs(o["lml"]), "grad": rel(r["grad"], o["grad"]),
           "int8_share": r["int8_share"], "engine_seconds": r["seconds"], "stats": r["stats"],
           "grad_max_abs": float(np.abs(o["grad"]).max())}
    if o["lml_value_only"] is not None:
        out["lml_value_only"] = abs(r["lml_value_only"] - o["lml_value_only"]) / abs(o["lml_value_only"])
    return out


def cases_for(n):
    d3 = (3, ("SE",), gp.SquaredExponential(), np.array([0.3, 0.1] + [np.log(0.35)] * 3))
    d5 = (5, ("RQ", "WHITE"), gp.RationalQuadratic() + gp.WhiteNoise(), np.array([0.2, 0.1, 1.0] + [np.log(0.3)] * 5 + [np.log(0.05)]))
    return [d3, d5] if n <= 16384 else [d5]


def run_scale(sizes):
    res = {"what": "engine vs CPU oracle (relative errors; grad norm-relative), per GEMM mode", "cases": []}
    for n in sizes:
        for d, comps, kernel, theta in cases_for(n):
            x, y, e = synth(2024 + n, n, d)
            q = np.random.default_rng(n).uniform(0, 1, (4096, d))      # enough rows for the blocked predict solve
            o = oracle_results(x, y, e, comps, theta, q)
            entry = {"N": n, "d": d, "kernel": "+".join(comps), "oracle_seconds": o["seconds_grad"] + o["seconds_fit_predict"],
                     "cond_lower_bound_from_L": float((o["L_diag"].max() / o["L_diag"].min()) ** 2), "lml": o["lml"]}
            for mode in ("default", "dmma"):
                entry[mode] = compare(engine_results(x, y, e, kernel, theta, q, mode), o)
            print(json.dumps(entry), flush=True)
            res["cases"].append(entry)
            json.dump(res, open(os.path.join(OUT, "parity_at_scale.json"), "w"), indent=1)
    return res


def run_cond(n=4096):
    """SE 3-D, noise sigma_n 0.1 -> 1e-4 and length-scales 0.3 / 1.0: cond(K) from ~1e3 to the jitter floor."""
    d = 3
    res = {"what": f"conditioning sweep at N={n}, SE d=3: relative errors vs the oracle per GEMM mode; cond = eigvalsh ratio", "cases": []}
    for ls in (0.3, 1.0):
        for sn in (0.1, 0.03, 0.01, 0.003, 0.001, 0.0003, 0.0001):
            x, y, e = synth(7, n, d, sn)
            theta = np.array([0.3, 0.1] + [np.log(ls)] * d)
            q = np.random.default_rng(1).uniform(0, 1, (4096, d))      # enough rows for the blocked predict solve
            tm, parts = orc.split_theta(theta, ("SE",), "const", n, d)
            w = np.linalg.eigvalsh(orc.train_cov(("SE",), parts, x, e**2))
            o = oracle_results(x, y, e, ("SE",), theta, q)
            entry = {"l": ls, "sigma_n": sn, "cond": float(w[-1] / w[0]), "lambda_min": float(w[0]),
                     "cond_lower_bound_from_L": float((o["L_diag"].max() / o["L_diag"].min()) ** 2)}
            for mode in ("default", "int8_forced", "dmma"):
                entry[mode] = compare(engine_results(x, y, e, gp.SquaredExponential(), theta, q, mode), o)
            print(json.dumps(entry), flush=True)
            res["cases"].append(entry)
            json.dump(res, open(os.path.join(OUT, "conditioning_sweep.json"), "w"), indent=1)
    return res


def run_phases(n=16384):
    """Which phase of marginal_likelihood_gradient is sensitive to the INT8 GEMM?  SE 3-D at size n: gradient error vs the
    oracle with the INT8 path allowed in exactly the phases of the mask (1 potrf, 2 trtri, 4 lauum), guard off."""
    d = 3
    x, y, e = synth(2024 + n, n, d)
    theta = np.array([0.3, 0.1] + [np.log(0.35)] * d)
    lml_o, grad_o = orc.marginal_likelihood_gradient_blocked(x, y, ("SE",), "const", theta, e**2)
    res = {"what": f"SE d=3 N={n}: gradient error vs oracle per INT8 phase mask (1 potrf, 2 trtri, 4 lauum), guard off",
           "grad_oracle": grad_o.tolist(), "cases": []}
    m = gp.GpRegressor(x, y, y_err=e, hyperpars=theta)
    for mask in (0, 1, 2, 4, 3, 6, 7):
        with _lib.options(i8_grad_guard=0, i8_grad_phases=mask):
            lml, grad = m.marginal_likelihood_gradient(theta)
        entry = {"mask": mask, "grad_rel": rel(grad, grad_o), "grad_abs": float(np.abs(grad - grad_o).max()),
                 "lml_rel": abs(lml - lml_o) / abs(lml_o), "per_param_abs": np.abs(grad - grad_o).tolist()}
        print(json.dumps(entry), flush=True)
        res["cases"].append(entry)
    for max_k in (16384, 8192, 4096, 2048):     # lauum alone on the INT8 path, k extent chunked (per-chunk row scales)
        with _lib.options(i8_grad_guard=0, i8_grad_phases=4, gemm_i8_max_k=max_k):
            lml, grad = m.marginal_likelihood_gradient(theta)
            ms = m.engine.timers().get("lauum")
        entry = {"lauum_only_max_k": max_k, "grad_rel": rel(grad, grad_o), "grad_abs": float(np.abs(grad - grad_o).max()), "lauum_ms": ms}
        print(json.dumps(entry), flush=True)
        res["cases"].append(entry)
    with _lib.options(i8_grad_guard=1):
        lml, grad = m.marginal_likelihood_gradient(theta)
        res["guarded"] = {"grad_rel": rel(grad, grad_o), "retries": m.engine.stat("grad_guard_retries"), "est": m.engine.stat("grad_guard_est")}
        print(json.dumps(res["guarded"]), flush=True)
    json.dump(res, open(os.path.join(OUT, f"grad_phase_sensitivity_N{n}.json"), "w"), indent=1)


def run_predict(n=32768):
    """Which predict-solve variant is accurate on a DENSE data set (SE 2-D, the config-5 model: sigma^2 cancels hard)?
    mu / sigma at 4096 points vs the oracle for: the default dispatch, both diagonal-block variants, the full-width
    recursion, and the DMMA path."""
    d = 2
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 1, (n, d))
    y = np.sin(3 * x).sum(axis=1) + rng.normal(0, 0.05, n)
    e = np.full(n, 0.05)
    theta = np.array([0.2, 0.1] + [np.log(0.3)] * d)
    q = np.random.default_rng(6).uniform(0, 1, (4096, d))
    t0 = time.perf_counter()
    fit = orc.Fit(x, y, ("SE",), "const", theta, e**2)
    mu_o, sig_o = fit.predict(q)
    res = {"what": f"SE d=2 N={n}: predict variants vs oracle (mu max-norm relative, sigma elementwise relative)",
           "oracle_seconds": time.perf_counter() - t0, "sigma_over_amp_min": float(sig_o.min() / np.exp(theta[1])), "cases": {}}
    m = gp.GpRegressor(x, y, y_err=e, hyperpars=theta)
    for name, opts in (("default", {}), ("diag_recursion", {"predict_diag": 1}), ("diag_int8_inverse", {"predict_diag": 2}),
                       ("full_recursion", {"predict_block": -1}), ("dmma", {"gemm_i8": 0})):
        with _lib.options(**opts):
            m.set_hyperparameters(theta)
            mu, sig = m(np.tile(q, (8, 1)))          # 32768 rows: the blocked path engages
            mu, sig = mu[:4096], sig[:4096]
            res["cases"][name] = {"mu": rel(mu, mu_o), "sigma": float(np.abs(sig / sig_o - 1).max()), "alpha": rel(m.alpha, fit.alpha),
                                  "block": m.engine.stat("predict_block"), "trsm_ms": m.engine.timers().get("trsm")}
        print(name, json.dumps(res["cases"][name]), flush=True)
    json.dump(res, open(os.path.join(OUT, f"predict_variants_N{n}.json"), "w"), indent=1)


def run_cfg1():
    """BASELINE config 1 shape: SE 1-D, N=200, predict at 1000 points; latency of the hot calls (the reference needs
    12.9 ms per marginal_likelihood_gradient on the survey host)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "cfg1_demo.npz"))
    x, y, e, theta, q = g["x"], g["y"], g["y_err"], g["theta"], g["q"]
    m = gp.GpRegressor(x, y, y_err=e, hyperpars=theta)
    out = {}
    for name, fn in (("marginal_likelihood_gradient", lambda: m.marginal_likelihood_gradient(theta)),
                     ("marginal_likelihood", lambda: m.marginal_likelihood(theta)),
                     ("set_hyperparameters", lambda: m.set_hyperparameters(theta)),
                     ("predict_1000_points", lambda: m(q)),
                     ("gradient_1000_points", lambda: m.gradient(q))):
        for _ in range(5):
            fn()
        ts = []
        for _ in range(50):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        out[name + "_ms"] = {"median": 1e3 * float(np.median(ts)), "min": 1e3 * float(np.min(ts))}
    lml, grad = m.marginal_likelihood_gradient(theta)
    out["parity"] = {"lml": abs(lml - g["lml_from_grad"]) / abs(g["lml_from_grad"]), "grad": rel(grad, g["lml_grad"]),
                     "mu": rel(m(q)[0], g["pred_mu"])}
    out["reference_ms_survey_host"] = {"marginal_likelihood_gradient": 12.9}
    json.dump(out, open(os.path.join(OUT, "cfg1_latency.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "scale"
    if what == "scale":
        run_scale([int(a) for a in sys.argv[2:]] or [8192])
    elif what == "cond":
        run_cond(int(sys.argv[2]) if len(sys.argv) > 2 else 4096)
    elif what == "cfg1":
        run_cfg1()
    elif what == "predict":
        run_predict(int(sys.argv[2]) if len(sys.argv) > 2 else 32768)
    elif what == "phases":
        run_phases(int(sys.argv[2]) if len(sys.argv) > 2 else 16384)
