"""Quick GPU bring-up check of the engine against the numpy oracle and the golden fixtures (gpurun only)."""
import ctypes as C, glob, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from inference_tools_b200 import _lib
from oracle import gp_oracle as orc

lib = _lib.load_test_library()
dp = C.POINTER(C.c_double)
lib.gpb_test_potrf.argtypes = [C.c_int, dp, dp, C.POINTER(C.c_int), C.c_int, dp]
lib.gpb_test_inverse.argtypes = [C.c_int, dp, dp, dp, C.c_int, dp, dp]
P = lambda a: a.ctypes.data_as(dp)
rel = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-300, np.abs(np.asarray(b)).max()))
res = {}
rng = np.random.default_rng(0)

# ---- primitives
for n in (128, 256, 384, 1024, 2048):
    x = rng.uniform(0, 1, (n, 3))
    K = np.exp(-0.5 * ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1) / 0.09) + 0.01 * np.eye(n)
    A = K.copy(); dinv = np.zeros((n, 128)); info = C.c_int(-1); ms = C.c_double(0)
    rc = lib.gpb_test_potrf(n, P(A), P(dinv), C.byref(info), 1, C.byref(ms))
    assert rc == 0, lib.gpb_last_error()
    L = np.tril(A); Lref = np.linalg.cholesky(K)
    res[f"potrf_{n}"] = {"info": info.value, "rel": rel(L, Lref), "ms": ms.value}
    d0 = np.linalg.inv(Lref[:128, :128])
    res[f"potrf_{n}"]["dinv0_rel"] = rel(dinv[:128], d0)
    W = np.zeros((n, n)); Ki = np.zeros((n, n)); m = 256
    X = rng.standard_normal((m, n)); X0 = X.copy(); ms3 = np.zeros(3)
    rc = lib.gpb_test_inverse(n, P(K), P(W), P(Ki), m, P(X), P(ms3))
    assert rc == 0, lib.gpb_last_error()
    Wref = np.linalg.inv(Lref); Kiref = np.linalg.inv(K)
    res[f"inverse_{n}"] = {"W_rel": rel(np.tril(W), Wref), "W_upper": float(np.abs(np.triu(W, 1)).max()),
                           "Kinv_rel_lower": rel(np.tril(Ki), np.tril(Kiref)),
                           "trsm_rel": rel(X, np.linalg.solve(Lref, X0.T).T), "ms": ms3.tolist()}
# non-PD detection
n = 256
A = np.eye(n); A[200, 200] = -1.0; info = C.c_int(-1)
lib.gpb_test_potrf(n, P(A), None, C.byref(info), 1, None)
res["potrf_nonpd_info"] = info.value

print(json.dumps(res, indent=1)); sys.stdout.flush()
# ---- engine vs golden fixtures
KIND = {"SE": 0, "RQ": 1, "WHITE": 2, "HETERO": 3}; MEAN = {"const": 0, "linear": 1, "quadratic": 2}
gold = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "*.npz")))
for path in gold:
    name = os.path.basename(path)[:-4]
    if name.startswith("fit_"): continue
    g = np.load(path)
    comps = [str(c) for c in g["comps"]]; mean = str(g["mean"])
    x = g["x"] if g["x"].ndim == 2 else g["x"].reshape(-1, 1)
    e = _lib.Engine(0)
    e.set_data(x, g["y"], g["y_err"] ** 2 if g["y_err"].size else None)
    e.set_model([KIND[c] for c in comps], MEAN[mean])
    th = g["theta"]; r = {}
    info = e.factor(th); r["info"] = info
    r["alpha"] = rel(e.get(_lib.GET_ALPHA), g["alpha"])
    if "L_diag" in g.files: r["Ldiag"] = rel(np.diagonal(e.get(_lib.GET_L)), g["L_diag"])
    if "K_xx" in g.files:
        r["K_xx"] = rel(e.get(_lib.GET_K_XX), g["K_xx"]); r["L"] = rel(e.get(_lib.GET_L), g["L"])
        k, dk = e.covariance_and_gradients(th[e.n_mean:]); r["K_cov"] = rel(k, g["K_cov"]); r["dK"] = rel(dk, g["dK"])
        r["K_qx"] = rel(e.cross_covariance(g["q"].reshape(-1, x.shape[1]), x, th[e.n_mean:]), g["K_qx"])
    lml, info = e.lml(th); r["lml"] = abs(lml - g["lml"]) / abs(g["lml"])
    lml2, grad, info = e.lml_grad(th); r["lml2"] = abs(lml2 - g["lml_from_grad"]) / abs(g["lml_from_grad"])
    r["grad"] = float(np.abs(grad - g["lml_grad"]).max() / np.abs(g["lml_grad"]).max())
    q = g["q"].reshape(-1, x.shape[1])
    if "pred_mu" in g.files:
        mu, sig = e.predict(q); r["mu"] = rel(mu, g["pred_mu"]); r["sig"] = rel(sig, g["pred_sig"])
        r["sig_pt"] = float(np.abs(sig / g["pred_sig"] - 1).max())
        if "post_mu" in g.files: pm, pc = e.posterior(q[:16]); r["post_mu"] = rel(pm, g["post_mu"]); r["post_cov"] = rel(pc, g["post_cov"])
    if "grad_mean" in g.files:
        gm, gc = e.gradient(q); r["grad_mean"] = rel(gm.squeeze(), g["grad_mean"]); r["grad_cov"] = rel(gc.squeeze(), g["grad_cov"])
        dm, dv = e.spatial_derivatives(q); r["sd_dmu"] = rel(dm.squeeze(), g["sd_dmu"]); r["sd_dvar"] = rel(dv.squeeze(), g["sd_dvar"])
        if "ei_q" not in g.files:
            res[name] = r; print(name, r); e.close(); continue
        qq = g["ei_q"].reshape(-1, x.shape[1]); ymax = g["y"].max()
        ei, _, _ = e.expected_improvement(qq, ymax, 0); r["ei"] = float(np.abs(ei / g["ei"] - 1).max())
        nl, _, _ = e.expected_improvement(qq, ymax, 1); r["ei_nl"] = rel(nl, g["ei_optfunc"])
        nl2, gr, _ = e.expected_improvement(qq, ymax, 2); r["ei_nl2"] = rel(nl2, g["ei_optfunc_g_val"]); r["ei_grad"] = rel(gr, g["ei_optfunc_g_grad"])
        r["n_Zlt-3"] = int((g["ei_Z"] < -3).sum())
    res[name] = r
    print(name, r); sys.stdout.flush()
    e.close()

# ---- medium-size timing vs oracle
for (n, d, comps, M) in [(4096, 3, ("SE",), 4096), (8192, 3, ("SE",), 2048)]:
    x = rng.uniform(0, 1, (n, d)); y = np.sin(3 * x).sum(1) + rng.normal(0, 0.05, n); ye = np.full(n, 0.05)
    th = np.array([0.3, 0.1] + [np.log(0.3)] * d)
    e = _lib.Engine(0); e.set_data(x, y, ye ** 2); e.set_model([0], 0)
    q = rng.uniform(0, 1, (M, d))
    t0 = time.time(); e.factor(th); t1 = time.time(); tf = e.timers()
    t2 = time.time(); lml, grad, info = e.lml_grad(th); t3 = time.time(); tg = e.timers()
    t4 = time.time(); mu, sig = e.predict(q); t5 = time.time(); tp = e.timers()
    t6 = time.time(); mu, sig = e.predict(q); t7 = time.time()
    r = {"factor_s": t1 - t0, "lml_grad_s": t3 - t2, "predict_s": t5 - t4, "predict2_s": t7 - t6, "timers_factor": tf, "timers_grad": tg, "timers_pred": tp}
    if n <= 4096:
        t0 = time.time(); f = orc.Fit(x, y, ("SE",), "const", th, ye ** 2); r["oracle_fit_s"] = time.time() - t0
        r["alpha"] = rel(e.get(_lib.GET_ALPHA), f.alpha)
        mo, so = f.predict(q[:512]); r["mu"] = rel(mu[:512], mo); r["sig"] = float(np.abs(sig[:512] / so - 1).max())
        t0 = time.time(); lo, go = orc.marginal_likelihood_gradient(x, y, ("SE",), "const", th, ye ** 2); r["oracle_grad_s"] = time.time() - t0
        r["lml"] = abs(lml - lo) / abs(lo); r["grad"] = float(np.abs(grad - go).max() / np.abs(go).max())
    res[f"size_{n}"] = r
    e.close()
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/engine_check.json", "w"), indent=1)
