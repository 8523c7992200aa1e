"""Generate tests/golden/*.npz from the UNMODIFIED reference (/root/reference), imported live in the
build container with the stub recipe of SURVEY.md appendix B (skip inference/__init__.py, fake
matplotlib).  The reference cannot travel to the GPU box, so its outputs are committed as small
fixtures; this script is the provenance of every number in them.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

Cases are seeded; inputs are stored in the fixture beside the outputs so the tests never need to
re-derive them.
"""
import os
import sys
import types
import warnings

import numpy as np

REF = os.environ.get("GPB_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    pkg = types.ModuleType("inference")
    pkg.__path__ = [os.path.join(REF, "inference")]
    sys.modules["inference"] = pkg
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].__path__ = []
    import inference.gp as gp
    return gp


KINDS = {"SE": "SquaredExponential", "RQ": "RationalQuadratic", "WHITE": "WhiteNoise", "HETERO": "HeteroscedasticNoise"}
MEANS = {"const": "ConstantMean", "linear": "LinearMean", "quadratic": "QuadraticMean"}


def make_kernel(gp, comps):
    """comps: kinds, or ("CP", axis, (region kinds...)) for a ChangePoint (one kernel per region here)"""
    k = None
    for c in comps:
        if isinstance(c, str):
            inst = getattr(gp, KINDS[c])()
        else:
            _, axis, regions = c
            inst = gp.ChangePoint(kernels=[make_kernel(gp, r) for r in regions], axis=axis)
        k = inst if k is None else k + inst
    return k


def comps_repr(comps):
    """string form stored in the fixture, e.g. 'CP:0:SE|RQ,WHITE'"""
    out = []
    for c in comps:
        out.append(c if isinstance(c, str) else "CP:%d:%s" % (c[1], "|".join("+".join(r) for r in c[2])))
    return ",".join(out)


def synth(seed, n, d, sigma_n=0.05):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (n, d))
    y = np.sin(3 * x).sum(axis=1) + rng.normal(0, sigma_n, n)
    return x, y, np.full(n, sigma_n)


def default_theta(comps, mean, n, d, rng):
    """A well-conditioned evaluation point with a small seeded perturbation."""
    tm = {"const": [0.3], "linear": [0.3] + [0.2] * d, "quadratic": [0.3] + [0.2] * d + [-0.1] * d}[mean]
    tc = []
    flat = []
    for c in comps:
        if isinstance(c, str):
            flat.append(c)
        else:
            for r in c[2]:
                flat.extend(r)
            nreg = len(c[2])
            flat.extend(["CPP"] * (nreg - 1))
    cp_seen = 0
    for c in flat:
        if c == "CPP":
            cp_seen += 1
            tc += [cp_seen / (flat.count("CPP") + 1.0), 0.08]
        elif c == "SE":
            tc += [0.1] + [np.log(0.3)] * d
        elif c == "RQ":
            tc += [0.1, 1.0] + [np.log(0.3)] * d
        elif c == "WHITE":
            tc += [np.log(0.05)]
        elif c == "HETERO":
            tc += list(np.log(0.05) + 0.3 * rng.standard_normal(n))
    th = np.array(tm + tc, dtype=float)
    if "CPP" in flat:
        th[:len(tm)] += 0.05 * rng.standard_normal(len(tm))
        return th
    th[:len(tm) + (0 if "HETERO" in comps else len(tc))] += 0.05 * rng.standard_normal(len(tm) + (0 if "HETERO" in comps else len(tc)))
    return th


def case(gp, name, seed, n, d, comps, mean, m_query=64, with_err=True, store_k=False, loo=False, theta=None):
    rng = np.random.default_rng(seed + 1000)
    x, y, y_err = synth(seed, n, d)
    theta = default_theta(comps, mean, n, d, rng) if theta is None else np.array(theta, dtype=float)
    kw = dict(kernel=make_kernel(gp, comps), mean=getattr(gp, MEANS[mean])(), hyperpars=theta)
    if with_err:
        kw["y_err"] = y_err
    g = gp.GpRegressor(x, y, **kw)
    q = rng.uniform(-0.1, 1.1, (m_query, d))
    out = dict(x=x, y=y, y_err=y_err if with_err else np.zeros(0), theta=theta, q=q,
               comps=np.array(comps_repr(comps).split(",")) if any(not isinstance(c, str) for c in comps) else np.array(comps),
               mean=np.array(mean), labels=np.array(g.hyperpar_labels),
               bounds=np.array(g.hp_bounds, dtype=float), alpha=g.alpha, mu_train=g.mu,
               L_diag=np.diagonal(g.L).copy())
    if store_k:
        out["K_xx"] = g.K_xx
        out["L"] = g.L
        kk, grads = g.cov.covariance_and_gradients(g.cov_hyperpars)
        out["K_cov"] = kk
        out["dK"] = np.array(grads)
        out["K_qx"] = g.cov(q, x, g.cov_hyperpars)
    if not (d > 1 and "HETERO" in comps):       # Hetero.__call__ uses u.size (covariance.py:672): breaks for d>1
        mu, sig = g(q)
        out["pred_mu"], out["pred_sig"] = mu, sig
        pm, pc = g.build_posterior(q[:16])
        out["post_mu"], out["post_cov"] = pm, pc
    out["lml"] = np.float64(g.marginal_likelihood(theta))
    lml2, grad = g.marginal_likelihood_gradient(theta)
    out["lml_from_grad"], out["lml_grad"] = np.float64(lml2), grad
    if loo:
        out["loo"] = np.float64(g.loo_likelihood(theta))
        l2, lg = g.loo_likelihood_gradient(theta)
        out["loo_from_grad"], out["loo_grad"] = np.float64(l2), lg
        lm, ls = g.loo_predictions()
        out["loo_mu"], out["loo_sig"] = lm, ls
    if comps == ("SE",):
        gm, gc = g.gradient(q)
        out["grad_mean"], out["grad_cov"] = gm, gc
        dm, dv = g.spatial_derivatives(q)
        out["sd_dmu"], out["sd_dvar"] = dm, dv
        # acquisition: EI value, -ln EI and its gradient, one point at a time as the reference does
        ei = gp.ExpectedImprovement()
        ei.update_gp(g)
        qq = np.concatenate([q, rng.uniform(0.3, 0.7, (32, d))])     # interior points: tiny sigma => Z < -3
        out["ei_q"] = qq
        out["ei"] = np.array([ei(p) for p in qq])
        out["ei_optfunc"] = np.array([ei.opt_func(p) for p in qq])
        vg = [ei.opt_func_gradient(p) for p in qq]
        out["ei_optfunc_g_val"] = np.array([float(v[0]) for v in vg])
        out["ei_optfunc_g_grad"] = np.array([np.atleast_1d(v[1]) for v in vg])
        mu_q, sig_q = g(qq)
        out["ei_Z"] = (mu_q - g.y.max()) / sig_q
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: N={n} d={d} comps={comps} mean={mean} lml={out['lml']:.6f}")


def fit_case(gp, name, seed, n, d, comps, mean):
    """Full multistart fit from the reference with the legacy global RNG seeded (regression.py:591)."""
    x, y, y_err = synth(seed, n, d)
    np.random.seed(seed)
    g = gp.GpRegressor(x, y, y_err=y_err, kernel=make_kernel(gp, comps), mean=getattr(gp, MEANS[mean])())
    rng = np.random.default_rng(seed + 7)
    q = rng.uniform(0, 1, (64, d))
    mu, sig = g(q)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, y=y, y_err=y_err, comps=np.array(comps),
                        mean=np.array(mean), np_seed=seed, theta_opt=np.array(g.hyperpars),
                        lml_opt=np.float64(g.marginal_likelihood(g.hyperpars)), q=q, pred_mu=mu, pred_sig=sig,
                        bounds=np.array(g.hp_bounds, dtype=float))
    print(f"{name}: fit theta={np.array(g.hyperpars)} lml={g.marginal_likelihood(g.hyperpars):.6f}")


def demo_case(gp):
    """cfg 1: the documentation shape -- SE 1-D, N=200 on [-3,9], predict on linspace(-4,10,1000)."""
    rng = np.random.default_rng(1)
    n = 200
    x = np.sort(rng.uniform(-3, 9, n))
    f = 1 / (1 + np.exp(-x)) + 0.1 * np.sin(2 * x)
    y_err = np.full(n, 0.08)
    y = f + rng.normal(0, 0.08, n)
    theta = np.array([0.5, -0.3, 0.4])
    g = gp.GpRegressor(x, y, y_err=y_err, hyperpars=theta)
    q = np.linspace(-4, 10, 1000)
    mu, sig = g(q)
    gm, gc = g.gradient(q)
    dm, dv = g.spatial_derivatives(q)
    lml, grad = g.marginal_likelihood_gradient(theta)
    np.savez_compressed(os.path.join(HERE, "cfg1_demo.npz"), x=x, y=y, y_err=y_err, theta=theta, q=q, pred_mu=mu,
                        pred_sig=sig, grad_mean=gm, grad_cov=gc, sd_dmu=dm, sd_dvar=dv,
                        lml=np.float64(g.marginal_likelihood(theta)), lml_from_grad=np.float64(lml), lml_grad=grad,
                        alpha=g.alpha, bounds=np.array(g.hp_bounds, dtype=float), labels=np.array(g.hyperpar_labels),
                        comps=np.array(("SE",)), mean=np.array("const"))
    print(f"cfg1_demo: lml={lml:.6f}")


def inverter_data():
    """tests/gp/test_GpLinearInverter.py:41-64: gaussian-blur forward model of a sum of Lorentzians"""
    from scipy.special import erfc

    def normal_cdf(x, mu=0.0, sigma=1.0):
        return 0.5 * erfc(-(x - mu) / (np.sqrt(2) * sigma))

    def lorentzian(x, A, w, c):
        return A / (1 + ((x - c) / w) ** 2)

    n_data, n_basis = 32, 64
    x = np.linspace(-1, 1, n_basis)
    data_axis = np.linspace(-1, 1, n_data)
    dx = 0.5 * (x[1] - x[0])
    solution = lorentzian(x, 1.0, 0.1, 0.0) + lorentzian(x, 0.8, 0.15, 0.3) + lorentzian(x, 0.3, 0.1, -0.45)
    A = np.zeros([n_data, n_basis])
    for k in range(n_basis):
        A[:, k] = normal_cdf(data_axis + dx, mu=x[k], sigma=0.075) - normal_cdf(data_axis - dx, mu=x[k], sigma=0.075)
    rng = np.random.default_rng(123)
    y = A @ solution + rng.normal(size=n_data, scale=0.02)
    return x.reshape(-1, 1), y, np.zeros(n_data) + 0.02, A


def inverter_case(gp, name, comps, mean, seed):
    x, y, y_err, A = inverter_data()
    inv = gp.GpLinearInverter(model_matrix=A, y=y, y_err=y_err, parameter_spatial_positions=x,
                              prior_covariance_function=make_kernel(gp, comps), prior_mean_function=getattr(gp, MEANS[mean])())
    rng = np.random.default_rng(seed)
    thetas = rng.uniform(0.1, 1.0, size=(4, inv.n_hyperpars))
    out = dict(x=x, y=y, y_err=y_err, A=A, comps=np.array(comps), mean=np.array(mean), thetas=thetas,
               labels=np.array(inv.hyperpar_labels))
    out["lml"] = np.array([inv.marginal_likelihood(t) for t in thetas])
    lg = [inv.marginal_likelihood_gradient(t) for t in thetas]
    out["lml_from_grad"] = np.array([v[0] for v in lg])
    out["lml_grad"] = np.array([v[1] for v in lg])
    post = [inv.calculate_posterior(t) for t in thetas]
    out["post_mean"] = np.array([p[0] for p in post])
    out["post_cov"] = np.array([p[1] for p in post])
    out["post_mean_alt"] = np.array([inv.calculate_posterior_mean(t) for t in thetas])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: lml={out['lml']}")


def acquisition_case(gp, name, seed, n, d, comps, kappa):
    """UpperConfidenceBound / MaxVariance (acquisition.py:143-232), one point per call as the reference evaluates
    them: __call__, opt_func, opt_func_gradient (SquaredExponential only) and convergence_metric."""
    rng = np.random.default_rng(seed + 1000)
    x, y, y_err = synth(seed, n, d)
    theta = default_theta(comps, "const", n, d, rng)
    g = gp.GpRegressor(x, y, y_err=y_err, kernel=make_kernel(gp, comps), hyperpars=theta)
    q = np.concatenate([rng.uniform(-0.1, 1.1, (48, d)), rng.uniform(0.3, 0.7, (16, d))])
    out = dict(x=x, y=y, y_err=y_err, theta=theta, q=q, comps=np.array(comps), mean=np.array("const"), kappa=np.float64(kappa))
    ucb, mv = gp.UpperConfidenceBound(kappa=kappa), gp.MaxVariance()
    for tag, acq in (("ucb", ucb), ("mv", mv)):
        acq.update_gp(g)
        out[tag] = np.array([acq(p) for p in q])
        out[tag + "_optfunc"] = np.array([acq.opt_func(p) for p in q])
        out[tag + "_metric"] = np.array([acq.convergence_metric(p) for p in q])
        if comps == ("SE",):
            vg = [acq.opt_func_gradient(p) for p in q]
            out[tag + "_optfunc_g_val"] = np.array([float(np.squeeze(v[0])) for v in vg])
            out[tag + "_optfunc_g_grad"] = np.array([np.atleast_1d(v[1]) for v in vg])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: N={n} d={d} comps={comps} ucb[0]={out['ucb'][0]:.6f} mv[0]={out['mv'][0]:.6e}")


def acquisition_cases(gp):
    acquisition_case(gp, "acq_se_d2_n60", 61, 60, 2, ("SE",), 2.0)
    acquisition_case(gp, "acq_se_d1_n40", 62, 40, 1, ("SE",), 0.7)
    acquisition_case(gp, "acq_rqwhite_d3_n80", 63, 80, 3, ("RQ", "WHITE"), 3.5)


def main():
    warnings.simplefilter("ignore")
    gp = load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "acq":   # only the acquisition fixtures (added in round 2)
        return acquisition_cases(gp)
    acquisition_cases(gp)
    demo_case(gp)
    # small dense cases: K, L, dK stored
    case(gp, "se_d2_n32_const", 11, 32, 2, ("SE",), "const", store_k=True, loo=True)
    case(gp, "rq_d2_n32_linear", 12, 32, 2, ("RQ",), "linear", store_k=True, loo=True)
    case(gp, "rqwhite_d3_n48_quadratic", 13, 48, 3, ("RQ", "WHITE"), "quadratic", store_k=True)
    case(gp, "sewhite_d1_n40_const", 14, 40, 1, ("SE", "WHITE"), "const", store_k=True)
    case(gp, "sehetero_d1_n24_const", 15, 24, 1, ("SE", "HETERO"), "const", store_k=True)
    case(gp, "se_d1_n20_noerr", 16, 20, 1, ("SE",), "const", with_err=False, store_k=True, theta=[0.2, 0.1, np.log(0.02)])
    # medium cases: vectors only
    case(gp, "se_d3_n200_const", 21, 200, 3, ("SE",), "const")
    case(gp, "se_d5_n300_quadratic", 22, 300, 5, ("SE",), "quadratic")
    case(gp, "rqwhite_d5_n257_const", 23, 257, 5, ("RQ", "WHITE"), "const")
    case(gp, "rq_d1_n129_linear", 24, 129, 1, ("RQ",), "linear")
    case(gp, "se_d3_n1024_const", 25, 1024, 3, ("SE",), "const", m_query=256)
    case(gp, "rqwhite_d5_n1024_const", 26, 1024, 5, ("RQ", "WHITE"), "const", m_query=256)
    case(gp, "se_d2_n700_linear", 27, 700, 2, ("SE",), "linear", m_query=128)
    # ChangePoint kernels (covariance.py:371-605)
    case(gp, "cp_sese_d1_n40_const", 41, 40, 1, (("CP", 0, (("SE",), ("SE",))),), "const", store_k=True, loo=True)
    case(gp, "cp_serq_white_d2_n36_linear", 42, 36, 2, (("CP", 1, (("SE",), ("RQ",))), "WHITE"), "linear", store_k=True, loo=True)
    case(gp, "cp_sesese_d1_n48_const", 43, 48, 1, (("CP", 0, (("SE",), ("SE",), ("SE",))),), "const", store_k=True)
    case(gp, "cp_sese_d3_n300_const", 44, 300, 3, (("CP", 2, (("SE",), ("SE",))),), "const")
    # GpLinearInverter (inversion.py)
    inverter_case(gp, "linv_se_const", ("SE",), "const", 51)
    inverter_case(gp, "linv_rq_linear", ("RQ",), "linear", 52)
    inverter_case(gp, "linv_white_const", ("WHITE",), "const", 53)
    inverter_case(gp, "linv_rqse_const", ("RQ", "SE"), "const", 54)
    # multistart fits
    fit_case(gp, "fit_se_d1_n60", 31, 60, 1, ("SE",), "const")
    fit_case(gp, "fit_rqwhite_d2_n80", 32, 80, 2, ("RQ", "WHITE"), "const")


if __name__ == "__main__":
    main()
