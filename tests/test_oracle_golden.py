"""CPU: the numpy oracle (oracle/gp_oracle.py) against the fixtures generated from the unmodified
reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, rel_err
from oracle import gp_oracle as orc

TOL = 1e-10


@pytest.mark.parametrize("name", golden_names())
def test_fit_and_predict_match_reference(name):
    g = load_golden(name)
    f = orc.Fit(g["x"], g["y"], g["comps"], g["mean"], g["theta"], g["noise_var"])
    assert rel_err(f.alpha, g["alpha"]) < TOL
    if "L_diag" in g:
        assert rel_err(np.diagonal(f.L), g["L_diag"]) < TOL
    if "L" in g:
        assert rel_err(f.L, g["L"]) < TOL
    if "pred_mu" in g:
        mu, sig = f.predict(g["q"])
        assert rel_err(mu, g["pred_mu"]) < TOL
        assert np.abs(sig / g["pred_sig"] - 1).max() < 1e-8
    if "post_mu" in g:
        pm, pc = f.posterior(g["q"][:16])
        assert rel_err(pm, g["post_mu"]) < TOL and rel_err(pc, g["post_cov"]) < 1e-9


@pytest.mark.parametrize("name", golden_names())
def test_marginal_likelihood_and_gradient_match_reference(name):
    g = load_golden(name)
    args = (g["x"], g["y"], g["comps"], g["mean"], g["theta"], g["noise_var"])
    lml = orc.marginal_likelihood(*args)
    assert abs(lml - g["lml"]) <= TOL * abs(g["lml"])
    lml2, grad = orc.marginal_likelihood_gradient(*args)
    assert abs(lml2 - g["lml_from_grad"]) <= TOL * abs(g["lml_from_grad"])
    assert np.abs(grad - g["lml_grad"]).max() <= 1e-9 * np.abs(g["lml_grad"]).max()


@pytest.mark.parametrize("name", [n for n in golden_names() if "K_xx" in load_golden(n)])
def test_covariance_matrices_match_reference(name):
    g = load_golden(name)
    n, d = g["x"].shape
    tm, parts = orc.split_theta(g["theta"], g["comps"], g["mean"], n, d)
    assert rel_err(orc.train_cov(g["comps"], parts, g["x"], g["noise_var"]), g["K_xx"]) < 1e-13
    k, grads = orc.cov_and_grads(g["comps"], parts, g["x"])
    assert rel_err(k, g["K_cov"]) < 1e-13
    assert rel_err(np.array(grads), g["dK"]) < 1e-12
    assert rel_err(orc.cross_cov(g["comps"], parts, g["q"], g["x"]), g["K_qx"]) < 1e-13


@pytest.mark.parametrize("name", [n for n in golden_names() if "grad_mean" in load_golden(n)])
def test_gradient_spatial_derivatives_and_ei_match_reference(name):
    g = load_golden(name)
    f = orc.Fit(g["x"], g["y"], g["comps"], g["mean"], g["theta"], g["noise_var"])
    gm, gc = f.gradient(g["q"])
    assert rel_err(gm, g["grad_mean"]) < 1e-9 and rel_err(gc, g["grad_cov"]) < 1e-9
    dm, dv = f.spatial_derivatives(g["q"])
    assert rel_err(dm, g["sd_dmu"]) < 1e-9 and rel_err(dv, g["sd_dvar"]) < 1e-9
    if "ei_q" not in g:
        return
    qq = g["ei_q"].reshape(-1, g["x"].shape[1])
    mu, sig = f.predict(qq)
    ymax = g["y"].max()
    assert np.allclose(orc.expected_improvement(mu, sig, ymax), g["ei"], rtol=1e-7, atol=1e-300)
    assert rel_err(orc.neg_log_ei(mu, sig, ymax), g["ei_optfunc"]) < 1e-9
    dmu, dvar = f.spatial_derivatives(qq)
    val, grad = orc.neg_log_ei_gradient(mu, sig, dmu, dvar, ymax)
    assert rel_err(val, g["ei_optfunc_g_val"]) < 1e-9
    assert rel_err(grad, g["ei_optfunc_g_grad"].reshape(grad.shape)) < 1e-8


@pytest.mark.parametrize("name", [n for n in golden_names() if "loo" in load_golden(n)])
def test_loo_matches_reference(name):
    g = load_golden(name)
    args = (g["x"], g["y"], g["comps"], g["mean"], g["theta"], g["noise_var"])
    assert abs(orc.loo_likelihood(*args) - g["loo"]) <= TOL * abs(g["loo"])
    val, grad = orc.loo_likelihood_gradient(*args)
    assert abs(val - g["loo_from_grad"]) <= TOL * abs(g["loo_from_grad"])
    assert np.abs(grad - g["loo_grad"]).max() <= 1e-8 * np.abs(g["loo_grad"]).max()
    f = orc.Fit(*args[:5], g["noise_var"])
    lm, ls = f.loo_predictions()
    assert rel_err(lm, g["loo_mu"]) < 1e-9 and rel_err(ls, g["loo_sig"]) < 1e-9


@pytest.mark.parametrize("name", golden_names())
def test_bounds_match_reference(name):
    g = load_golden(name)
    b = orc.mean_bounds(g["mean"], g["x"], g["y"]) + orc.cov_bounds(g["comps"], g["x"], g["y"])
    assert np.allclose(np.array(b, dtype=float), g["bounds"], rtol=1e-12, atol=1e-12)


def test_non_positive_definite_returns_sentinel():
    x = np.linspace(0, 1, 12)[:, None]
    x[5] = x[4]                               # duplicate point, no noise, huge length-scale => singular
    y = np.sin(3 * x[:, 0])
    val = orc.marginal_likelihood(x, y, ("SE",), "const", np.array([0.0, 0.0, 5.0]))
    assert val == -1e50 or np.isfinite(val)


def test_ei_fixtures_cover_both_branches():
    lo = hi = 0
    for n in golden_names():
        g = load_golden(n)
        if "ei_Z" in g:
            lo += int((g["ei_Z"] < -3).sum())
            hi += int((g["ei_Z"] >= -3).sum())
    assert lo > 50 and hi > 50


LINV = ["linv_se_const", "linv_rq_linear", "linv_white_const", "linv_rqse_const"]


@pytest.mark.parametrize("name", LINV)
def test_linear_inverter_matches_reference(name):
    import os
    from conftest import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    comps = tuple(str(c) for c in g["comps"])
    inv = orc.LinearInverter(g["y"], g["y_err"], g["A"], g["x"], comps, str(g["mean"]))
    for i, th in enumerate(g["thetas"]):
        assert abs(inv.marginal_likelihood(th) - g["lml"][i]) <= 1e-10 * abs(g["lml"][i])
        lml, grad = inv.marginal_likelihood_gradient(th)
        assert abs(lml - g["lml_from_grad"][i]) <= 1e-10 * abs(g["lml_from_grad"][i])
        assert np.abs(grad - g["lml_grad"][i]).max() <= 1e-9 * np.abs(g["lml_grad"][i]).max()
        mu, cov = inv.calculate_posterior(th)
        assert rel_err(mu, g["post_mean"][i]) < 1e-9 and rel_err(cov, g["post_cov"][i]) < 1e-9


@pytest.mark.parametrize("name", ["acq_se_d2_n60", "acq_se_d1_n40", "acq_rqwhite_d3_n80"])
def test_ucb_and_max_variance_match_reference(name):
    """UpperConfidenceBound / MaxVariance (acquisition.py:143-232) restated in the oracle against reference outputs."""
    g = load_golden(name)
    f = orc.Fit(g["x"], g["y"], g["comps"], g["mean"], g["theta"], g["noise_var"])
    kappa = float(g["kappa"])
    mu, sig = f.predict(g["q"])
    assert rel_err(orc.upper_confidence_bound(mu, sig, kappa), g["ucb"]) < 1e-9
    assert rel_err(-orc.upper_confidence_bound(mu, sig, kappa), g["ucb_optfunc"]) < 1e-9
    assert rel_err(orc.upper_confidence_bound(mu, sig, kappa) - g["y"].max(), g["ucb_metric"]) < 1e-8
    assert np.abs(orc.max_variance(sig) / g["mv"] - 1).max() < 1e-8
    assert np.abs(np.sqrt(orc.max_variance(sig)) / g["mv_metric"] - 1).max() < 1e-8
    if "ucb_optfunc_g_grad" in g:
        dm, dv = f.spatial_derivatives(g["q"])
        val, grad = orc.upper_confidence_bound_gradient(mu, sig, dm, dv, kappa)
        assert rel_err(val, g["ucb_optfunc_g_val"]) < 1e-9
        assert rel_err(grad, g["ucb_optfunc_g_grad"].reshape(grad.shape)) < 1e-8
        assert rel_err(-np.asarray(dv).reshape(grad.shape), g["mv_optfunc_g_grad"].reshape(grad.shape)) < 1e-8
        assert np.abs(-orc.max_variance(sig) / g["mv_optfunc_g_val"] - 1).max() < 1e-8


@pytest.mark.parametrize("name", ["se_d3_n200_const", "rqwhite_d5_n257_const", "rqwhite_d3_n48_quadratic", "sewhite_d1_n40_const"])
def test_row_blocked_gradient_matches_dense_oracle_and_reference(name):
    """The memory-lean gradient used by tools/parity_at_scale.py (N = 16384 / 32768) is the same arithmetic."""
    g = load_golden(name)
    args = (g["x"], g["y"], g["comps"], g["mean"], g["theta"], g["noise_var"])
    lml, grad = orc.marginal_likelihood_gradient_blocked(*args, block=37)
    lml_d, grad_d = orc.marginal_likelihood_gradient(*args)
    assert abs(lml - lml_d) <= 1e-13 * abs(lml_d) and np.abs(grad - grad_d).max() <= 1e-12 * np.abs(grad_d).max()
    assert abs(lml - g["lml_from_grad"]) <= TOL * abs(g["lml_from_grad"])
    assert np.abs(grad - g["lml_grad"]).max() <= 1e-9 * np.abs(g["lml_grad"]).max()
