"""GPU parity at the sizes where the default dispatch runs on the INT8 tensor-core GEMM (pytest -m gpu).

BASELINE.json config 2 size (N = 8192): alpha, mu / sigma at 256 points, the log marginal likelihood AND its gradient
against the CPU oracle at the north_star tolerance (1e-9 relative; the gradient norm-relative), with the default GEMM
dispatch and an assertion that the INT8 path carried the flops -- so the path the headline benchmark runs on is the path
that is checked.  The oracle side is the row-blocked restatement pinned in tests/test_oracle_golden.py.
tools/parity_at_scale.py repeats this at N = 16384 and 32768 (minutes of CPU; results under profiles/)."""
import numpy as np
import pytest
from numpy.linalg import LinAlgError

from conftest import rel_err, synth
import inference_tools_b200.gp as gp
from inference_tools_b200 import _lib
from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.mark.parametrize("d,comps", [(3, ("SE",)), (5, ("RQ", "WHITE"))])
def test_config2_size_against_oracle_on_the_int8_path(d, comps):
    n = 8192
    x, y, e = synth(4242 + d, n, d)
    if comps == ("SE",):
        kernel, theta = gp.SquaredExponential(), np.array([0.3, 0.1] + [np.log(0.35)] * d)
    else:
        kernel, theta = gp.RationalQuadratic() + gp.WhiteNoise(), np.array([0.2, 0.1, 1.0] + [np.log(0.3)] * d + [np.log(0.05)])
    q = np.random.default_rng(d).uniform(0, 1, (256, d))
    lib = _lib.load_library()
    f0, i0 = lib.gpb_gemm_flops(), lib.gpb_gemm_flops_int8()
    m = gp.GpRegressor(x, y, y_err=e, kernel=kernel, hyperpars=theta)
    lml_v = m.marginal_likelihood(theta)
    share = (lib.gpb_gemm_flops_int8() - i0) / (lib.gpb_gemm_flops() - f0)
    assert share > 0.8, f"INT8 path carried only {share:.2f} of the factorisation's GEMM flops"
    mu, sig = m(q)
    lml, grad = m.marginal_likelihood_gradient(theta)

    lml_o, grad_o = orc.marginal_likelihood_gradient_blocked(x, y, comps, "const", theta, e**2)
    ref = orc.Fit(x, y, comps, "const", theta, e**2)
    mu_o, sig_o = ref.predict(q)
    assert rel_err(m.alpha, ref.alpha) < TOL
    assert rel_err(mu, mu_o) < TOL and np.abs(sig / sig_o - 1).max() < TOL
    assert abs(lml - lml_o) <= TOL * abs(lml_o) and abs(lml_v - lml_o) <= TOL * abs(lml_o)
    assert np.abs(grad - grad_o).max() <= TOL * np.abs(grad_o).max()
    # and the two GEMM implementations agree with each other far inside the tolerance
    with _lib.options(gemm_i8=0):
        lml_d, grad_d = m.marginal_likelihood_gradient(theta)
    assert abs(lml - lml_d) <= 1e-11 * abs(lml_d) and np.abs(grad - grad_d).max() <= 1e-10 * np.abs(grad_d).max()


@pytest.mark.parametrize("sigma_n,ls", [(0.003, 0.3), (0.0003, 1.0)])
def test_ill_conditioned_kernels_int8_vs_dmma_vs_oracle(sigma_n, ls):
    """cond(K) ~ 1e7 .. 1e11 (near-noise-free smooth kernel, SURVEY.md 8d parity floor): the INT8 path must not report a
    spurious non-PD matrix, and must stay as close to the oracle as the DMMA path does (within a factor 10 or 1e-9)."""
    n, d = 4096, 3
    x, y, e = synth(7, n, d, sigma_n)
    theta = np.array([0.3, 0.1] + [np.log(ls)] * d)
    q = np.random.default_rng(1).uniform(0, 1, (128, d))
    lml_o, grad_o = orc.marginal_likelihood_gradient_blocked(x, y, ("SE",), "const", theta, e**2)
    ref = orc.Fit(x, y, ("SE",), "const", theta, e**2)
    mu_o, sig_o = ref.predict(q)
    err = {}
    for mode, opts in (("int8", {"gemm_i8": 2, "gemm_i8_min_k": 64}), ("dmma", {"gemm_i8": 0})):
        with _lib.options(**opts):
            m = gp.GpRegressor(x, y, y_err=e, hyperpars=theta)       # LinAlgError here = spurious info > 0
            lml, grad = m.marginal_likelihood_gradient(theta)
            mu, sig = m(q)
            err[mode] = {"lml": abs(lml - lml_o) / abs(lml_o), "grad": rel_err(grad, grad_o), "mu": rel_err(mu, mu_o),
                         "alpha": rel_err(m.alpha, ref.alpha)}
    for k in err["int8"]:
        assert err["int8"][k] <= max(10 * err["dmma"][k], TOL), (k, err)


def test_spurious_non_pd_on_the_int8_path_is_rechecked_on_dmma():
    """gpb_factor / gpb_lml(_grad): info > 0 from a factorisation whose GEMMs ran on the INT8 path is confirmed on the
    DMMA kernels before it is reported (option "i8_fallback"); a genuinely indefinite matrix still fails."""
    n, d = 2048, 2
    x, y, _ = synth(11, n, d, 0.01)
    x[1] = x[0]                                                         # duplicate point, no noise: singular up to jitter
    theta = np.array([0.0, 0.0, np.log(0.5), np.log(0.5)])
    with _lib.options(gemm_i8=2, gemm_i8_min_k=64):
        m_ok = None
        try:
            m_ok = gp.GpRegressor(x, y, hyperpars=theta)
        except LinAlgError:
            pass
    with _lib.options(gemm_i8=0):
        try:
            gp.GpRegressor(x, y, hyperpars=theta)
            dmma_ok = True
        except LinAlgError:
            dmma_ok = False
    assert (m_ok is not None) == dmma_ok                                 # same verdict as the FP64 tensor path


def test_blocked_predict_solve_on_cached_planes_matches_recursion_and_oracle():
    """The predict solve against the cached digit planes of L (api.cu: predict_solve_blocked; a-priori row scales for the
    solved rows) must agree with the recursive solve it replaces and with the oracle, for plain rows and for the stacked
    gradient rows (SquaredExponential), and must be the path that ran."""
    n, d, mq = 4096, 3, 40000
    x, y, e = synth(31, n, d)
    theta = np.array([0.3, 0.1] + [np.log(0.35)] * d)
    m = gp.GpRegressor(x, y, y_err=e, hyperpars=theta)
    q = np.random.default_rng(2).uniform(-0.05, 1.05, (mq, d))
    mu_b, sig_b = m(q)
    assert "planes" in m.engine.timers()                     # the blocked path built its planes
    with _lib.options(predict_diag=2):                       # the other diagonal-block variant gives the same answer
        mu_v, sig_v = m(q)
    assert rel_err(mu_v, mu_b) < 1e-12 and np.abs(sig_v / sig_b - 1).max() < 1e-10
    dm_b, dv_b = m.spatial_derivatives(q[:12000])
    with _lib.options(predict_block=-1):
        mu_r, sig_r = m(q)
        assert "planes" not in m.engine.timers()
        dm_r, dv_r = m.spatial_derivatives(q[:12000])
    assert rel_err(mu_b, mu_r) < 1e-12 and np.abs(sig_b / sig_r - 1).max() < 1e-10
    assert rel_err(dm_b, dm_r) < 1e-12 and rel_err(dv_b, dv_r) < 1e-10
    ref = orc.Fit(x, y, ("SE",), "const", theta, e**2)
    mu_o, sig_o = ref.predict(q[:256])
    assert rel_err(mu_b[:256], mu_o) < TOL and np.abs(sig_b[:256] / sig_o - 1).max() < TOL
    dm_o, dv_o = ref.spatial_derivatives(q[:64])
    assert rel_err(dm_b[:64], dm_o) < TOL and rel_err(dv_b[:64], dv_o) < TOL
    # a re-fit at other hyper-parameters must rebuild the planes (stale planes would reproduce the old predictions)
    theta2 = theta + np.array([0.0, 0.2, 0.1, -0.1, 0.05])
    m.set_hyperparameters(theta2)
    mu2, sig2 = m(q[:30000])
    ref2 = orc.Fit(x, y, ("SE",), "const", theta2, e**2)
    mu2_o, sig2_o = ref2.predict(q[:128])
    assert rel_err(mu2[:128], mu2_o) < TOL and np.abs(sig2[:128] / sig2_o - 1).max() < TOL


def test_gradient_with_the_substitution_inverse_matches_the_oracle():
    """Option "grad_inverse" = 1: K^-1 from the rows of L^-T (blocked substitution, potrf.cu: trtri_rows_lower) instead of
    the recursive triangular inverse -- same gradient, same alpha / LML, on the INT8 path and on DMMA."""
    n, d = 4096, 3
    x, y, e = synth(91, n, d)
    theta = np.array([0.3, 0.1] + [np.log(0.35)] * d)
    lml_o, grad_o = orc.marginal_likelihood_gradient_blocked(x, y, ("SE",), "const", theta, e**2)
    m = gp.GpRegressor(x, y, y_err=e, hyperpars=theta)
    lml0, grad0 = m.marginal_likelihood_gradient(theta)
    for opts in ({"grad_inverse": 1}, {"grad_inverse": 1, "gemm_i8": 0}, {"grad_inverse": 1, "gemm_i8": 2, "gemm_i8_min_k": 64}):
        with _lib.options(**opts):
            lml, grad = m.marginal_likelihood_gradient(theta)
        assert abs(lml - lml_o) <= TOL * abs(lml_o)
        assert np.abs(grad - grad_o).max() <= TOL * np.abs(grad_o).max()
        assert np.abs(grad - grad0).max() <= 1e-10 * np.abs(grad0).max()
