"""CPU baseline leg of bench.py -- TEST/BENCH INFRASTRUCTURE (never imported by the product).

Times the reference's algorithm for one fit + LML-gradient + predict step (the numpy/scipy calls of
regression.py:551-566, 239-244, 208-216 restated in oracle/gp_oracle.py) on a BOUNDED sample of the
benchmark workload, component by component, and extrapolates each component to the full configuration
with its own complexity exponent (assembly and traces ~ N^2, LAPACK/BLAS ~ N^3, predict ~ M N^2).  The
unmodified reference cannot run the full configuration at all: it stores two (N,N,d) float64 arrays
(covariance.py:315-316), 2 x 43 GB at N=32768, d=5."""
from __future__ import annotations

import os
import time

import numpy as np
from numpy.linalg import cholesky
from scipy.linalg import solve_triangular

from oracle import gp_oracle as orc


def synth(seed, n, d, sigma_n=0.05):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (n, d))
    y = np.sin(3 * x).sum(axis=1) + rng.normal(0, sigma_n, n)
    return x, y, np.full(n, sigma_n)


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def timed_step(n_s, d, comps, mean, theta, m_s, seed=0):
    """One reference-algorithm step at sample size (n_s, m_s); returns component wall times in seconds."""
    x, y, e = synth(seed, n_s, d)
    nv = e**2
    q = np.random.default_rng(seed + 1).uniform(0, 1, (m_s, d))
    tm, parts = orc.split_theta(theta, comps, mean, n_s, d)
    t = {}
    # ---- marginal_likelihood_gradient (regression.py:551-566)
    t0 = time.perf_counter()
    k, grad_k = orc.cov_and_grads(comps, parts, x)
    k = k + np.diag(nv)
    t["grad_assemble"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    L = cholesky(k)
    ik = solve_triangular(L, np.eye(n_s), lower=True)
    ik = ik.T @ ik
    t["grad_lapack"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    mu = orc.mean_vec(mean, tm, x, x.mean(axis=0))
    alpha = ik @ (y - mu)
    lml = -0.5 * ((y - mu) @ alpha) - np.log(np.diagonal(L)).sum()
    Q = alpha[:, None] * alpha[None, :] - ik
    grad = [0.5 * (Q * g.T).sum() for g in grad_k]
    t["grad_trace"] = time.perf_counter() - t0
    del k, grad_k, ik, Q, L
    # ---- set_hyperparameters (regression.py:239-244)
    t0 = time.perf_counter()
    k = orc.train_cov(comps, parts, x, nv)
    t["fit_assemble"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    L = cholesky(k)
    alpha = solve_triangular(L.T, solve_triangular(L, y - mu, lower=True))
    t["fit_lapack"] = time.perf_counter() - t0
    # ---- __call__ (regression.py:208-216): one dtrtrs per query point, as the reference loops
    kqq = orc.prior_var(comps, parts)
    t0 = time.perf_counter()
    out = []
    for p in q[:, None, :]:
        kqx = orc.cross_cov(comps, parts, p, x)
        m_ = (kqx @ alpha)[0] + tm[0]
        v = solve_triangular(L, kqx.T, lower=True)
        out.append((m_, kqq - (v**2).sum()))
    t["predict_per_point"] = (time.perf_counter() - t0) / max(1, m_s)
    t["_check"] = float(lml) + float(np.sum(grad)) + float(out[-1][0])
    return t


def extrapolate(t, n_s, n, m_total):
    """Seconds of the full step at (n, m_total) from the sample timings, per-component exponents."""
    r = n / n_s
    fit_grad = (t["grad_assemble"] + t["grad_trace"] + t["fit_assemble"]) * r**2 + (t["grad_lapack"] + t["fit_lapack"]) * r**3
    predict = t["predict_per_point"] * r**2 * m_total
    return fit_grad + predict, fit_grad, predict


# ----------------------------------------------------------------------------------------------------------------------
# The UNMODIFIED reference (baseline/_ref, loaded by oracle/ref_loader.py) on the same step, at sizes it can allocate.
def reference_timed_step(n_s, d, theta, m_s, seed=0):
    """marginal_likelihood_gradient + set_hyperparameters + __call__ at m_s points of the reference's own GpRegressor
    (RationalQuadratic + WhiteNoise, the benchmark's model) at sample size n_s; wall seconds per call, or None when no copy
    of the reference is present on this host.  Construction (pass_spatial_data, bounds) is outside the step, as it is for
    the engine."""
    from oracle.ref_loader import load_reference_gp

    ref = load_reference_gp()
    if ref is None:
        return None
    x, y, e = synth(seed, n_s, d)
    q = np.random.default_rng(seed + 1).uniform(0, 1, (m_s, d))
    t0 = time.perf_counter()
    g = ref.GpRegressor(x, y, y_err=e, kernel=ref.RationalQuadratic() + ref.WhiteNoise(), hyperpars=theta)
    t = {"construct": time.perf_counter() - t0}
    t0 = time.perf_counter()
    lml, grad = g.marginal_likelihood_gradient(theta)
    t["grad"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    g.set_hyperparameters(theta)
    t["fit"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    mu, sig = g(q)
    t["predict_per_point"] = (time.perf_counter() - t0) / max(1, m_s)
    t["_check"] = float(lml) + float(np.sum(grad)) + float(mu[-1])
    return t


def reference_extrapolate(t1, n1, t2, n2, n, m_total):
    """Full-step seconds at (n, m_total) from two sample sizes: every call is modelled as a N^2 + b N^3 (assembly, traces
    and per-point solves are N^2; dpotrf, the explicit inverse and iK.T @ iK are N^3) with a, b >= 0 fitted through the two
    measurements; the per-point predict cost is pure N^2 (one dtrtrs against L per point, regression.py:213)."""
    out = {}
    for key in ("grad", "fit"):
        a1, a2 = t1[key], t2[key]
        # a n1^2 + b n1^3 = a1 ; a n2^2 + b n2^3 = a2
        det = n1**2 * n2**3 - n2**2 * n1**3
        a = (a1 * n2**3 - a2 * n1**3) / det
        b = (n1**2 * a2 - n2**2 * a1) / det
        if a < 0 or b < 0:       # degenerate fit: fall back to the pessimistic-for-us pure N^2 / optimistic pure N^3 mix
            a, b = 0.0, a2 / n2**3
        out[key] = a * n**2 + b * n**3
    out["predict"] = t2["predict_per_point"] * (n / n2) ** 2 * m_total
    out["step"] = out["grad"] + out["fit"] + out["predict"]
    return out
