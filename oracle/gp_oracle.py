"""CPU oracle for the GpRegressor hot path -- TEST INFRASTRUCTURE, never imported by the product.

A functional numpy/scipy restatement of the arithmetic the reference performs on this path.  Every
function cites the reference lines it follows (paths relative to /root/reference/inference/gp/).
The third-party arithmetic (LAPACK dpotrf / dtrtrs through numpy.linalg.cholesky and
scipy.linalg.solve_triangular, BLAS through `@`) is called exactly as the reference calls it; the
difference to the reference is only structural: no (N,N,d) arrays are stored, covariance blocks are
built on demand from x, so the same code runs at N = 16k-32k where the reference cannot allocate.

Parity pinning: the reference holds no golden vectors for this path.  This
restatement is pinned instead against outputs of the unmodified reference imported in the build
container: tests/golden/make_golden.py generates tests/golden/*.npz from the live reference and
tests/test_oracle_golden.py checks this module against those fixtures.

Model description used throughout:
    comps : sequence of terms; a term is a kernel kind "SE" | "RQ" | "WHITE" | "HETERO", or a change-point term
            ("CP", axis, (region_0, region_1, ...)) whose regions are sequences of kernel kinds (covariance.py:371-605)
    mean  : "const" | "linear" | "quadratic"
    theta : [mean params | cov params] (regression.py:150-155), cov params concatenated in
            component order (covariance.py:61, 692-697)
"""
from __future__ import annotations

import numpy as np
from numpy.linalg import cholesky
from scipy.linalg import solve_triangular
from scipy.special import erf, erfcx

EPS_JITTER = 1e-12  # covariance.py:221, 318


# ----------------------------------------------------------------------------- parameter layout
def n_cov_params(kind: str, n: int, d: int) -> int:
    """covariance.py:222 (SE d+1), :319 (RQ d+2), :140 (White 1), :654 (Hetero N)."""
    return {"SE": d + 1, "RQ": d + 2, "WHITE": 1, "HETERO": n}[kind]


def n_mean_params(mean: str, d: int) -> int:
    """mean.py:34 (1), :62 (1+d), :95 (1+2d)."""
    return {"const": 1, "linear": 1 + d, "quadratic": 1 + 2 * d}[mean]


class Leaves(list):
    """Flattened covariance description: list of (kind, theta_slice_values, region_index_or_None, theta_offset) plus the
    change-point data `cp` = (axis, locations, widths, theta_offset_of_first_cp_param) or None.  Parameter order follows
    the reference: per term in sum order (covariance.py:61); inside a ChangePoint the region kernels first, then
    (location, width) per change-point (covariance.py:484-489)."""
    cp = None
    n_theta = 0


def flatten(comps, tc, n, d):
    tc = np.asarray(tc, dtype=float)
    out = Leaves()
    o = 0
    for term in comps:
        if isinstance(term, str):
            p = n_cov_params(term, n, d)
            out.append((term, tc[o:o + p], None, o))
            o += p
        else:
            _, axis, regions = term
            assert out.cp is None, "one ChangePoint per model"
            for r, region in enumerate(regions):
                for kind in region:
                    p = n_cov_params(kind, n, d)
                    out.append((kind, tc[o:o + p], r, o))
                    o += p
            ncp = len(regions) - 1
            cpp = tc[o:o + 2 * ncp].reshape(ncp, 2)
            out.cp = (axis, cpp[:, 0].copy(), cpp[:, 1].copy(), o, len(regions))
            o += 2 * ncp
    out.n_theta = o
    assert o == tc.size, "wrong number of hyper-parameters"
    return out


def split_theta(theta, comps, mean, n, d):
    theta = np.asarray(theta, dtype=float)
    pm = n_mean_params(mean, d)
    return theta[:pm], flatten(comps, theta[pm:], n, d)


def _logistic(xa, loc, width):
    """covariance.py:592-595"""
    return 1.0 / (1.0 + np.exp(-(xa - loc) / width))


def region_weights(leaves, pts):
    """g_r(x) with coeff_r(u, v) = g_r(u) g_r(v): g_0 = 1 - f_0, g_r = f_{r-1} (1 - f_r), g_last = f_last
    (covariance.py:520-531: kernel_coeffs[-1] *= w1; kernel_coeffs.append(w2))."""
    axis, locs, widths, _, nreg = leaves.cp
    f = [_logistic(pts[:, axis], locs[a], widths[a]) for a in range(nreg - 1)]
    g = []
    for r in range(nreg):
        w = np.ones(pts.shape[0])
        if r > 0:
            w = w * f[r - 1]
        if r < nreg - 1:
            w = w * (1 - f[r])
        g.append(w)
    return g


# ----------------------------------------------------------------------------- covariance functions
def _scaled_half_sqdist(u, v, ls):
    """Z_ij = sum_k 0.5 (u_ik - v_jk)^2 / l_k^2, accumulated dimension by dimension in the same
    order as numpy's `.sum(axis=2)` over the last axis (covariance.py:243-244, 339-340)."""
    z = np.zeros((u.shape[0], v.shape[0]))
    for k in range(u.shape[1]):
        dk = u[:, k, None] - v[None, :, k]
        z += (0.5 * dk**2) / ls[k] ** 2
    return z


def cross_cov(comps, parts, u, v):
    """cov(u, v, theta): covariance.py:240-245 (SE), :335-341 (RQ); noise kernels return zeros
    (:160-161, :671-672); composite = sum (:86-89)."""
    out = np.zeros((u.shape[0], v.shape[0]))
    if parts.cp is not None:
        gu, gv = region_weights(parts, u), region_weights(parts, v)
    for kind, th, reg, _ in parts:
        if kind == "SE":
            a, ls = np.exp(th[0]), np.exp(th[1:])
            k = a**2 * np.exp(-_scaled_half_sqdist(u, v, ls))
        elif kind == "RQ":
            a, q, ls = np.exp(th[0]), np.exp(th[1]), np.exp(th[2:])
            z = _scaled_half_sqdist(u, v, ls)
            k = a**2 * (1 + z / q) ** (-q)
        else:
            continue
        out += k if reg is None else gu[reg][:, None] * k * gv[reg][None, :]
    return out


def prior_var(comps, parts, q=None):
    """k(q,q): a^2 summed over the smooth components (noise kernels contribute 0), times the squared region weight
    of the query point under a ChangePoint."""
    if parts.cp is None:
        return sum(np.exp(th[0]) ** 2 for kind, th, _, _ in parts if kind in ("SE", "RQ"))
    g = region_weights(parts, q)
    return sum(np.exp(th[0]) ** 2 * (1.0 if reg is None else g[reg] ** 2)
               for kind, th, reg, _ in parts if kind in ("SE", "RQ"))


def train_cov_rows(comps, parts, x, r0, r1, noise_var=None):
    """Rows [r0,r1) of K(theta)+sig on the training data: build_covariance (covariance.py:247-255,
    343-348, 163-169, 674-680, 91-95) plus diag(y_err^2) (regression.py:239, 320).  The 1e-12
    jitter is added before the a^2 scaling (covariance.py:254-255, 348)."""
    n = x.shape[0]
    blk = cross_cov(comps, parts, x[r0:r1], x)
    idx = np.arange(r0, r1)
    g = region_weights(parts, x[r0:r1]) if parts.cp is not None else None
    for kind, th, reg, _ in parts:
        w2 = 1.0 if reg is None else g[reg] ** 2
        if kind in ("SE", "RQ"):
            blk[idx - r0, idx] += w2 * np.exp(th[0]) ** 2 * EPS_JITTER
        elif kind == "WHITE":
            blk[idx - r0, idx] += w2 * np.exp(2 * th[0])
        elif kind == "HETERO":
            blk[idx - r0, idx] += w2 * np.exp(2 * th[r0:r1])
    if noise_var is not None:
        blk[idx - r0, idx] += noise_var[r0:r1]
    assert blk.shape == (r1 - r0, n)
    return blk


def train_cov(comps, parts, x, noise_var=None, y_cov=None, block=2048):
    n = x.shape[0]
    k = np.empty((n, n))
    for r0 in range(0, n, block):
        r1 = min(n, r0 + block)
        k[r0:r1] = train_cov_rows(comps, parts, x, r0, r1, noise_var)
    if y_cov is not None:
        k += y_cov
    return k


def _leaf_cov_and_grads(kind, th, x):
    """One plain kernel on the training data: (K, [dK/dtheta]) (covariance.py:268-276, 350-365, 171-175, 682-686)."""
    n, d = x.shape
    eye = np.eye(n)
    grads = []
    if kind == "SE":
        a, ls = np.exp(th[0]), np.exp(th[1:])
        k = a**2 * (np.exp(-_scaled_half_sqdist(x, x, ls)) + EPS_JITTER * eye)
        grads.append(2.0 * k)
        for i in range(d):
            dx2 = (x[:, i, None] - x[None, :, i]) ** 2
            grads.append((dx2 / ls[i] ** 2) * k)
    elif kind == "RQ":
        a, q, ls = np.exp(th[0]), np.exp(th[1]), np.exp(th[2:])
        z = _scaled_half_sqdist(x, x, ls)
        f = 1 + z / q
        lnf = np.log(f)
        k = a**2 * (np.exp(-q * lnf) + EPS_JITTER * eye)
        grads.append(2.0 * k)
        grads.append(-k * (lnf * q - z / f))
        g = 2 * k / f
        for i in range(d):
            grads.append(g * (0.5 * (x[:, i, None] - x[None, :, i]) ** 2 / ls[i] ** 2))
    elif kind == "WHITE":
        k = np.exp(2 * th[0]) * eye
        grads.append(2.0 * k)
    else:
        s2 = np.exp(2 * th)
        k = np.diag(s2)
        for i in range(n):
            g = np.zeros((n, n))
            g[i, i] = 2.0 * s2[i]
            grads.append(g)
    return k, grads


def cov_and_grads(comps, parts, x):
    """covariance_and_gradients WITHOUT sig, gradients in theta order: plain kernels (covariance.py:268-276, 350-365,
    171-175, 682-686), sums (:97-105), ChangePoint (:541-584; the change-point gradients use the UNWEIGHTED region
    covariances exactly as the reference does, which is exact for two regions).  Dense (small-N use only)."""
    n, d = x.shape
    ktot = np.zeros((n, n))
    grads = []
    g = region_weights(parts, x) if parts.cp is not None else None
    region_k = {}
    pending_cp = parts.cp is not None
    for idx, (kind, th, reg, _) in enumerate(parts):
        k, gk = _leaf_cov_and_grads(kind, th, x)
        if reg is None:
            if pending_cp and region_k:        # first term after the ChangePoint block: emit its own gradients first
                grads.extend(_cp_grads(parts, x, region_k))
                pending_cp = False
            ktot = ktot + k
            grads.extend(gk)
        else:
            coeff = g[reg][:, None] * g[reg][None, :]
            ktot = ktot + k * coeff
            grads.extend([m * coeff for m in gk])
            region_k[reg] = region_k.get(reg, 0) + k
    if pending_cp and region_k:
        grads.extend(_cp_grads(parts, x, region_k))
    return ktot, grads


def _cp_grads(parts, x, region_k):
    """covariance.py:574-583 with logistic_and_gradient (:597-602)."""
    axis, locs, widths, _, nreg = parts.cp
    out = []
    for a in range(nreg - 1):
        z = (x[:, axis] - locs[a]) / widths[a]
        w = 1.0 / (1.0 + np.exp(-z))
        dfdc = -w * (1 - w) / widths[a]
        for dw in (dfdc, dfdc * z):
            A = -dw[:, None] * (1 - w)[None, :]
            B = dw[:, None] * w[None, :]
            out.append(region_k[a] * (A + A.T) + region_k[a + 1] * (B + B.T))
    return out


# ----------------------------------------------------------------------------- mean functions
def mean_vec(mean, tm, x, xbar=None):
    """build_mean: mean.py:47-48, :77-78, :117-120.  xbar = column mean of the TRAINING x."""
    if mean == "const":
        return np.zeros(x.shape[0]) + tm[0]
    d = x.shape[1]
    dx = x - xbar[None, :]
    if mean == "linear":
        return tm[0] + dx @ tm[1:]
    return tm[0] + dx @ tm[1:d + 1] + (dx**2) @ tm[d + 1:2 * d + 1]


def mean_grads(mean, x, xbar=None):
    """mean_and_gradients: mean.py:50-51, :80-83, :122-126."""
    n, d = x.shape
    g = [np.ones(n)]
    if mean in ("linear", "quadratic"):
        dx = x - xbar[None, :]
        g.extend(dx.T)
        if mean == "quadratic":
            g.extend((dx**2).T)
    return g


# ----------------------------------------------------------------------------- fit / likelihood
class Fit:
    """State produced by set_hyperparameters (regression.py:218-244)."""

    def __init__(self, x, y, comps, mean, theta, noise_var=None, y_cov=None):
        self.x = np.ascontiguousarray(x, dtype=float)
        if self.x.ndim == 1:
            self.x = self.x.reshape(-1, 1)
        self.y = np.asarray(y, dtype=float).squeeze()
        self.n, self.d = self.x.shape
        self.comps, self.mean = tuple(comps), mean
        self.noise_var = None if noise_var is None else np.asarray(noise_var, dtype=float)
        self.y_cov = y_cov
        self.xbar = self.x.mean(axis=0)
        self.theta = np.asarray(theta, dtype=float)
        self.tm, self.parts = split_theta(theta, comps, mean, self.n, self.d)
        k = train_cov(self.comps, self.parts, self.x, self.noise_var, y_cov)
        self.mu = mean_vec(mean, self.tm, self.x, self.xbar)
        self.L = cholesky(k)                                                       # regression.py:241
        del k
        self.alpha = solve_triangular(                                              # regression.py:242-244
            self.L.T, solve_triangular(self.L, self.y - self.mu, lower=True))

    # regression.py:188-216 -- one dtrtrs per query in the reference; here `chunk` RHS at a time
    def predict(self, q, chunk=256):
        q = self._points(q)
        mu = np.empty(q.shape[0])
        var = np.empty(q.shape[0])
        for s in range(0, q.shape[0], chunk):
            qq = q[s:s + chunk]
            kqq = prior_var(self.comps, self.parts, qq)
            kqx = cross_cov(self.comps, self.parts, qq, self.x)
            mu[s:s + chunk] = kqx @ self.alpha + mean_vec(self.mean, self.tm, qq, self.xbar)
            v = solve_triangular(self.L, kqx.T, lower=True)
            var[s:s + chunk] = kqq - (v**2).sum(axis=0)
        return mu, np.sqrt(np.abs(var))

    def _points(self, q):
        q = np.asarray(q, dtype=float)
        if q.ndim <= 1 and self.d == 1:
            q = q.reshape(-1, 1)
        elif q.ndim == 1 and q.size == self.d:
            q = q.reshape(1, -1)
        return q

    def _se_theta(self):
        if self.comps != ("SE",):
            raise NotImplementedError("gradient_terms exists only for SquaredExponential (covariance.py:38-44, 257-266)")
        th = self.parts[0][1]
        return np.exp(th[0]), np.exp(th[1:])

    # regression.py:351-385 with covariance.py:257-266.  R is a length-d vector broadcast over
    # rows (reference behaviour, SURVEY.md section 7 "quirks"), reproduced as is.
    def gradient(self, q):
        a, ls = self._se_theta()
        q = self._points(q)
        means, covs = [], []
        for pnt in q:
            kqx = cross_cov(self.comps, self.parts, pnt[None, :], self.x)
            A = ((self.x - pnt[None, :]) / ls[None, :] ** 2).T
            R = (a / ls) ** 2
            Q = solve_triangular(self.L, (A * kqx).T, lower=True)
            means.append(A @ (kqx * self.alpha).T)
            covs.append(R - Q.T @ Q)
        return np.array(means).squeeze(), np.array(covs).squeeze()

    # regression.py:387-419
    def spatial_derivatives(self, q):
        a, ls = self._se_theta()
        q = self._points(q)
        dmu, dvar = [], []
        for pnt in q:
            kqx = cross_cov(self.comps, self.parts, pnt[None, :], self.x)
            A = ((self.x - pnt[None, :]) / ls[None, :] ** 2).T
            Q = solve_triangular(self.L.T, solve_triangular(self.L, kqx.T, lower=True))
            dmu.append(A @ (kqx * self.alpha).T)
            dvar.append(-2 * (A * kqx[None, :]) @ Q)
        return np.array(dmu).squeeze(), np.array(dvar).squeeze()

    # regression.py:421-449
    def posterior(self, q):
        q = self._points(q)
        kqx = cross_cov(self.comps, self.parts, q, self.x)
        kqq = cross_cov(self.comps, self.parts, q, q)
        mu = kqx @ self.alpha + mean_vec(self.mean, self.tm, q, self.xbar)
        Q = solve_triangular(self.L, kqx.T, lower=True)
        return mu, kqq - Q.T @ Q

    # regression.py:451-466
    def loo_predictions(self):
        ik = solve_triangular(self.L, np.eye(self.n), lower=True)
        ik = ik.T @ ik
        var = 1.0 / np.diag(ik)
        return self.y - self.alpha * var, np.sqrt(var)


def marginal_likelihood(x, y, comps, mean, theta, noise_var=None, y_cov=None):
    """regression.py:528-542 (value only; LinAlgError -> -1e50)."""
    x = np.asarray(x, dtype=float)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    n, d = x.shape
    tm, parts = split_theta(theta, comps, mean, n, d)
    k = train_cov(comps, parts, x, noise_var, y_cov)
    mu = mean_vec(mean, tm, x, x.mean(axis=0))
    try:
        L = cholesky(k)
    except np.linalg.LinAlgError:
        return -1e50
    v = solve_triangular(L, y - mu, lower=True)
    return -0.5 * (v @ v) - np.log(np.diagonal(L)).sum()


def marginal_likelihood_gradient(x, y, comps, mean, theta, noise_var=None, y_cov=None):
    """regression.py:544-567: explicit inverse through dtrtrs on the identity + `iK.T @ iK`,
    Q = alpha alpha^T - K^-1, grad_i = 0.5 sum(Q * dK_i^T)."""
    x = np.asarray(x, dtype=float)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    n, d = x.shape
    tm, parts = split_theta(theta, comps, mean, n, d)
    xbar = x.mean(axis=0)
    k, grad_k = cov_and_grads(comps, parts, x)
    if noise_var is not None:
        k = k + np.diag(noise_var)
    if y_cov is not None:
        k = k + y_cov
    mu = mean_vec(mean, tm, x, xbar)
    grad_mu = mean_grads(mean, x, xbar)
    L = cholesky(k)
    ik = solve_triangular(L, np.eye(n), lower=True)
    ik = ik.T @ ik
    alpha = ik @ (y - mu)
    lml = -0.5 * ((y - mu).T @ alpha) - np.log(np.diagonal(L)).sum()
    grad = np.zeros(len(theta))
    pm = len(tm)
    grad[:pm] = [(alpha * g).sum() for g in grad_mu]
    Q = alpha[:, None] * alpha[None, :] - ik
    grad[pm:] = [0.5 * (Q * g.T).sum() for g in grad_k]
    return lml, grad


def _leaf_rows_and_grads(kind, th, x, r0, r1):
    """Rows [r0, r1) of one plain kernel on the training data and of its gradient planes: the same expressions as
    _leaf_cov_and_grads (covariance.py:268-276, 350-365, 171-175), built for a row block only."""
    n, d = x.shape
    xr = x[r0:r1]
    idx = np.arange(r0, r1)
    grads = []
    if kind == "SE":
        a, ls = np.exp(th[0]), np.exp(th[1:])
        e = np.exp(-_scaled_half_sqdist(xr, x, ls))
        e[idx - r0, idx] += EPS_JITTER
        k = a**2 * e
        grads.append(2.0 * k)
        for i in range(d):
            grads.append(((xr[:, i, None] - x[None, :, i]) ** 2 / ls[i] ** 2) * k)
    elif kind == "RQ":
        a, q, ls = np.exp(th[0]), np.exp(th[1]), np.exp(th[2:])
        z = _scaled_half_sqdist(xr, x, ls)
        f = 1 + z / q
        lnf = np.log(f)
        e = np.exp(-q * lnf)
        e[idx - r0, idx] += EPS_JITTER
        k = a**2 * e
        grads.append(2.0 * k)
        grads.append(-k * (lnf * q - z / f))
        g = 2 * k / f
        for i in range(d):
            grads.append(g * (0.5 * (xr[:, i, None] - x[None, :, i]) ** 2 / ls[i] ** 2))
    elif kind == "WHITE":
        k = np.zeros((r1 - r0, n))
        k[idx - r0, idx] = np.exp(2 * th[0])
        grads.append(2.0 * k)
    else:
        raise NotImplementedError("row-blocked gradients: SE / RQ / WHITE only")
    return k, grads


def marginal_likelihood_gradient_blocked(x, y, comps, mean, theta, noise_var=None, block=1024, want_parts=False):
    """marginal_likelihood_gradient (regression.py:544-567) for sizes whose p dense gradient planes do not fit: the same
    LAPACK / BLAS calls on K (cholesky, dtrtrs on the identity, iK.T @ iK, iK @ r), but K and the dK planes are built
    one row block at a time and the traces 0.5 sum(Q * dK^T) are accumulated per block (dK is symmetric).  Plain SE / RQ /
    WHITE sums only.  tests/test_oracle_golden.py pins it to marginal_likelihood_gradient and to the reference fixtures."""
    x = np.asarray(x, dtype=float)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    n, d = x.shape
    tm, parts = split_theta(theta, comps, mean, n, d)
    assert parts.cp is None
    xbar = x.mean(axis=0)
    k = train_cov(comps, parts, x, noise_var)
    mu = mean_vec(mean, tm, x, xbar)
    L = cholesky(k)
    del k
    logdet = np.log(np.diagonal(L)).sum()
    ik = solve_triangular(L, np.eye(n), lower=True, overwrite_b=True, check_finite=False)
    del L
    ik = ik.T @ ik
    alpha = ik @ (y - mu)
    lml = -0.5 * ((y - mu).T @ alpha) - logdet
    grad = np.zeros(len(theta))
    pm = len(tm)
    grad[:pm] = [(alpha * g).sum() for g in mean_grads(mean, x, xbar)]
    for r0 in range(0, n, block):
        r1 = min(n, r0 + block)
        Q = alpha[r0:r1, None] * alpha[None, :] - ik[r0:r1]
        for kind, th, _, off in parts:
            _, gk = _leaf_rows_and_grads(kind, th, x, r0, r1)
            for j, g in enumerate(gk):
                grad[pm + off + j] += 0.5 * (Q * g).sum()
    if want_parts:
        return lml, grad, alpha
    return lml, grad


def loo_likelihood(x, y, comps, mean, theta, noise_var=None):
    """regression.py:468-487."""
    x = np.asarray(x, dtype=float)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    n, d = x.shape
    tm, parts = split_theta(theta, comps, mean, n, d)
    k = train_cov(comps, parts, x, noise_var)
    mu = mean_vec(mean, tm, x, x.mean(axis=0))
    try:
        L = cholesky(k)
    except np.linalg.LinAlgError:
        return -1e50
    ik = solve_triangular(L, np.eye(n), lower=True)
    ik = ik.T @ ik
    alpha = ik @ (y - mu)
    var = 1.0 / np.diag(ik)
    return -0.5 * (var * alpha**2 + np.log(var)).sum()


def loo_likelihood_gradient(x, y, comps, mean, theta, noise_var=None):
    """regression.py:489-526."""
    x = np.asarray(x, dtype=float)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    n, d = x.shape
    tm, parts = split_theta(theta, comps, mean, n, d)
    xbar = x.mean(axis=0)
    k, grad_k = cov_and_grads(comps, parts, x)
    if noise_var is not None:
        k = k + np.diag(noise_var)
    mu = mean_vec(mean, tm, x, xbar)
    grad_mu = mean_grads(mean, x, xbar)
    L = cholesky(k)
    ik = solve_triangular(L, np.eye(n), lower=True)
    ik = ik.T @ ik
    alpha = ik @ (y - mu)
    var = 1.0 / np.diag(ik)
    loo = -0.5 * (var * alpha**2 + np.log(var)).sum()
    c1 = alpha * var
    c2 = 0.5 * var * (1 + var * alpha**2)
    grad = np.zeros(len(theta))
    pm = len(tm)
    grad[pm:] = [(c1 * ((ik @ g) @ alpha) - c2 * np.diag((ik @ g) @ ik)).sum() for g in grad_k]
    grad[:pm] = [(c1 * (ik @ g)).sum() for g in grad_mu]
    return loo, grad


# ----------------------------------------------------------------------------- acquisition
_IR2PI = 1 / np.sqrt(2 * np.pi)
_IR2 = 1.0 / np.sqrt(2)
_RPI2 = np.sqrt(0.5 * np.pi)
_LN2PI = np.log(2 * np.pi)


def expected_improvement(mu, sig, y_max):
    """Vectorised ExpectedImprovement.__call__ (acquisition.py:76-86): branch at Z < -3."""
    mu, sig = np.asarray(mu, dtype=float), np.asarray(sig, dtype=float)
    Z = (mu - y_max) / sig
    out = np.empty_like(Z)
    lo = Z < -3
    zl = Z[lo]
    out[lo] = np.exp(np.log(1 + zl * _RPI2 * erfcx(-zl * _IR2)) - 0.5 * (zl**2 + _LN2PI) + np.log(sig[lo]))
    zh = Z[~lo]
    out[~lo] = sig[~lo] * (zh * 0.5 * (1.0 + erf(zh * _IR2)) + np.exp(-0.5 * zh**2) * _IR2PI)
    return out


def neg_log_ei(mu, sig, y_max):
    """Vectorised ExpectedImprovement.opt_func (acquisition.py:88-97)."""
    mu, sig = np.asarray(mu, dtype=float), np.asarray(sig, dtype=float)
    Z = (mu - y_max) / sig
    out = np.empty_like(Z)
    lo = Z < -3
    zl = Z[lo]
    out[lo] = np.log(1 + zl * _RPI2 * erfcx(-zl * _IR2)) - 0.5 * (zl**2 + _LN2PI) + np.log(sig[lo])
    zh = Z[~lo]
    out[~lo] = np.log(sig[~lo] * (zh * 0.5 * (1.0 + erf(zh * _IR2)) + np.exp(-0.5 * zh**2) * _IR2PI))
    return -out


def neg_log_ei_gradient(mu, sig, dmu, dvar, y_max):
    """Vectorised ExpectedImprovement.opt_func_gradient (acquisition.py:99-125).
    mu, sig: (M,), dmu, dvar: (M,d) -> (-ln EI (M,), -grad ln EI (M,d))."""
    mu, sig = np.asarray(mu, dtype=float), np.asarray(sig, dtype=float)
    dmu = np.asarray(dmu, dtype=float).reshape(mu.size, -1)
    dvar = np.asarray(dvar, dtype=float).reshape(mu.size, -1)
    Z = (mu - y_max) / sig
    val = np.empty_like(Z)
    grad = np.empty_like(dmu)
    for i in range(Z.size):
        z, s = Z[i], sig[i]
        if z < -3:
            R = _RPI2 * erfcx(-z * _IR2)
            H = 1 + z * R
            val[i] = np.log(H) - 0.5 * (z**2 + _LN2PI) + np.log(s)
            grad[i] = (0.5 * dvar[i] / s + R * dmu[i]) / (H * s)
        else:
            pdf = np.exp(-0.5 * z**2) * _IR2PI
            cdf = 0.5 * (1.0 + erf(z * _IR2))
            ei = s * (z * cdf + pdf)
            val[i] = np.log(ei)
            grad[i] = (0.5 * pdf * dvar[i] / s + dmu[i] * cdf) / ei
    return -val, -grad


def upper_confidence_bound(mu, sig, kappa):
    """UpperConfidenceBound.__call__ (acquisition.py:169-171): mu + kappa sigma; opt_func (:173-175) is its negative."""
    return np.asarray(mu) + kappa * np.asarray(sig)


def upper_confidence_bound_gradient(mu, sig, dmu, dvar, kappa):
    """UpperConfidenceBound.opt_func_gradient (acquisition.py:177-189): (-(mu + kappa sigma), -(dmu + 0.5 kappa dvar / sigma))
    for (M,) mu / sigma and (M, d) dmu / dvar."""
    mu, sig = np.asarray(mu), np.asarray(sig)
    dmu, dvar = np.asarray(dmu).reshape(mu.size, -1), np.asarray(dvar).reshape(mu.size, -1)
    return -(mu + kappa * sig), -(dmu + 0.5 * kappa * dvar / sig[:, None])


def max_variance(sig):
    """MaxVariance.__call__ (acquisition.py:213-215): sigma^2; opt_func (:217-219) is its negative; opt_func_gradient
    (:221-229) returns (-sigma^2, -dvar); convergence_metric (:231-232) is sigma."""
    return np.asarray(sig) ** 2


# ----------------------------------------------------------------------------- bounds
def _kind_bounds(kind, x, y):
    n, d = x.shape
    out = []
    if kind in ("SE", "RQ"):
        s = np.log(y.std())
        out.append((s - 4, s + 4))
        if kind == "RQ":
            out.append((-2, 6))
        for i in range(d):
            col = x[:, i]
            mad = 0.0
            for r0 in range(0, n, 1024):
                mad += np.abs(col[r0:r0 + 1024, None] - col[None, :]).sum()
            mad /= n * n
            out.append((np.log(mad) - 4, np.log(col.max() - col.min()) + 2))
    else:
        s = np.log(np.ptp(y))
        out.extend([(s - 8, s + 2)] * (1 if kind == "WHITE" else n))
    return out


def cov_bounds(comps, x, y):
    """estimate_hyperpar_bounds: covariance.py:228-238 (SE), :323-333 (RQ), :151-158 (White), :662-669 (Hetero),
    :496-511 (ChangePoint: region kernels, then (location, width) bounds per change-point).  mean_ij|dx| is over all
    N^2 ordered pairs including the zero diagonal."""
    out = []
    for term in comps:
        if isinstance(term, str):
            out.extend(_kind_bounds(term, x, y))
        else:
            _, axis, regions = term
            for region in regions:
                for kind in region:
                    out.extend(_kind_bounds(kind, x, y))
            lo, hi = x[:, axis].min(), x[:, axis].max()
            for _ in range(len(regions) - 1):
                out.append((lo, hi))
                out.append((5e-3 * (hi - lo), 0.5 * (hi - lo)))
    return out


def mean_bounds(mean, x, y):
    """mean.py:40-42, :68-72, :104-109."""
    w = y.max() - y.min()
    if mean == "const":
        return [(y.min() - w, y.max() + w)]
    dx = x - x.mean(axis=0)[None, :]
    gb = 10 * w / (dx.max(axis=0) - dx.min(axis=0))
    out = [(y.min() - 2 * w, y.max() + 2 * w)]
    out.extend([(-b, b) for b in gb])
    if mean == "quadratic":
        out.extend([(-b, b) for b in gb])
    return out


# ----------------------------------------------------------------------------- GpLinearInverter (inversion.py)
class LinearInverter:
    """Restatement of inference/gp/inversion.py: posterior of y = A x + noise with a GP prior on x.

    comps / mean / theta as above; x_pos = parameter_spatial_positions (n x d); A = model_matrix (m x n)."""

    def __init__(self, y, y_err, A, x_pos, comps, mean):
        self.y, self.y_err = np.asarray(y, dtype=float), np.asarray(y_err, dtype=float)
        self.A = np.asarray(A, dtype=float)
        self.x = np.asarray(x_pos, dtype=float)
        self.comps, self.mean = tuple(comps), mean
        self.n, self.d = self.x.shape
        self.xbar = self.x.mean(axis=0)
        self.sigma = np.diag(self.y_err**2)            # inversion.py:134
        self.inv_sigma = np.diag(self.y_err**-2.0)     # :135

    def _prior(self, theta):
        tm, parts = split_theta(theta, self.comps, self.mean, self.n, self.d)
        K = train_cov(self.comps, parts, self.x)       # cov.build_covariance(theta) -- no data-error term (:149)
        return K, mean_vec(self.mean, tm, self.x, self.xbar), tm, parts

    def calculate_posterior(self, theta):
        """inversion.py:138-155"""
        K, mu, _, _ = self._prior(theta)
        W = self.A.T @ self.inv_sigma @ self.A
        u = self.A.T @ (self.inv_sigma @ (self.y - self.A @ mu))
        from scipy.linalg import solve
        cov = solve(np.eye(self.n) + K @ W, K)
        return cov @ u + mu, cov

    def marginal_likelihood(self, theta):
        """inversion.py:174-188"""
        K, mu, _, _ = self._prior(theta)
        L = cholesky(self.A @ K @ self.A.T + self.sigma)
        v = solve_triangular(L, self.y - self.A @ mu, lower=True)
        return -0.5 * (v @ v) - np.log(np.diagonal(L)).sum()

    def marginal_likelihood_gradient(self, theta):
        """inversion.py:190-217"""
        tm, parts = split_theta(theta, self.comps, self.mean, self.n, self.d)
        K, grad_K = cov_and_grads(self.comps, parts, self.x)
        J = self.A @ K @ self.A.T + self.sigma
        grad_J = [self.A @ dK @ self.A.T for dK in grad_K]
        mu = mean_vec(self.mean, tm, self.x, self.xbar)
        grad_f = [self.A @ du for du in mean_grads(self.mean, self.x, self.xbar)]
        f = self.A @ mu
        L = cholesky(J)
        iJ = solve_triangular(L, np.eye(L.shape[0]), lower=True)
        iJ = iJ.T @ iJ
        alpha = iJ @ (self.y - f)
        lml = -0.5 * ((self.y - f) @ alpha) - np.log(np.diagonal(L)).sum()
        grad = np.zeros(len(theta))
        pm = len(tm)
        grad[:pm] = [(alpha * df).sum() for df in grad_f]
        Q = alpha[:, None] * alpha[None, :] - iJ
        grad[pm:] = [0.5 * (Q * dJ.T).sum() for dJ in grad_J]
        return lml, grad
