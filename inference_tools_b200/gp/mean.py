"""Mean-function plug-ins with the interface of ``inference.gp.mean`` (reference mean.py:5-126).

Inside GpRegressor these are descriptors: the residual y - mu(theta), the mean of a query point and
the mean-parameter gradients are fused into the CUDA kernels (kernels.cu ``mean_at``, lml.cu
``diag_terms_kernel``).  The O(N d) stand-alone methods of the reference protocol are kept as host
helpers for API compatibility; nothing on the engine's path calls them.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from inference_tools_b200 import _lib


class MeanFunction(ABC):
    """Plug-in protocol of the reference (mean.py:5-28): ``pass_spatial_data``, ``estimate_hyperpar_bounds``,
    ``build_mean`` plus ``n_params``, ``bounds``, ``hyperpar_labels``.  ``kind`` selects the CUDA implementation."""

    kind: int = -1
    bounds = None
    n_params: int
    hyperpar_labels: list

    @abstractmethod
    def pass_spatial_data(self, x: np.ndarray):
        pass

    @abstractmethod
    def estimate_hyperpar_bounds(self, y: np.ndarray):
        pass

    @abstractmethod
    def build_mean(self, theta: np.ndarray):
        pass


class _PolynomialMean(MeanFunction):
    """The three reference means are one family: mu(x) = theta_0 + sum_{p=1..order} (x - xbar)^p . theta_p, with the
    powers taken per coordinate and theta_p a block of d coefficients (order 0: mean.py:31-51, 1: :54-83, 2: :86-126).
    Everything below is written once for an arbitrary ``order``; the public classes only fix it and the label texts."""

    order = 0
    _labels = ("ConstantMean",)  # label of theta_0, then one prefix per power

    def __init__(self, hyperpar_bounds=None):
        self.bounds = hyperpar_bounds
        if self.order == 0:
            self.n_params, self.hyperpar_labels = 1, [self._labels[0]]

    # -- data-dependent bookkeeping
    def pass_spatial_data(self, x):
        self.n_data, self.n_dim = x.shape[0], (x.shape[1] if x.ndim > 1 else 1)
        if self.order == 0:
            return
        self.x_mean = x.mean(axis=0)
        self.dx = x - self.x_mean[None, :]
        self._powers = [self.dx**p for p in range(1, self.order + 1)]
        if self.order >= 2:
            self.dx_sqr = self._powers[1]
        d = self.n_dim
        self.n_params = 1 + self.order * d
        self.hyperpar_labels = [self._labels[0]] + [f"{pre}{i}" for pre in self._labels[1:] for i in range(d)]
        self._blocks = [slice(1 + (p - 1) * d, 1 + p * d) for p in range(1, self.order + 1)]
        if self.order >= 2:
            self.lin_slc, self.quad_slc = self._blocks[0], self._blocks[1]

    def estimate_hyperpar_bounds(self, y):
        lo, hi = y.min(), y.max()
        span = hi - lo
        if self.order == 0:
            self.bounds = [(lo - span, hi + span)]  # mean.py:40-42
            return
        # mean.py:68-72, 104-109: level within two spans, every coefficient within +-10 spans per unit extent
        reach = 10 * span / (self.dx.max(axis=0) - self.dx.min(axis=0))
        coeff = [(-r, r) for r in reach]
        self.bounds = [(lo - 2 * span, hi + 2 * span)] + coeff * self.order

    # -- stand-alone evaluation (API compatibility; the engine fuses these into its kernels)
    def __call__(self, q, theta):
        out = theta[0]
        if self.order:
            dq = q - self.x_mean
            for p, blk in enumerate(self._blocks, start=1):
                out = out + np.dot(dq**p, theta[blk]).squeeze()
        return out

    def build_mean(self, theta):
        out = np.full(self.n_data, float(theta[0]))
        for pw, blk in zip(getattr(self, "_powers", ()), getattr(self, "_blocks", ())):
            out = out + pw @ theta[blk]
        return out

    def mean_and_gradients(self, theta):
        grads = [np.ones(self.n_data)]
        for pw in getattr(self, "_powers", ()):
            grads.extend(pw.T)
        return self.build_mean(theta), grads


class ConstantMean(_PolynomialMean):
    """mu = theta_0 (reference mean.py:31-51)."""

    kind = _lib.MEAN_CONST
    order = 0
    _labels = ("ConstantMean",)


class LinearMean(_PolynomialMean):
    """mu = theta_0 + (x - xbar) . theta_1..d (reference mean.py:54-83)."""

    kind = _lib.MEAN_LINEAR
    order = 1
    _labels = ("LinearMean background", "LinearMean gradient ")


class QuadraticMean(_PolynomialMean):
    """mu = theta_0 + (x - xbar) . theta_lin + (x - xbar)^2 . theta_quad (reference mean.py:86-126)."""

    kind = _lib.MEAN_QUADRATIC
    order = 2
    _labels = ("mean_background", "mean_linear_coeff_", "mean_quadratic_coeff_")


def as_engine_mean(mean) -> MeanFunction:
    from inspect import isclass

    m = mean() if isclass(mean) else mean
    if not isinstance(m, MeanFunction) or m.kind < 0:
        raise TypeError(
            f"""\n
            [ GpRegressor error ]
            >> The mean function {type(m)} cannot run in the CUDA engine and there is no CPU
            >> fallback. Supported: ConstantMean, LinearMean, QuadraticMean.
            """
        )
    return m
