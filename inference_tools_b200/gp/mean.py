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

    def _offsets(self, x):
        self.x_mean = x.mean(axis=0)
        self.dx = x - self.x_mean[None, :]
        self.n_data, self.n_dim = x.shape

    def _slope_bounds(self, y):
        # mean.py:68-72, 104-109
        w = y.max() - y.min()
        g = 10 * w / (self.dx.max(axis=0) - self.dx.min(axis=0))
        return (y.min() - 2 * w, y.max() + 2 * w), [(-b, b) for b in g]


class ConstantMean(MeanFunction):
    """mu = theta_0 (reference mean.py:31-51)."""

    kind = _lib.MEAN_CONST

    def __init__(self, hyperpar_bounds=None):
        self.bounds = hyperpar_bounds
        self.n_params = 1
        self.hyperpar_labels = ["ConstantMean"]

    def pass_spatial_data(self, x):
        self.n_data = x.shape[0]

    def estimate_hyperpar_bounds(self, y):
        w = y.max() - y.min()
        self.bounds = [(y.min() - w, y.max() + w)]

    def __call__(self, q, theta):
        return theta[0]

    def build_mean(self, theta):
        return np.zeros(self.n_data) + theta[0]

    def mean_and_gradients(self, theta):
        return self.build_mean(theta), [np.ones(self.n_data)]


class LinearMean(MeanFunction):
    """mu = theta_0 + (x - xbar) . theta_1..d (reference mean.py:54-83)."""

    kind = _lib.MEAN_LINEAR

    def __init__(self, hyperpar_bounds=None):
        self.bounds = hyperpar_bounds

    def pass_spatial_data(self, x):
        self._offsets(x)
        self.n_params = 1 + self.n_dim
        self.hyperpar_labels = ["LinearMean background"]
        self.hyperpar_labels.extend(f"LinearMean gradient {i}" for i in range(self.n_dim))

    def estimate_hyperpar_bounds(self, y):
        level, slopes = self._slope_bounds(y)
        self.bounds = [level, *slopes]

    def __call__(self, q, theta):
        return theta[0] + np.dot(q - self.x_mean, theta[1:]).squeeze()

    def build_mean(self, theta):
        return theta[0] + np.dot(self.dx, theta[1:])

    def mean_and_gradients(self, theta):
        return self.build_mean(theta), [np.ones(self.n_data), *self.dx.T]


class QuadraticMean(MeanFunction):
    """mu = theta_0 + (x - xbar) . theta_lin + (x - xbar)^2 . theta_quad (reference mean.py:86-126)."""

    kind = _lib.MEAN_QUADRATIC

    def __init__(self, hyperpar_bounds=None):
        self.bounds = hyperpar_bounds

    def pass_spatial_data(self, x):
        self._offsets(x)
        n = self.n_dim
        self.dx_sqr = self.dx**2
        self.n_params = 1 + 2 * n
        self.hyperpar_labels = ["mean_background"]
        self.hyperpar_labels.extend(f"mean_linear_coeff_{i}" for i in range(n))
        self.hyperpar_labels.extend(f"mean_quadratic_coeff_{i}" for i in range(n))
        self.lin_slc = slice(1, n + 1)
        self.quad_slc = slice(n + 1, 2 * n + 1)

    def estimate_hyperpar_bounds(self, y):
        level, slopes = self._slope_bounds(y)
        self.bounds = [level, *slopes, *slopes]

    def __call__(self, q, theta):
        d = q - self.x_mean
        return theta[0] + np.dot(d, theta[self.lin_slc]).squeeze() + np.dot(d**2, theta[self.quad_slc]).squeeze()

    def build_mean(self, theta):
        return theta[0] + np.dot(self.dx, theta[self.lin_slc]) + np.dot(self.dx_sqr, theta[self.quad_slc])

    def mean_and_gradients(self, theta):
        return self.build_mean(theta), [np.ones(self.n_data), *self.dx.T, *self.dx_sqr.T]


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
