"""Covariance-function plug-ins with the interface of ``inference.gp.covariance`` (reference
covariance.py:8-44) -- descriptors for the CUDA engine instead of numpy array factories.

Inside :class:`~inference_tools_b200.gp.regression.GpRegressor` a covariance object is only a
*description* (component kinds + parameter layout): the regressor hands the kinds to one engine
context and never calls back into Python on the hot path.  The stand-alone methods of the reference
protocol (``__call__``, ``build_covariance``, ``covariance_and_gradients``) are kept for drop-in
compatibility and run on the GPU through a private engine context created by ``pass_spatial_data``.
Nothing here stores an (N, N, d) array (reference covariance.py:218-219, 315-316).
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from inference_tools_b200 import _lib

SUPPORTED = "SquaredExponential, RationalQuadratic, WhiteNoise, HeteroscedasticNoise and sums of them"


def mean_abs_difference(col: np.ndarray) -> float:
    """mean_ij |c_i - c_j| over all N^2 ordered pairs (zeros on the diagonal included), i.e. the value of
    ``abs(dx[:, :, i]).mean()`` in covariance.py:236, 331, from the sorted-sample identity
    sum_ij |c_i - c_j| = 2 sum_n (2n - N + 1) c_(n) in O(N log N) instead of O(N^2)."""
    c = np.sort(np.asarray(col, dtype=float))
    n = c.size
    w = 2.0 * np.arange(n) - n + 1.0
    return 2.0 * float(np.dot(w, c)) / (n * n)


class CovarianceFunction(ABC):
    """Base class (reference covariance.py:8-44).  Only the engine-native kinds can be used."""

    kind: int = -1
    bounds = None
    n_params: int
    hyperpar_labels: list

    def __init__(self):
        self._engine = None
        self._x = None

    # ---- composition (covariance.py:33-36)
    def __add__(self, other):
        if not isinstance(other, CovarianceFunction):
            raise TypeError(f"cannot add {type(other)} to a covariance function; supported: {SUPPORTED}")
        k1 = self.components if isinstance(self, CompositeCovariance) else [self]
        k2 = other.components if isinstance(other, CompositeCovariance) else [other]
        return CompositeCovariance([*k1, *k2])

    # ---- engine description
    def kinds(self) -> list:
        return [self.kind]

    def _components(self):
        return [self]

    def _ensure_engine(self):
        if self._x is None:
            raise RuntimeError("pass_spatial_data(x) must be called before evaluating the covariance")
        if self._engine is None:
            eng = _lib.Engine()
            eng.set_data(self._x, np.zeros(self._x.shape[0]))
            eng.set_model(self.kinds(), _lib.MEAN_CONST)
            self._engine = eng
        return self._engine

    def _register_data(self, x: np.ndarray):
        self._x = np.ascontiguousarray(x, dtype=float)
        self._engine = None

    @abstractmethod
    def pass_spatial_data(self, x: np.ndarray):
        pass

    @abstractmethod
    def estimate_hyperpar_bounds(self, y: np.ndarray):
        pass

    # ---- reference protocol, evaluated on the GPU
    def __call__(self, u: np.ndarray, v: np.ndarray, theta: np.ndarray) -> np.ndarray:
        """cov(u, v, theta) (covariance.py:240-245, 335-341, 160-161, 671-672, 86-89)."""
        u = np.asarray(u, dtype=float)
        v = np.asarray(v, dtype=float)
        return self._ensure_engine().cross_covariance(u, v, theta)

    def build_covariance(self, theta: np.ndarray) -> np.ndarray:
        """Data covariance K(theta) without the error term (covariance.py:247-255, 343-348, 163-169, 674-680)."""
        return self._ensure_engine().build_covariance(theta, add_sig=False)

    def covariance_and_gradients(self, theta: np.ndarray):
        """K(theta) and the list of dK/dtheta_i (covariance.py:268-276, 350-365, 171-175, 682-686, 97-105)."""
        k, dk = self._ensure_engine().covariance_and_gradients(theta)
        return k, [g for g in dk]

    def gradient_terms(self, v, x, theta):
        raise NotImplementedError(
            f"""
            Gradient calculations are not yet available for the
            {type(self)} covariance function.
            """
        )

    def get_bounds(self):
        return self.bounds

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None  # device handles are not picklable; rebuilt lazily
        return state


class CompositeCovariance(CovarianceFunction):
    """Sum of covariance functions (covariance.py:47-105)."""

    def __init__(self, covariance_components: list):
        super().__init__()
        for c in covariance_components:
            if not isinstance(c, CovarianceFunction) or isinstance(c, CompositeCovariance):
                raise TypeError(f"unsupported covariance component {type(c)}; supported: {SUPPORTED}")
        if len(covariance_components) > _lib.MAX_COMP:
            raise ValueError(f"at most {_lib.MAX_COMP} covariance components are supported")
        self.components = covariance_components
        self.bounds = None

    def kinds(self):
        return [c.kind for c in self.components]

    def _components(self):
        return self.components

    def pass_spatial_data(self, x: np.ndarray):
        for comp in self.components:
            comp.pass_spatial_data(x)
        self._register_data(x)
        self.slices = slice_builder([c.n_params for c in self.components])
        self.hyperpar_labels = []
        for i, comp in enumerate(self.components):
            self.hyperpar_labels.extend(f"K{i+1}: {s}" for s in comp.hyperpar_labels)
        self.n_params = sum(c.n_params for c in self.components)

    def estimate_hyperpar_bounds(self, y: np.ndarray):
        self.bounds = []
        for comp in self.components:
            if comp.bounds is None:
                comp.estimate_hyperpar_bounds(y)
            self.bounds.extend(comp.bounds)


class _SmoothKernel(CovarianceFunction):
    """Shared data handling of the two stationary kernels."""

    def __init__(self, hyperpar_bounds=None):
        super().__init__()
        self.bounds = hyperpar_bounds

    def _length_scale_bounds(self):
        # covariance.py:235-238, 330-333: (ln mean|dx| - 4, ln max dx + 2) per dimension
        out = []
        for i in range(self._x.shape[1]):
            col = self._x[:, i]
            out.append((np.log(mean_abs_difference(col)) - 4, np.log(col.max() - col.min()) + 2))
        return out


class SquaredExponential(_SmoothKernel):
    r"""K(u, v) = A^2 exp(-1/2 sum_i ((u_i - v_i) / l_i)^2); theta = [ln A, ln l_1, ..., ln l_n]
    (reference covariance.py:181-279)."""

    kind = _lib.COV_SE

    def pass_spatial_data(self, x: np.ndarray):
        self._register_data(x)
        self.n_params = x.shape[1] + 1
        self.hyperpar_labels = ["SqrExp log-amplitude"]
        self.hyperpar_labels.extend(f"SqrExp log-scale {i}" for i in range(x.shape[1]))

    def estimate_hyperpar_bounds(self, y: np.ndarray):
        s = np.log(y.std())
        self.bounds = [(s - 4, s + 4), *self._length_scale_bounds()]

    def gradient_terms(self, v: np.ndarray, x: np.ndarray, theta: np.ndarray):
        """A = (x - v) / l^2 as (d, N) and R = (a / l)^2 (covariance.py:257-266).  O(N d) host helper kept
        for API parity; GpRegressor.gradient / spatial_derivatives fuse these terms inside the CUDA path."""
        a = np.exp(theta[0])
        ls = np.exp(theta[1:])
        return ((x - v[None, :]) / ls[None, :] ** 2).T, (a / ls) ** 2


class RationalQuadratic(_SmoothKernel):
    r"""K(u, v) = A^2 (1 + 1/(2 alpha) sum_i ((u_i - v_i) / l_i)^2)^(-alpha);
    theta = [ln A, ln alpha, ln l_1, ..., ln l_n] (reference covariance.py:282-368)."""

    kind = _lib.COV_RQ

    def pass_spatial_data(self, x: np.ndarray):
        self._register_data(x)
        self.n_params = x.shape[1] + 2
        self.hyperpar_labels = ["RQ log-amplitude", "RQ log-alpha"]
        self.hyperpar_labels.extend(f"RQ log-scale {i}" for i in range(x.shape[1]))

    def estimate_hyperpar_bounds(self, y: np.ndarray):
        s = np.log(y.std())
        self.bounds = [(s - 4, s + 4), (-2, 6), *self._length_scale_bounds()]


class WhiteNoise(CovarianceFunction):
    r"""K(x_i, x_j) = delta_ij sigma_n^2; theta = [ln sigma_n] (reference covariance.py:108-178)."""

    kind = _lib.COV_WHITE

    def __init__(self, hyperpar_bounds=None):
        super().__init__()
        self.bounds = hyperpar_bounds
        self.n_params = 1
        self.hyperpar_labels = ["WhiteNoise log-sigma"]

    def pass_spatial_data(self, x: np.ndarray):
        self._register_data(x)

    def estimate_hyperpar_bounds(self, y: np.ndarray):
        s = np.log(np.ptp(y))
        self.bounds = [(s - 8, s + 2)]


class HeteroscedasticNoise(CovarianceFunction):
    r"""K(x_i, x_j) = delta_ij sigma_i^2; theta = [ln sigma_1, ..., ln sigma_m]
    (reference covariance.py:608-689)."""

    kind = _lib.COV_HETERO

    def __init__(self, hyperpar_bounds=None):
        super().__init__()
        self.bounds = hyperpar_bounds

    def pass_spatial_data(self, x: np.ndarray):
        self._register_data(x)
        self.n_params = x.shape[0]
        self.hyperpar_labels = [f"log_sigma_{i+1}" for i in range(self.n_params)]

    def estimate_hyperpar_bounds(self, y: np.ndarray):
        s = np.log(np.ptp(y))
        self.bounds = [(s - 8, s + 2) for _ in range(self.n_params)]


def slice_builder(lengths: list) -> list:
    """Consecutive parameter slices of the components (covariance.py:692-697)."""
    out, start = [], 0
    for n in lengths:
        out.append(slice(start, start + n))
        start += n
    return out


def as_engine_covariance(kernel) -> CovarianceFunction:
    """Accept a class or an instance (regression.py:136) and reject anything the CUDA engine cannot run."""
    from inspect import isclass

    cov = kernel() if isclass(kernel) else kernel
    if not isinstance(cov, CovarianceFunction) or any(c.kind < 0 for c in cov._components()):
        raise TypeError(
            f"""\n
            [ GpRegressor error ]
            >> The covariance function {type(cov)} cannot run in the CUDA engine and there is
            >> no CPU fallback. Supported kernels: {SUPPORTED}.
            """
        )
    return cov
