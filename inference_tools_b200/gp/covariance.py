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

SUPPORTED = "SquaredExponential, RationalQuadratic, WhiteNoise, HeteroscedasticNoise, ChangePoint and sums of them"


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
        return [leaf["kind"] for leaf in self.layout()["leaves"]]

    def _components(self):
        return [self]

    def layout(self) -> dict:
        """Flattened description handed to the engine (gpb_set_model_ex): `leaves` = plain kernels in evaluation order,
        each with its parameter offset inside this function's theta and the ChangePoint region it belongs to (-1 = none);
        `cp` = None or dict(axis, theta_off, n_regions).  Needs pass_spatial_data first (parameter counts)."""
        return {"leaves": [{"kind": self.kind, "off": 0, "region": -1}], "cp": None, "n_params": self.n_params}

    def engine_layout(self) -> dict:
        lay = self.layout()
        cp = lay["cp"] or {"axis": 0, "theta_off": 0, "n_regions": 0}
        return {"theta_offs": [leaf["off"] for leaf in lay["leaves"]], "regions": [leaf["region"] for leaf in lay["leaves"]],
                "n_regions": cp["n_regions"], "cp_axis": cp["axis"], "cp_theta_off": cp["theta_off"], "n_params": lay["n_params"]}

    def _ensure_engine(self):
        if self._x is None:
            raise RuntimeError("pass_spatial_data(x) must be called before evaluating the covariance")
        if self._engine is None:
            eng = _lib.Engine()
            eng.set_data(self._x, np.zeros(self._x.shape[0]))
            eng.set_model(self.kinds(), _lib.MEAN_CONST, self.engine_layout())
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
        self.components = covariance_components
        self.bounds = None

    def _components(self):
        out = []
        for comp in self.components:
            out.extend(comp._components())
        return out

    def layout(self):
        leaves, cp, off = [], None, 0
        for comp in self.components:
            lay = comp.layout()
            leaves.extend({"kind": leaf["kind"], "off": leaf["off"] + off, "region": leaf["region"]} for leaf in lay["leaves"])
            if lay["cp"] is not None:
                if cp is not None:
                    raise ValueError("at most one ChangePoint kernel per covariance function is supported")
                cp = dict(lay["cp"], theta_off=lay["cp"]["theta_off"] + off)
            off += lay["n_params"]
        return {"leaves": leaves, "cp": cp, "n_params": off}

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


class ChangePoint(CovarianceFunction):
    r"""Divides the input space into regions along one axis, each modelled by its own kernel
    (reference covariance.py:371-605):  K_cp(u, v) = sum_r g_r(u) K_r(u, v) g_r(v) with logistic region weights
    f_i(x) = 1 / (1 + exp(-(x - c_i) / w_i)), g_0 = 1 - f_0, g_r = f_{r-1} (1 - f_r), g_last = f_last.
    theta = [theta of K_0, ..., theta of K_{n-1}, c_0, w_0, c_1, w_1, ...] (locations and widths are not in log space).

    :param kernels: the kernel classes / objects (K0, K1, ...), one per region (plain kernels or sums of them)
    :param int axis: the spatial axis along which the regions are divided
    :param location_bounds: n-1 (lower, upper) pairs for the change-point locations
    :param width_bounds: n-1 (lower, upper) pairs for the change-point widths
    """

    kind = -2  # not a leaf: described through layout()

    def __init__(self, kernels, axis: int = 0, location_bounds=None, width_bounds=None):
        super().__init__()
        from inspect import isclass

        self.cov = [K() if isclass(K) and issubclass(K, CovarianceFunction) else K for K in kernels]
        for K in self.cov:
            if not isinstance(K, CovarianceFunction) or isinstance(K, ChangePoint):
                raise TypeError(
                    """\n
                    \r[ ChangePoint error ]
                    \r>> Each of the specified covariance kernels must be an instance of
                    \r>> a class which inherits from the 'CovarianceFunction' abstract
                    \r>> base-class.
                    """
                )
        self.n_kernels = len(kernels)
        if not 2 <= self.n_kernels <= _lib.MAX_REG:
            raise ValueError(f"[ ChangePoint error ] between 2 and {_lib.MAX_REG} region kernels are supported")
        for name, given in (("location_bounds", location_bounds), ("width_bounds", width_bounds)):
            if given is not None and len(given) != self.n_kernels - 1:
                raise ValueError(
                    f"""\n
                    \r[ ChangePoint error ]
                    \r>> The length of '{name}' must be one less than the number of kernels
                    """
                )
        self.location_bounds = None if location_bounds is None else [check_bounds(b) for b in location_bounds]
        self.width_bounds = None if width_bounds is None else [check_bounds(b) for b in width_bounds]
        self.axis = axis
        self.bounds = None

    def _components(self):
        out = []
        for K in self.cov:
            out.extend(K._components())
        return out

    def layout(self):
        leaves, off = [], 0
        for r, K in enumerate(self.cov):
            lay = K.layout()
            if lay["cp"] is not None:
                raise ValueError("ChangePoint kernels cannot be nested")
            leaves.extend({"kind": leaf["kind"], "off": leaf["off"] + off, "region": r} for leaf in lay["leaves"])
            off += lay["n_params"]
        cp = {"axis": self.axis, "theta_off": off, "n_regions": self.n_kernels}
        return {"leaves": leaves, "cp": cp, "n_params": off + 2 * (self.n_kernels - 1)}

    def pass_spatial_data(self, x: np.ndarray):
        for K in self.cov:
            K.pass_spatial_data(x)
        self._register_data(x)
        counts = [K.n_params for K in self.cov] + [2] * (self.n_kernels - 1)
        self.n_params = sum(counts)
        slices = slice_builder(counts)
        self.cov_slc, self.cp_slc = slices[: self.n_kernels], slices[self.n_kernels:]
        self.hyperpar_labels = []
        for i, K in enumerate(self.cov):
            self.hyperpar_labels.extend(f"ChngPnt K{i}: {lab}" for lab in K.hyperpar_labels)
        for i in range(self.n_kernels - 1):
            self.hyperpar_labels.extend([f"ChngPnt{i} location", f"ChngPnt{i} width"])
        self.x_cp = x[:, self.axis]

    def estimate_hyperpar_bounds(self, y: np.ndarray):
        xr = self.x_cp.min(), self.x_cp.max()
        dx = xr[1] - xr[0]
        self.bounds = []
        for K in self.cov:
            K.estimate_hyperpar_bounds(y)
            self.bounds.extend(K.bounds)
        if self.location_bounds is None:
            self.location_bounds = [xr] * (self.n_kernels - 1)
        if self.width_bounds is None:
            self.width_bounds = [(5e-3 * dx, 0.5 * dx)] * (self.n_kernels - 1)
        for loc, wid in zip(self.location_bounds, self.width_bounds):
            self.bounds.extend([loc, wid])

    @staticmethod
    def logistic(x, theta):
        z = (x - theta[0]) / theta[1]
        return 1.0 / (1.0 + np.exp(-z))


def check_bounds(bounds):
    """covariance.py:700-705"""
    if bounds is not None:
        assert type(bounds) in [list, tuple, np.ndarray]
        assert len(bounds) == 2
        assert bounds[1] > bounds[0]
    return bounds


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
    if not isinstance(cov, CovarianceFunction) or any(c.kind < 0 for c in cov._components()) \
            or len(cov._components()) > _lib.MAX_COMP:
        raise TypeError(
            f"""\n
            [ GpRegressor error ]
            >> The covariance function {type(cov)} cannot run in the CUDA engine and there is
            >> no CPU fallback. Supported kernels: {SUPPORTED}.
            """
        )
    return cov
