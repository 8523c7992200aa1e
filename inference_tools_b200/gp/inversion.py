"""``GpLinearInverter`` with the reference's public surface (reference inference/gp/inversion.py:11-249), computed by
the CUDA engine (``csrc/inverter.cu`` behind ``gpb_linv_*``).

The reference evaluates every quantity with dense numpy / LAPACK calls on the host: ``A K A^T`` products, a Cholesky
factorisation per objective evaluation, one ``A dK_p A^T`` per hyper-parameter for the gradient and a general
``solve(I + K W, K)`` for the posterior.  Here the matrices live in HBM, the products run on the FP64 tensor-core GEMM,
and the gradient is contracted in parameter space, ``0.5 sum (a a^T - M) o dK_p`` with ``a = A^T alpha`` and
``M = A^T inv(J) A``, so the same fused trace kernels as ``GpRegressor.marginal_likelihood_gradient`` apply and no
``dK_p`` is ever stored.
"""
from __future__ import annotations

import numpy as np
from numpy import ndarray
from numpy.linalg import LinAlgError
from scipy.optimize import minimize

from inference_tools_b200 import _lib
from inference_tools_b200.gp.covariance import CovarianceFunction, SquaredExponential, as_engine_covariance
from inference_tools_b200.gp.mean import ConstantMean, MeanFunction, as_engine_mean


def _error_text(lines) -> str:
    """The reference's message layout: a blank line, the class tag, then the '>>' lines, all indented alike."""
    pad = " " * 16
    return "\n\n" + pad + "[ GpLinearInverter error ]\n" + "".join(pad + ln + "\n" for ln in lines) + pad


class GpLinearInverter:
    """Gaussian-process linear inversion: posterior of ``y = A x + noise`` under a GP prior on ``x``.

    Arguments as the reference constructor (inversion.py:55-63): ``y``, ``y_err`` (1D, equal size), ``model_matrix``
    (2D, ``(y.size, n_parameters)``), ``parameter_spatial_positions`` (2D, ``(n_parameters, n_dimensions)``),
    ``prior_covariance_function`` and ``prior_mean_function`` as a class or an instance.  ``device`` selects the GPU
    (default ``$GPB200_DEVICE`` / ``$LOCAL_RANK`` / 0, as ``GpRegressor``).

    Difference from the reference: ``calculate_posterior(_mean)`` factor ``A K A^T + Sigma`` (Woodbury form) and raise
    ``numpy.linalg.LinAlgError`` when that matrix is not positive definite, where the reference's general
    ``solve(I + K W, K)`` (inversion.py:150-153) would still return a (meaningless) result.
    """

    def __init__(
        self,
        y: ndarray,
        y_err: ndarray,
        model_matrix: ndarray,
        parameter_spatial_positions: ndarray,
        prior_covariance_function: CovarianceFunction = SquaredExponential,
        prior_mean_function: MeanFunction = ConstantMean,
        device: int = None,
    ):
        # the reference's checks, in its order and with its texts (inversion.py:64-112)
        A, pos = model_matrix, parameter_spatial_positions
        checks = (
            (A.ndim != 2, [">> 'model_matrix' argument must be a 2D numpy.ndarray"]),
            (y.ndim != y_err.ndim != 1 or y.size != y_err.size,
             [">> 'y' and 'y_err' arguments must be 1D numpy.ndarray", ">> of equal size."]),
            (A.ndim == 2 and A.shape[0] != y.size,
             [">> The size of the first dimension of 'model_matrix' must",
              ">> equal the size of 'y', however they have shapes",
              f">> {A.shape}, {y.shape}", ">> respectively."]),
            (pos.ndim != 2,
             [">> 'parameter_spatial_positions' must be a 2D numpy.ndarray, with the",
              ">> size of first dimension being equal to the number of model parameters",
              ">> and the size of the second dimension being equal to the number of",
              ">> spatial dimensions."]),
            (A.ndim == 2 and pos.ndim == 2 and A.shape[1] != pos.shape[0],
             [">> The size of the second dimension of 'model_matrix' must be equal",
              ">> to the size of the first dimension of 'parameter_spatial_positions',",
              ">> however they have shapes", f">> {A.shape}, {pos.shape}", ">> respectively."]),
            (pos.ndim == 2 and pos.shape[1] > _lib.MAX_DIM,
             [f">> the CUDA engine supports at most {_lib.MAX_DIM} spatial dimensions"]),
        )
        for failed, lines in checks:
            if failed:
                raise ValueError(_error_text(lines))

        self.A = np.ascontiguousarray(model_matrix, dtype=float)
        self.y = np.ascontiguousarray(y, dtype=float)
        self.y_err = np.ascontiguousarray(y_err, dtype=float)
        self.x = np.ascontiguousarray(parameter_spatial_positions, dtype=float)

        self.cov = as_engine_covariance(prior_covariance_function)
        self.cov.pass_spatial_data(self.x)
        if self.cov.bounds is None:
            self.cov.bounds = [(None, None)] * self.cov.n_params

        self.mean = as_engine_mean(prior_mean_function)
        self.mean.pass_spatial_data(self.x)
        if self.mean.bounds is None:
            self.mean.bounds = [(None, None)] * self.mean.n_params

        self.n_hyperpars = self.mean.n_params + self.cov.n_params
        self.mean_slice = slice(0, self.mean.n_params)
        self.cov_slice = slice(self.mean.n_params, self.n_hyperpars)
        self.hyperpar_labels = [*self.mean.hyperpar_labels, *self.cov.hyperpar_labels]

        self._device = device
        self._engine = None

    # ------------------------------------------------------------------ engine plumbing
    @property
    def engine(self) -> _lib.Engine:
        if self._engine is None:
            eng = _lib.Engine(self._device)
            # the engine's "training inputs" are the parameter positions; its y is unused on this path
            eng.set_data(self.x, np.zeros(self.x.shape[0]), None, None)
            eng.set_model(self.cov.kinds(), self.mean.kind, self.cov.engine_layout())
            if eng.n_mean + eng.n_cov != self.n_hyperpars:
                raise RuntimeError("engine / host hyper-parameter layout mismatch")
            eng.linv_set_problem(self.A, self.y, self.y_err)
            self._engine = eng
        return self._engine

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None
        return state

    def _theta(self, theta) -> ndarray:
        theta = np.asarray(theta, dtype=float)
        if theta.size != self.n_hyperpars:
            raise ValueError(_error_text([f">> There are a total of {self.n_hyperpars} hyper-parameters,",
                                          f">> but {theta.size} values were given."]))
        return theta

    # ------------------------------------------------------------------ reference API
    def calculate_posterior(self, theta: ndarray):
        """Posterior mean and covariance for the given hyper-parameters (inversion.py:138-155)."""
        mean, cov, info = self.engine.linv_posterior(self._theta(theta), want_cov=True)
        if info != 0:
            raise LinAlgError("Matrix is not positive definite")
        return mean, cov

    def calculate_posterior_mean(self, theta: ndarray) -> ndarray:
        """Posterior mean only (inversion.py:157-168); skips the n x n covariance solve and copy."""
        mean, _, info = self.engine.linv_posterior(self._theta(theta), want_cov=False)
        if info != 0:
            raise LinAlgError("Matrix is not positive definite")
        return mean

    def marginal_likelihood(self, theta: ndarray) -> float:
        """Log-marginal likelihood (inversion.py:170-187); raises ``LinAlgError`` where numpy's cholesky would."""
        val, info = self.engine.linv_lml(self._theta(theta))
        if info != 0:
            raise LinAlgError("Matrix is not positive definite")
        return val

    def marginal_likelihood_gradient(self, theta: ndarray):
        """Log-marginal likelihood and its gradient (inversion.py:189-217)."""
        val, grad, info = self.engine.linv_lml_grad(self._theta(theta))
        if info != 0:
            raise LinAlgError("Matrix is not positive definite")
        return val, grad

    def optimize_hyperparameters(self, initial_guess: ndarray) -> ndarray:
        """Maximise the marginal likelihood with bounded Nelder-Mead (inversion.py:219-249); every objective
        evaluation is one engine call."""
        initial_guess = np.asarray(initial_guess, dtype=float)
        if initial_guess.size != self.n_hyperpars:
            raise ValueError(_error_text([f">> There are a total of {self.n_hyperpars} hyper-parameters,",
                                          f">> but {initial_guess.size} values were given in 'initial_guess'."]))
        hp_bounds = [*self.mean.bounds, *self.cov.bounds]
        result = minimize(
            fun=lambda t: -self.marginal_likelihood(t),
            x0=initial_guess,
            method="Nelder-Mead",
            bounds=hp_bounds,
        )
        return result.x
