"""``GpRegressor`` with the public surface of ``inference.gp.regression.GpRegressor`` (reference
regression.py:16-612), driven by the CUDA engine in ``libgpb200.so``.

What stays on the host: argument validation, hyper-parameter bookkeeping (bounds, labels, slices) and
the scipy optimisers (a handful of parameters).  Every array operation of the reference -- covariance
assembly, Cholesky, triangular solves, the explicit inverse and the gradient traces, the per-point
prediction loop -- is one C-ABI call into hand-written sm_100a kernels; there is no numpy/LAPACK
fallback.  Dense attributes the reference exposes (``K_xx``, ``L``, ``sig``) are materialised lazily
from device state on first access.
"""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor
from warnings import warn

import numpy as np
from numpy.linalg import LinAlgError
from scipy.optimize import differential_evolution, fmin_l_bfgs_b

from inference_tools_b200 import _lib
from inference_tools_b200.gp.covariance import CovarianceFunction, SquaredExponential, as_engine_covariance
from inference_tools_b200.gp.mean import ConstantMean, MeanFunction, as_engine_mean


class GpRegressor:
    """Gaussian-process regression in one or more dimensions (reference regression.py:16-77).

    :param x: x-data as an array of shape (number of points, number of dimensions), or 1D for 1D data.
    :param y: y-data values as a 1D array.
    :param y_err: Gaussian standard deviations of the y-data as a 1D array (optional).
    :param y_cov: full covariance matrix of the y-data, alternative to ``y_err``.
    :param hyperpars: hyper-parameter values; when omitted they are optimised.
    :param kernel: covariance function class or instance (``SquaredExponential`` by default).
    :param mean: mean function class or instance (``ConstantMean`` by default).
    :param bool cross_val: select hyper-parameters by leave-one-out cross-validation instead of the marginal likelihood.
    :param str optimizer: ``"bfgs"`` (multi-start L-BFGS-B) or ``"diffev"`` (differential evolution).
    :param int n_processes: concurrent L-BFGS-B restarts.  The reference forks a ``multiprocessing.Pool``
        (regression.py:600-601); CUDA state does not survive ``fork``, so restarts run on worker threads, each
        with its own engine context, spread round-robin over the visible GPUs; the threads pull start points from
        a shared queue, so an uneven mix of short and long restarts still balances.  With ``optimizer="diffev"``
        and ``n_processes > 1`` every generation's population is evaluated in one batched call sharded over the
        same contexts (``gpb_lml_grad_batch``).
    :param int n_starts: number of L-BFGS-B starting positions.
    :param int device: CUDA device ordinal (extension; default ``$GPB200_DEVICE`` / ``$LOCAL_RANK`` / 0).
    :param distributed: extension for training sets beyond one GPU's memory (BASELINE config 5).  ``True`` = this
        process is one rank of an initialised ``torch.distributed`` group (one process per GPU); or a tuple
        ``(rank, world, nccl_unique_id_bytes)`` when another channel shares the id.  Every rank constructs the
        regressor with the same data; the covariance matrix is then assembled, factored and solved in a
        block-column-cyclic layout over the ranks (NCCL panel broadcasts, ``csrc/dist.cu``).  In this mode
        ``set_hyperparameters``, ``marginal_likelihood``, ``marginal_likelihood_gradient``, ``alpha`` and
        ``__call__`` are COLLECTIVE calls (every rank makes them in the same order; ``__call__`` takes this rank's
        own slab of query points).  Without ``hyperpars`` the ranks run the optimiser in lockstep: the start points
        (and differential-evolution populations) come from a generator seeded by the replicated data instead of
        numpy's global one, and every evaluation returns the same all-reduced value and gradient on every rank.
    :param int dist_block: panel width of the distributed layout (multiple of 128, default 1024).
    """

    def __init__(
        self,
        x: np.ndarray,
        y: np.ndarray,
        y_err: np.ndarray = None,
        y_cov: np.ndarray = None,
        hyperpars: np.ndarray = None,
        kernel: CovarianceFunction = SquaredExponential,
        mean: MeanFunction = ConstantMean,
        cross_val: bool = False,
        optimizer: str = "bfgs",
        n_processes: int = 1,
        n_starts: int = None,
        device: int = None,
        distributed=False,
        dist_block: int = 1024,
    ):
        self.x = x if isinstance(x, np.ndarray) else np.array(x)
        self.y = (y if isinstance(y, np.ndarray) else np.array(y)).squeeze()
        if self.y.ndim != 1:
            raise ValueError(
                f"""\n
                \r[ GpRegressor error ]
                \r>> 'y' argument must be a 1D array, but instead has shape {self.y.shape}
                """
            )
        self.n_points = self.y.size
        if self.x.ndim == 2:
            self.n_dimensions = self.x.shape[1]
        elif self.x.ndim <= 1:
            self.n_dimensions = 1
            self.x = self.x.reshape([self.x.size, 1])
        else:
            raise ValueError(
                f"""\n
                \r[ GpRegressor Error ]
                \r>> 'x' argument must be a 2D array, but instead has
                \r>> {self.x.ndim} dimensions and shape {self.x.shape}.
                """
            )
        if self.x.shape[0] != self.n_points:
            raise ValueError(
                f"""\n
                \r[ GpRegressor Error ]
                \r>> The first dimension of the 'x' array must be equal in size
                \r>> to the 'y' array.
                \r>> 'x' has shape {self.x.shape}, but 'y' has size {self.y.size}.
                """
            )
        if self.n_dimensions > _lib.MAX_DIM:
            raise ValueError(f"[ GpRegressor error ] the CUDA engine supports at most {_lib.MAX_DIM} spatial dimensions")

        # data-error term: kept as a length-N variance vector (or the dense matrix the caller gave)
        self._noise_var, self._y_cov = self.check_error_data(y_err, y_cov)

        self.cov = as_engine_covariance(kernel)
        self.mean = as_engine_mean(mean)
        self.cov.pass_spatial_data(self.x)
        self.mean.pass_spatial_data(self.x)
        if self.cov.bounds is None:
            self.cov.estimate_hyperpar_bounds(self.y)
        if self.mean.bounds is None:
            self.mean.estimate_hyperpar_bounds(self.y)
        self.hp_bounds = [*self.mean.bounds, *self.cov.bounds]
        self.n_hyperpars = len(self.hp_bounds)
        self.mean_slice = slice(0, self.mean.n_params)
        self.cov_slice = slice(self.mean.n_params, self.n_hyperpars)
        self.hyperpar_labels = [*self.mean.hyperpar_labels, *self.cov.hyperpar_labels]

        self._device = device
        self._dist = self._resolve_distributed(distributed)
        self._dist_block = int(dist_block)
        self._dist_theta = None   # hyper-parameters the sharded factor currently belongs to
        self._engine = None
        self._pool = []
        self._n_processes = n_processes
        self._cache = {}
        self.hyperpars = None

        if cross_val:
            self.model_selector = self.loo_likelihood
            self.model_selector_gradient = self.loo_likelihood_gradient
        else:
            self.model_selector = self.marginal_likelihood
            self.model_selector_gradient = self.marginal_likelihood_gradient

        if hyperpars is None and self._dist is not None and (cross_val or n_processes != 1):
            raise ValueError(
                """\n
                [ GpRegressor error ]
                >> distributed=True optimises the marginal likelihood collectively: every rank runs the
                >> same optimiser on the same (all-reduced) values, so 'cross_val' and 'n_processes' > 1
                >> are not available; pass 'hyperpars' to skip the optimisation.
                """
            )
        if hyperpars is None:
            if optimizer not in ["bfgs", "diffev"]:
                optimizer = "bfgs"
                warn(
                    """
                    An invalid option was passed to the 'optimizer' keyword argument.
                    The default option 'bfgs' was used instead.
                    Valid options are 'bfgs' and 'diffev'.
                    """
                )
            if optimizer == "diffev":
                hyperpars = self.differential_evo()
            else:
                hyperpars = self.multistart_bfgs(n_processes=n_processes, starts=n_starts)

        self.set_hyperparameters(hyperpars)

    # ------------------------------------------------------------------ engine plumbing
    def _new_engine(self, device=None) -> _lib.Engine:
        eng = _lib.Engine(self._device if device is None else device)
        eng.set_data(self.x, self.y, self._noise_var, self._y_cov)
        eng.set_model(self.cov.kinds(), self.mean.kind, self.cov.engine_layout())
        if eng.n_mean + eng.n_cov != self.n_hyperpars:
            raise RuntimeError("engine / host hyper-parameter layout mismatch")
        return eng

    @staticmethod
    def _resolve_distributed(distributed):
        """None (single GPU) or (rank, world, unique id).  ``True`` takes rank / world from torch.distributed and shares
        the NCCL unique id of the engine's own communicator through it (rendezvous plumbing only)."""
        if not distributed:
            return None
        if distributed is True:
            import torch.distributed as dist

            if not dist.is_initialized():
                raise RuntimeError("GpRegressor(distributed=True) needs an initialised torch.distributed process group")
            rank, world = dist.get_rank(), dist.get_world_size()
            box = [_lib.nccl_unique_id() if (rank == 0 and world > 1) else None]
            if world > 1:
                dist.broadcast_object_list(box, src=0)
            return rank, world, box[0]
        rank, world, uid = distributed
        return int(rank), int(world), uid

    def _dist_factor(self, theta):
        """(re)factor the sharded covariance at ``theta`` unless it already belongs to it; collective"""
        theta = np.asarray(theta, dtype=float)
        if self._dist_theta is not None and np.array_equal(self._dist_theta, theta):
            return
        self._dist_theta = None
        info, _ = self._engine.dist_factor(theta, self._dist_block)
        if info > 0:
            raise LinAlgError("Matrix is not positive definite")
        self._dist_theta = theta.copy()

    @property
    def engine(self) -> _lib.Engine:
        if self._engine is None:
            self._engine = self._new_engine()
            if self._dist is not None:
                self._engine.dist_init(*self._dist)
            if self.hyperpars is not None:
                self._factor(self.hyperpars)
        return self._engine

    def _factor(self, theta):
        if self._dist is not None:
            return self._dist_factor(theta)
        info = self._engine.factor(theta)
        if info > 0:
            raise LinAlgError("Matrix is not positive definite")

    def _engine_pool(self, n: int):
        """``n`` engine contexts holding this model, the first being ``self.engine``, the others on the next visible
        GPUs round-robin (more contexts than GPUs share devices).  Kept for the life of the object."""
        first = self.engine
        n_dev = _lib.device_count()
        while len(self._pool) < n - 1:
            self._pool.append(self._new_engine((first.device + len(self._pool) + 1) % n_dev))
        return [first, *self._pool[: n - 1]]

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None
        state["_pool"] = []
        state["_cache"] = {}
        state["_dist_theta"] = None
        return state

    # ------------------------------------------------------------------ reference API
    def __call__(self, points: np.ndarray):
        """Mean and standard deviation of the regression estimate at ``points`` (regression.py:188-216)."""
        p = self.process_points(points)
        if self._dist is not None:      # collective: this rank's slab against the sharded factor
            eng = self.engine
            self._dist_factor(self.hyperpars)
            return eng.dist_predict(p)
        return self.engine.predict(p)

    def set_hyperparameters(self, hyperpars: np.ndarray):
        """Update the hyper-parameters and re-factor the model (regression.py:218-244)."""
        if len(hyperpars) != self.n_hyperpars:
            raise ValueError(
                f"""\n
                [ GpRegressor error ]
                >> An incorrect number of hyper-parameter values were passed via the
                >> 'hyperpars' keyword argument:
                >> There are {self.n_hyperpars} hyper-parameters but {len(hyperpars)} values were given.
                """
            )
        self.hyperpars = hyperpars
        self.mean_hyperpars = self.hyperpars[self.mean_slice]
        self.cov_hyperpars = self.hyperpars[self.cov_slice]
        self._cache = {}
        if self._engine is None:
            self._engine = self._new_engine()
            if self._dist is not None:
                self._engine.dist_init(*self._dist)
        self._factor(np.asarray(hyperpars, dtype=float))

    def append(self, new_x, new_y, new_y_err=None):
        """Extension: add ONE training point with the hyper-parameters kept.  The factor gets one new row from a single
        forward substitution (O(N^2)) instead of the O(N^3) rebuild the reference performs for every added evaluation
        (optimisation.py:177-186); afterwards the object is the regressor of the enlarged data set at the same theta."""
        if self._dist is not None or self._y_cov is not None:
            raise NotImplementedError("append() needs a single-GPU regressor without a dense y_cov")
        new_x = np.asarray(new_x, dtype=float).reshape(1, self.n_dimensions)
        if (self._noise_var is None) != (new_y_err is None):
            raise ValueError("[ GpRegressor error ] 'new_y_err' must be given exactly when the regressor was built with 'y_err'")
        nv = 0.0 if new_y_err is None else float(np.asarray(new_y_err).squeeze()) ** 2
        info = self.engine.append_point(new_x, float(np.asarray(new_y).squeeze()), nv)
        if info > 0:
            raise LinAlgError("Matrix is not positive definite")
        self.x = np.append(self.x, new_x, axis=0)
        self.y = np.append(self.y, float(np.asarray(new_y).squeeze()))
        if self._noise_var is not None:
            self._noise_var = np.append(self._noise_var, nv)
        self.n_points += 1
        self.cov.pass_spatial_data(self.x)
        self.mean.pass_spatial_data(self.x)
        for eng in self._pool:
            eng.close()
        self._pool = []
        self._cache = {}

    # dense attributes of the reference object, fetched from the device on demand
    def _fetch(self, which):
        if which not in self._cache:
            if self._dist is not None:
                eng = self.engine
                if which == _lib.GET_ALPHA:     # collective block back-substitution over the sharded factor
                    self._dist_factor(self.hyperpars)
                    self._cache[which] = eng.dist_alpha()
                elif which == _lib.GET_MU:
                    self._cache[which] = self.mean.build_mean(np.asarray(self.mean_hyperpars, dtype=float))
                else:
                    raise NotImplementedError("K_xx and L are sharded over the ranks in distributed mode; they are not gathered")
            else:
                self._cache[which] = self.engine.get(which)
        return self._cache[which]

    @property
    def K_xx(self):
        return self._fetch(_lib.GET_K_XX)

    @property
    def L(self):
        return self._fetch(_lib.GET_L)

    @property
    def alpha(self):
        return self._fetch(_lib.GET_ALPHA)

    @property
    def mu(self):
        return self._fetch(_lib.GET_MU)

    @property
    def sig(self):
        """The data-error covariance the reference stores densely (regression.py:133, 320-322)."""
        if self._y_cov is not None:
            return self._y_cov
        if self._noise_var is not None:
            return np.diag(self._noise_var)
        return np.zeros([self.n_points, self.n_points])

    def check_error_data(self, y_err, y_cov):
        """Validation of the error arguments (regression.py:246-322); returns (variance vector, dense cov)."""
        if y_cov is not None:
            if type(y_cov) in (list, tuple):
                y_cov = np.array(y_cov).squeeze()
            elif type(y_cov) is not np.ndarray:
                raise TypeError(
                    f"""\n
                    [ GpRegressor error ]
                    >> The 'y_cov' keyword argument should be given as a numpy array:
                    >> Expected type {np.ndarray} but type {type(y_cov)} was given.
                    """
                )
            if y_cov.shape != (self.n_points, self.n_points):
                raise ValueError(
                    """\n
                    [ GpRegressor error ]
                    >> The 'y_cov' keyword argument was passed an array with an incorrect
                    >> shape. 'y_cov' must be a 2D array of shape (N,N), where 'N' is the
                    >> number of given y-data values.
                    """
                )
            if not (y_cov == y_cov.T).all():
                raise ValueError(
                    """\n
                    [ GpRegressor error ]
                    >> The covariance matrix passed to the 'y_cov' keyword argument
                    >> is not symmetric.
                    """
                )
            if y_err is not None:
                warn(
                    """\n
                    [ GpRegressor warning ]
                    >> Only one of the 'y_err' and 'y_cov' keyword arguments should
                    >> be specified. Only the input to 'y_cov' will be used - the
                    >> input to 'y_err' will be ignored.
                    """
                )
            return None, np.ascontiguousarray(y_cov, dtype=float)

        if y_err is not None:
            if type(y_err) in (list, tuple):
                y_err = np.array(y_err).squeeze()
            elif type(y_err) is not np.ndarray:
                raise TypeError(
                    f"""\n
                    [ GpRegressor error ]
                    >> The 'y_err' keyword argument should be given as a numpy array:
                    >> Expected type {np.ndarray} but type {type(y_err)} was given.
                    """
                )
            if y_err.shape != (self.n_points,):
                raise ValueError(
                    """\n
                    [ GpRegressor error ]
                    >> The 'y_err' keyword argument was passed an array with an
                    >> incorrect shape. 'y_err' must be a 1D array of length 'N',
                    >> where 'N' is the number of given y-data values.
                    """
                )
            return np.asarray(y_err, dtype=float) ** 2, None
        return None, None

    def process_points(self, points: np.ndarray) -> np.ndarray:
        """Reshape rules for query points (regression.py:324-349)."""
        x = points if isinstance(points, np.ndarray) else np.array(points)
        if x.ndim <= 1 and self.n_dimensions == 1:
            x = x.reshape([x.size, 1])
        elif x.ndim == 1 and x.size == self.n_dimensions:
            x = x.reshape([1, x.size])
        elif x.ndim > 2:
            raise ValueError(
                f"""\n
                [ GpRegressor error ]
                >> 'points' argument must be a 2D array, but given array
                >> has {x.ndim} dimensions and shape {x.shape}.
                """
            )
        if x.shape[1] != self.n_dimensions:
            raise ValueError(
                f"""\n
                [ GpRegressor error ]
                >> The second dimension of the 'points' array must have size
                >> equal to the number of dimensions of the input data.
                >> The input data have {self.n_dimensions} dimensions but 'points' has shape {x.shape}.
                """
            )
        return x

    def gradient(self, points: np.ndarray):
        """Mean and covariance of the gradient of the regression estimate (regression.py:351-385).
        The covariance keeps the reference's ``R - Q^T Q`` with ``R`` broadcast along rows."""
        p = self.process_points(points)
        mean, cov = self.engine.gradient(p)
        return mean.reshape(p.shape[0], self.n_dimensions, 1).squeeze(), cov.squeeze()

    def spatial_derivatives(self, points: np.ndarray):
        """Gradients of the predictive mean and variance (regression.py:387-419)."""
        p = self.process_points(points)
        dmu, dvar = self.engine.spatial_derivatives(p)
        return dmu.squeeze(), dvar.squeeze()

    def build_posterior(self, points: np.ndarray, mean_only=False):
        """Posterior mean vector and covariance matrix at ``points`` (regression.py:421-449)."""
        v = self.process_points(points)
        mu, sigma = self.engine.posterior(v, mean_only=mean_only)
        return mu if mean_only else (mu, sigma)

    def loo_predictions(self):
        """Leave-one-out predictions (regression.py:451-466)."""
        return self.engine.loo_predictions()

    def loo_likelihood(self, theta: np.ndarray) -> float:
        """Leave-one-out log-likelihood (regression.py:468-487)."""
        val, _, info = self.engine.loo(np.asarray(theta, dtype=float), want_grad=False)
        if info > 0:
            warn("Cholesky decomposition failure in loo_likelihood")
            return -1e50
        return np.float64(val)

    def loo_likelihood_gradient(self, theta: np.ndarray):
        """Leave-one-out log-likelihood and its gradient (regression.py:489-526)."""
        val, grad, info = self.engine.loo(np.asarray(theta, dtype=float), want_grad=True)
        if info > 0:
            raise LinAlgError("Matrix is not positive definite")
        return np.float64(val), grad

    def marginal_likelihood(self, theta: np.ndarray) -> float:
        """Log-marginal likelihood (regression.py:528-542)."""
        if self._dist is not None:      # collective; the sharded factor then belongs to `theta` until it is next needed
            eng = self.engine
            self._dist_theta = None
            val, info, _ = eng.dist_lml(np.asarray(theta, dtype=float), self._dist_block)
            if info == 0:
                self._dist_theta = np.asarray(theta, dtype=float).copy()
        else:
            val, info = self.engine.lml(np.asarray(theta, dtype=float))
        if info > 0:
            warn("Cholesky decomposition failure in marginal_likelihood")
            return -1e50
        return np.float64(val)

    def marginal_likelihood_gradient(self, theta: np.ndarray):
        """Log-marginal likelihood and its gradient (regression.py:544-567); a failed factorisation raises
        ``LinAlgError`` as numpy.linalg.cholesky does at regression.py:555."""
        if self._dist is not None:      # collective: this rank's rows of K^-1, its share of the traces, one all-reduce
            self._dist_theta = None
            val, grad, info, _ = self.engine.dist_lml_grad(np.asarray(theta, dtype=float), self._dist_block)
            if info > 0:
                raise LinAlgError("Matrix is not positive definite")
            self._dist_theta = np.asarray(theta, dtype=float).copy()
            return np.float64(val), grad
        val, grad, info = self.engine.lml_grad(np.asarray(theta, dtype=float))
        if info > 0:
            raise LinAlgError("Matrix is not positive definite")
        return np.float64(val), grad

    def marginal_likelihood_batch(self, thetas: np.ndarray, n_devices: int = None) -> np.ndarray:
        """``marginal_likelihood`` for every row of ``thetas`` (R, p) in one call, sharded over ``n_devices`` GPUs
        (default: all visible).  Failed factorisations give -1e50 as the scalar method does."""
        engines = self._engine_pool(min(n_devices or _lib.device_count(), max(1, len(thetas))))
        val, _, info = _lib.lml_grad_batch(engines, thetas, want_grad=False)
        return np.where(info > 0, -1e50, val)

    def marginal_likelihood_gradient_batch(self, thetas: np.ndarray, n_devices: int = None):
        """``marginal_likelihood_gradient`` for every row of ``thetas``: ``(lml[R], grad[R, p])``."""
        engines = self._engine_pool(min(n_devices or _lib.device_count(), max(1, len(thetas))))
        val, grad, info = _lib.lml_grad_batch(engines, thetas, want_grad=True)
        if (info > 0).any():
            raise LinAlgError("Matrix is not positive definite")
        return val, grad

    # ------------------------------------------------------------------ optimisers (host side)
    def differential_evo(self) -> np.ndarray:
        """regression.py:569-574.  ``n_processes > 1``: scipy's vectorised mode hands over a whole generation
        (p, S) per call and the population is evaluated by one batched C-ABI call sharded over the GPUs
        (population updates are then deferred to the end of a generation instead of immediate)."""
        if self._n_processes > 1 and self.model_selector == self.marginal_likelihood:
            def population_cost(pop):
                pop = np.atleast_2d(np.asarray(pop, dtype=float).T)
                return -self.marginal_likelihood_batch(pop, n_devices=self._n_processes)

            res = differential_evolution(func=population_cost, bounds=self.hp_bounds, vectorized=True, updating="deferred")
            return res.x
        # distributed: every rank must propose the same populations (the evaluations are collective calls)
        seed = self._lockstep_seed() if self._dist is not None else None
        res = differential_evolution(func=lambda t: -self.model_selector(t), bounds=self.hp_bounds, seed=seed)
        return res.x

    def _lockstep_seed(self) -> int:
        """Seed shared by all ranks of a distributed fit without communication: a checksum of the (replicated) targets."""
        import zlib
        return zlib.crc32(np.ascontiguousarray(self.y).tobytes()) & 0x7FFFFFFF

    def bfgs_cost_func(self, theta: np.ndarray):
        val, grad = self.model_selector_gradient(theta)
        return -val, -grad

    def launch_bfgs(self, x0: np.ndarray):
        return fmin_l_bfgs_b(func=self.bfgs_cost_func, x0=x0, approx_grad=False, bounds=self.hp_bounds)

    def multistart_bfgs(self, starts: int = None, n_processes: int = 1):
        """Multi-start L-BFGS-B (regression.py:585-605): ``starts - 1`` uniform draws from numpy's legacy global
        RNG plus the centre of the bounds hyper-cube; the best local optimum wins."""
        if starts is None:
            starts = int(2 * np.sqrt(len(self.hp_bounds))) + 1
        lwr, upr = (np.array([b[i] for b in self.hp_bounds]) for i in (0, 1))
        # distributed: the ranks run this optimiser in lockstep (every evaluation is a collective call), so they must draw
        # the same start points -- from a generator seeded by the replicated data instead of the process-global one
        draw = np.random.RandomState(self._lockstep_seed()).random_sample if self._dist is not None else np.random.random
        x0s = [lwr + (upr - lwr) * draw(size=len(self.hp_bounds)) for _ in range(starts - 1)]
        x0s.append(0.5 * (lwr + upr))

        if n_processes == 1 or len(x0s) == 1:
            results = [self.launch_bfgs(x0) for x0 in x0s]
        else:
            results = self._threaded_restarts(x0s, n_processes)
        return sorted(results, key=lambda r: r[1])[0][0]

    def _threaded_restarts(self, x0s, n_workers):
        """Independent restarts on worker threads: one engine context per thread (devices round-robin, kept in the
        object's pool), start points pulled from a shared queue.  ctypes releases the GIL inside every C-ABI call."""
        n_workers = min(n_workers, len(x0s))
        engines = self._engine_pool(n_workers)
        use_loo = self.model_selector_gradient == self.loo_likelihood_gradient
        results = [None] * len(x0s)
        lock = threading.Lock()
        next_start = [0]

        def worker(eng):
            def cost(theta):
                th = np.asarray(theta, dtype=float)
                if use_loo:
                    val, grad, info = eng.loo(th, want_grad=True)
                else:
                    val, grad, info = eng.lml_grad(th)
                if info > 0:
                    raise LinAlgError("Matrix is not positive definite")
                return -val, -grad

            while True:
                with lock:
                    i = next_start[0]
                    next_start[0] += 1
                if i >= len(x0s):
                    return
                results[i] = fmin_l_bfgs_b(func=cost, x0=x0s[i], approx_grad=False, bounds=self.hp_bounds)

        with ThreadPoolExecutor(n_workers) as pool:
            for f in [pool.submit(worker, e) for e in engines]:
                f.result()
        return results

    def __str__(self):
        pad = max(len(label) for label in self.hyperpar_labels) + 2
        lines = ["\n[ GpRegressor hyperparameters ]\n"]
        lines.extend(f"{label:>{pad}} = {val:.4}\n" for label, val in zip(self.hyperpar_labels, self.hyperpars))
        return "".join(lines)
