"""``GpOptimiser`` with the interface of ``inference.gp.optimisation.GpOptimiser`` (reference
optimisation.py:14-249; the matplotlib ``plot_results`` helper is out of scope).  The regressor it wraps
and the acquisition evaluations run in the CUDA engine; the proposal search is host-side scipy as in the
reference, plus a batched alternative (``optimizer="sweep"``, BASELINE.json config 4): one dense
acquisition sweep over millions of candidates on the GPU followed by L-BFGS-B restarts from the best ones.
"""
from __future__ import annotations

from collections.abc import Sequence
from inspect import isclass

import numpy as np
from scipy.optimize import differential_evolution, fmin_l_bfgs_b

from inference_tools_b200.gp.acquisition import AcquisitionFunction, ExpectedImprovement
from inference_tools_b200.gp.covariance import CovarianceFunction, SquaredExponential
from inference_tools_b200.gp.mean import ConstantMean, MeanFunction
from inference_tools_b200.gp.regression import GpRegressor
from inference_tools_b200.gp._lockstep import lockstep_lbfgs


class GpOptimiser:
    """Gaussian-process (Bayesian) optimisation: maximise an expensive function from few evaluations.

    Arguments as in the reference (optimisation.py:80-93): ``x, y, bounds, y_err, hyperpars, kernel, mean,
    cross_val, acquisition, optimizer, n_processes``; extra ``device``, ``sweep_candidates``, ``sweep_restarts``,
    ``refit_every`` (re-optimise the hyper-parameters only every k-th added evaluation, appending in between).
    """

    def __init__(
        self,
        x: np.ndarray,
        y: np.ndarray,
        bounds: Sequence,
        y_err: np.ndarray = None,
        hyperpars: np.ndarray = None,
        kernel: CovarianceFunction = SquaredExponential,
        mean: MeanFunction = ConstantMean,
        cross_val: bool = False,
        acquisition: AcquisitionFunction = ExpectedImprovement,
        optimizer: str = "bfgs",
        n_processes: int = 1,
        device: int = None,
        sweep_candidates: int = 1 << 20,
        sweep_restarts: int = 64,
        refit_every: int = 1,
    ):
        self.x = x if isinstance(x, np.ndarray) else np.array(x)
        if self.x.ndim == 1:
            self.x = self.x.reshape(self.x.size, 1)
        self.y = y if isinstance(y, np.ndarray) else np.array(y)
        self.y_err = y_err if isinstance(y_err, (np.ndarray, type(None))) else np.array(y_err)
        self.bounds = bounds
        self.kernel = kernel
        self.mean = mean
        self.cross_val = cross_val
        self.n_processes = n_processes
        self.optimizer = optimizer
        self.device = device
        self.sweep_candidates = sweep_candidates
        self.sweep_restarts = sweep_restarts
        self.refit_every = max(1, int(refit_every))
        self._since_refit = 0
        self._fit_optimizer = optimizer if optimizer in ("bfgs", "diffev") else "bfgs"

        self.gp = self._build_gp(hyperpars)
        self.acquisition = acquisition() if isclass(acquisition) else acquisition
        self.acquisition.update_gp(self.gp)

        self.acquisition_max_history = []
        self.convergence_metric_history = []
        self.iteration_history = []

    def _build_gp(self, hyperpars=None):
        return GpRegressor(
            x=self.x, y=self.y, y_err=self.y_err, hyperpars=hyperpars, kernel=self.kernel,
            mean=self.mean, cross_val=self.cross_val, optimizer=self._fit_optimizer,
            n_processes=self.n_processes, device=self.device,
        )

    def __call__(self, x):
        return self.gp(x)

    def add_evaluation(self, new_x: np.ndarray, new_y: np.ndarray, new_y_err: np.ndarray = None):
        """Append an evaluation and re-fit the regressor (optimisation.py:136-190)."""
        new_x = np.array(new_x, dtype=float).reshape(1, self.x.shape[1])
        new_y = new_y if isinstance(new_y, np.ndarray) else np.array(new_y)
        if not isinstance(new_y_err, (np.ndarray, type(None))):
            new_y_err = np.array(new_y_err)

        self.acquisition_max_history.append(self.acquisition(new_x))
        self.convergence_metric_history.append(self.acquisition.convergence_metric(new_x))
        self.iteration_history.append(self.y.size + 1)

        self.x = np.append(self.x, new_x, axis=0)
        self.y = np.append(self.y, new_y)
        if self.y_err is not None:
            if new_y_err is None:
                raise ValueError(
                    """\n
                    \r[ GpOptimiser error ]
                    \r>> 'new_y_err' argument of the 'add_evaluation' method must be
                    \r>> specified if the 'y_err' argument was specified when the
                    \r>> instance of GpOptimiser was initialised.
                    """
                )
            self.y_err = np.append(self.y_err, new_y_err)

        # The reference re-optimises the hyper-parameters and re-factors from scratch after every evaluation
        # (optimisation.py:177-186).  refit_every = k > 1 (extension) does that only every k-th evaluation; in between the
        # new point is appended to the existing factor at the current hyper-parameters: one O(N^2) row update on the GPU.
        self._since_refit += 1
        if self._since_refit < self.refit_every:
            self.gp.append(new_x, new_y, new_y_err)
        else:
            self._since_refit = 0
            self.gp = self._build_gp()
        self.mu_max = self.y.max()
        self.acquisition.update_gp(self.gp)

    # ------------------------------------------------------------------ proposal search
    def diff_evo(self):
        res = differential_evolution(self.acquisition.opt_func, self.bounds, popsize=30)
        val = res.fun[0] if hasattr(res.fun, "__len__") else res.fun
        return res.x, val

    def launch_bfgs(self, x0: np.ndarray):
        return fmin_l_bfgs_b(self.acquisition.opt_func_gradient, x0, approx_grad=False, bounds=self.bounds, pgtol=1e-10)

    def multistart_bfgs(self, starting_positions=None):
        """optimisation.py:211-223.  The restarts run scipy's L-BFGS-B as in the reference, but in lockstep: their
        evaluation requests are answered by ONE batched device call per round (``_lockstep.py``) instead of one
        single-point prediction each; every restart follows the trajectory it would follow alone."""
        if starting_positions is None:
            starting_positions = self.acquisition.starting_positions(self.bounds)
        starts = list(starting_positions)
        batch = getattr(self.acquisition, "opt_func_gradient_batch", None)
        if batch is not None and len(starts) > 1:
            results = lockstep_lbfgs(batch, starts, self.bounds, pgtol=1e-10)
        else:
            results = [self.launch_bfgs(x0) for x0 in starts]
        best = sorted(results, key=lambda r: float(r[1]))[0]
        return best[0], float(best[1])

    def sweep(self, n_candidates: int = None, n_restarts: int = None, rng=None):
        """Dense batched search: evaluate ``opt_func`` (-ln acquisition) at ``n_candidates`` uniform points in one
        GPU call, then polish the ``n_restarts`` best with L-BFGS-B (analytic gradient when the acquisition has
        one for this kernel, else the sweep optimum is returned unpolished)."""
        n_candidates = n_candidates or self.sweep_candidates
        n_restarts = n_restarts or self.sweep_restarts
        rng = np.random.default_rng() if rng is None else rng
        lwr, upr = (np.array([b[i] for b in self.bounds], dtype=float) for i in (0, 1))
        cand = lwr + (upr - lwr) * rng.random((n_candidates, lwr.size))
        score = self.acquisition.opt_func_batch(cand)
        order = np.argsort(score)[:n_restarts]
        try:
            return self.multistart_bfgs(cand[order])
        except NotImplementedError:
            return cand[order[0]], float(score[order[0]])

    def propose_evaluation(self, optimizer=None):
        """Location of the next evaluation = maximiser of the acquisition function (optimisation.py:225-249)."""
        opt = optimizer if optimizer is not None else self.optimizer
        if opt == "bfgs":
            proposed, _ = self.multistart_bfgs()
        elif opt == "sweep":
            proposed, _ = self.sweep()
        else:
            proposed, _ = self.diff_evo()
        if hasattr(proposed, "__len__") and len(proposed) == 1:
            proposed = proposed[0]
        return proposed
