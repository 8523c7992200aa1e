"""Acquisition functions with the interface of ``inference.gp.acquisition`` (reference
acquisition.py:8-232).  All three evaluate on the GPU: the predictive mean / sigma (and their spatial
derivatives) come from the batched solve, and the acquisition formulae -- EI / log-EI including the
``Z < -3`` log-space branch through ``erfcx`` (acquisition.py:79-81, 104-114), UCB (:169-189), MaxVariance
(:213-229) -- run in the ``acquisition_kernel`` epilogue; the best candidate of a batch is found by a device
reduction.  Besides the reference's one-point-per-call methods every class offers ``batch(points)``
for millions of candidates in one call (BASELINE.json config 4).
"""
from __future__ import annotations

import numpy as np
from numpy.random import random

from inference_tools_b200 import _lib
from inference_tools_b200.gp.regression import GpRegressor


class AcquisitionFunction:
    gp: GpRegressor
    mu_max: float

    def starting_positions(self, bounds):
        """Starting points for the acquisition optimiser (acquisition.py:13-37): for every training point
        inside the bounds the best of 20 random neighbours, ranked here by ONE batched ``opt_func`` call."""
        lwr, upr = (np.array([b[i] for b in bounds], dtype=float) for i in (0, 1))
        widths = upr - lwr
        lwr = lwr + widths * 0.01
        upr = upr - widths * 0.01
        L = len(widths)
        starts = [None] * len(self.gp.x)
        cand, owner = [], []
        for i, x0 in enumerate(self.gp.x):
            if ((x0 >= lwr) & (x0 <= upr)).all():
                s = x0[None, :] + 0.02 * widths[None, :] * (2 * random(size=(20, L)) - 1)
                cand.append(np.minimum(upr, np.maximum(lwr, s)))
                owner.append(i)
            else:
                starts[i] = lwr + (upr - lwr) * random(size=L)
        if cand:
            pts = np.concatenate(cand)
            score = self.opt_func_batch(pts).reshape(len(owner), 20)
            best = score.argmin(axis=1)
            for row, i in enumerate(owner):
                starts[i] = cand[row][best[row]]
        return starts

    def update_gp(self, gp: GpRegressor):
        self.gp = gp
        self.mu_max = gp.y.max()

    def opt_func_batch(self, points):
        return np.array([self.opt_func(p) for p in points])


class ExpectedImprovement(AcquisitionFunction):
    r"""EI(x) = (z F(z) + P(z)) sigma(x), z = (mu(x) - y_max) / sigma(x)  (reference acquisition.py:44-140)."""

    def __init__(self):
        self.name = "Expected improvement"
        self.convergence_description = r"$\mathrm{EI}_{\mathrm{max}} \; / \; (y_{\mathrm{max}} - y_{\mathrm{min}})$"

    def _run(self, x, mode):
        p = self.gp.process_points(x)
        return self.gp.engine.expected_improvement(p, self.mu_max, mode)

    def __call__(self, x) -> float:
        return self._run(x, _lib.EI_VALUE)[0][0]

    def opt_func(self, x) -> float:
        return self._run(x, _lib.EI_NEG_LOG)[0][0]

    def opt_func_gradient(self, x):
        val, grad, _ = self._run(x, _lib.EI_NEG_LOG_GRAD)
        return np.array(val[0]), grad[0].squeeze()

    # batched surface (new): values for all candidates in one call + index of the best one
    def batch(self, points, log: bool = False):
        val, _, best = self._run(points, _lib.EI_NEG_LOG if log else _lib.EI_VALUE)
        return (-val if log else val), best

    def opt_func_batch(self, points):
        return self._run(points, _lib.EI_NEG_LOG)[0]

    def opt_func_gradient_batch(self, points):
        val, grad, _ = self._run(points, _lib.EI_NEG_LOG_GRAD)
        return val, grad

    def convergence_metric(self, x):
        return self.__call__(x) / (self.mu_max - self.gp.y.min())


class UpperConfidenceBound(AcquisitionFunction):
    r"""UCB(x) = mu(x) + kappa sigma(x)  (reference acquisition.py:143-192); evaluated by the same fused kernel as EI."""

    def __init__(self, kappa: float = 2.0):
        self.kappa = kappa
        self.name = "Upper confidence bound"
        self.convergence_description = r"$\mathrm{UCB}_{\mathrm{max}} - y_{\mathrm{max}}$"

    def _run(self, x, mode):
        p = self.gp.process_points(x)
        return self.gp.engine.acquisition(_lib.ACQ_UCB, self.kappa, p, mode)

    def __call__(self, x) -> float:
        return self._run(x, _lib.ACQ_VALUE)[0][0]

    def opt_func(self, x) -> float:
        return self._run(x, _lib.ACQ_OPT)[0][0]

    def opt_func_gradient(self, x):
        val, grad, _ = self._run(x, _lib.ACQ_OPT_GRAD)
        return np.array(val[0]), grad[0].squeeze()

    def batch(self, points):
        val, _, best = self._run(points, _lib.ACQ_VALUE)
        return val, best

    def opt_func_batch(self, points):
        return self._run(points, _lib.ACQ_OPT)[0]

    def opt_func_gradient_batch(self, points):
        val, grad, _ = self._run(points, _lib.ACQ_OPT_GRAD)
        return val, grad

    def convergence_metric(self, x):
        return self.__call__(x) - self.mu_max


class MaxVariance(AcquisitionFunction):
    r"""mv(x) = sigma^2(x)  (reference acquisition.py:195-232)."""

    def __init__(self):
        self.name = "Max variance"
        self.convergence_description = r"$\sqrt{\mathrm{Var}\left[x\right]}$"

    def _run(self, x, mode):
        p = self.gp.process_points(x)
        return self.gp.engine.acquisition(_lib.ACQ_MAXVAR, 0.0, p, mode)

    def __call__(self, x) -> float:
        return self._run(x, _lib.ACQ_VALUE)[0][0]

    def opt_func(self, x) -> float:
        return self._run(x, _lib.ACQ_OPT)[0][0]

    def opt_func_gradient(self, x):
        val, grad, _ = self._run(x, _lib.ACQ_OPT_GRAD)
        return np.array(val[0]), grad[0].squeeze()

    def batch(self, points):
        val, _, best = self._run(points, _lib.ACQ_VALUE)
        return val, best

    def opt_func_batch(self, points):
        return self._run(points, _lib.ACQ_OPT)[0]

    def opt_func_gradient_batch(self, points):
        val, grad, _ = self._run(points, _lib.ACQ_OPT_GRAD)
        return val, grad

    def convergence_metric(self, x):
        return np.sqrt(self.__call__(x))
