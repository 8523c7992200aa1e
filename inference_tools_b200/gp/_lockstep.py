"""Many L-BFGS-B restarts, one batched device call per round.

The reference polishes one start point per training point with scipy's ``fmin_l_bfgs_b``, one after the other (or on a
process pool), every function evaluation a single-point prediction (optimisation.py:202-223).  A single-point prediction
on the GPU is pure launch latency, while the batched entry points (``opt_func_gradient_batch``) answer thousands of
points in about the same time.  Here every restart still runs scipy's own L-BFGS-B -- on its own thread, so each one
follows exactly the trajectory it would follow alone -- but the threads' evaluation requests are collected and answered
together: one batched call per round of the slowest restart.
"""
from __future__ import annotations

import threading

import numpy as np
from scipy.optimize import fmin_l_bfgs_b


def lockstep_lbfgs(batch_value_and_grad, x0s, bounds, max_concurrent: int = 256, **lbfgs_kwargs):
    """``[fmin_l_bfgs_b(f, x0, approx_grad=False, bounds=bounds, **lbfgs_kwargs) for x0 in x0s]`` where ``f`` is answered
    through ``batch_value_and_grad(points (R, d)) -> (values (R,), grads (R, d))``.  Results come back in the order of
    ``x0s``.  An exception of the batched call is re-raised here (after every worker has been released)."""
    x0s = [np.asarray(x0, dtype=float) for x0 in x0s]
    out = [None] * len(x0s)
    for lo in range(0, len(x0s), max_concurrent):
        chunk = range(lo, min(lo + max_concurrent, len(x0s)))
        _run_chunk(batch_value_and_grad, x0s, bounds, chunk, out, lbfgs_kwargs)
    return out


def _run_chunk(batch_fn, x0s, bounds, chunk, out, lbfgs_kwargs):
    cond = threading.Condition()
    pending, answers = {}, {}
    state = {"active": len(chunk), "error": None}

    def worker(i):
        def cost(x):
            with cond:
                pending[i] = np.array(x, dtype=float)
                cond.notify_all()
                while i not in answers and state["error"] is None:
                    cond.wait()
                if state["error"] is not None:
                    raise _Released()
                return answers.pop(i)

        try:
            out[i] = fmin_l_bfgs_b(cost, x0s[i], approx_grad=False, bounds=bounds, **lbfgs_kwargs)
        except _Released:
            pass
        finally:
            with cond:
                state["active"] -= 1
                cond.notify_all()

    threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in chunk]
    for t in threads:
        t.start()
    while True:
        with cond:
            while state["active"] > 0 and len(pending) < state["active"]:
                cond.wait()
            if state["active"] == 0:
                break
            ids = sorted(pending)
            pts = np.stack([pending.pop(i) for i in ids])
        try:
            vals, grads = batch_fn(pts)
            vals, grads = np.asarray(vals, dtype=float).reshape(len(ids)), np.asarray(grads, dtype=float).reshape(len(ids), -1)
        except BaseException as exc:  # noqa: BLE001 -- release the workers, then re-raise on the caller's thread
            with cond:
                state["error"] = exc
                cond.notify_all()
            break
        with cond:
            for k, i in enumerate(ids):
                answers[i] = (float(vals[k]), grads[k].copy())
            cond.notify_all()
    for t in threads:
        t.join()
    if state["error"] is not None:
        raise state["error"]


class _Released(Exception):
    """unwinds a worker whose evaluation can no longer be answered"""
