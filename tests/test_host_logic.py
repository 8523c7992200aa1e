"""CPU: host-side logic of the drop-in layer (no GPU compute): C-ABI exports, descriptor bookkeeping,
bounds, labels, validation errors, and that nothing falls back to the CPU."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden, make_kernel, make_mean
from inference_tools_b200 import _lib
import inference_tools_b200.gp as gp
from inference_tools_b200.gp.covariance import CovarianceFunction, mean_abs_difference, slice_builder


def header_symbols():
    text = open(os.path.join(ROOT, "include", "gpb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpb_[a-z_0-9]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load_library()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gpb200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)


def test_no_cpu_fallback_without_device():
    lib = _lib.load_library()
    import ctypes as C
    cnt = C.c_int(-1)
    rc = lib.gpb_device_count(C.byref(cnt))
    if rc == 0 and cnt.value > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(_lib.EngineError):
        _lib.Engine(0)
    x = np.linspace(0, 1, 10)
    with pytest.raises(_lib.EngineError):
        gp.GpRegressor(x, np.sin(x), hyperpars=[0.0, 0.0, 0.0])


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "inference_tools_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().lower().replace("# oracle", ""), f


def test_mean_abs_difference_identity():
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 200):
        c = rng.normal(size=n)
        brute = np.abs(c[:, None] - c[None, :]).mean()
        assert mean_abs_difference(c) == pytest.approx(brute, rel=1e-12, abs=1e-15)


@pytest.mark.parametrize("name", golden_names())
def test_labels_and_bounds_match_reference(name):
    g = load_golden(name)
    cov, mean = make_kernel(gp, g["comps"]), make_mean(gp, g["mean"])
    cov.pass_spatial_data(g["x"])
    mean.pass_spatial_data(g["x"])
    cov.estimate_hyperpar_bounds(g["y"])
    mean.estimate_hyperpar_bounds(g["y"])
    bounds = np.array([*mean.bounds, *cov.bounds], dtype=float)
    assert np.allclose(bounds, g["bounds"], rtol=1e-11, atol=1e-12)
    if "labels" in g:
        assert [*mean.hyperpar_labels, *cov.hyperpar_labels] == [str(s) for s in g["labels"]]
    assert cov.n_params + mean.n_params == bounds.shape[0]


def test_composite_bookkeeping():
    k = gp.RationalQuadratic() + gp.WhiteNoise() + gp.SquaredExponential()
    x = np.random.default_rng(0).uniform(size=(20, 3))
    k.pass_spatial_data(x)
    assert k.kinds() == [_lib.COV_RQ, _lib.COV_WHITE, _lib.COV_SE]
    assert k.n_params == 5 + 1 + 4
    assert k.hyperpar_labels[0] == "K1: RQ log-amplitude" and k.hyperpar_labels[5] == "K2: WhiteNoise log-sigma"
    assert k.slices == [slice(0, 5), slice(5, 6), slice(6, 10)] == slice_builder([5, 1, 4])
    user_bounds = gp.WhiteNoise(hyperpar_bounds=[(-3.0, 1.0)])
    k2 = gp.SquaredExponential() + user_bounds
    k2.pass_spatial_data(x)
    k2.estimate_hyperpar_bounds(np.arange(20.0))
    assert k2.bounds[-1] == (-3.0, 1.0) and len(k2.bounds) == 5


def test_user_defined_kernels_are_rejected_not_run_on_cpu():
    class Mine(CovarianceFunction):
        def pass_spatial_data(self, x):
            pass

        def estimate_hyperpar_bounds(self, y):
            pass

    x = np.linspace(0, 1, 8)
    with pytest.raises(TypeError, match="no CPU fallback"):
        gp.GpRegressor(x, x, kernel=Mine)
    with pytest.raises(TypeError):
        gp.GpRegressor(x, x, kernel=gp.SquaredExponential() + 3)


def test_input_consistency_checking():
    """the reference's tests/gp/test_GpRegressor.py:154-160 plus the error-argument checks"""
    with pytest.raises(ValueError):
        gp.GpRegressor(x=np.zeros(3), y=np.zeros(2))
    with pytest.raises(ValueError):
        gp.GpRegressor(x=np.zeros([4, 3]), y=np.zeros(3))
    with pytest.raises(ValueError):
        gp.GpRegressor(x=np.zeros([3, 1]), y=np.zeros([3, 2]))
    with pytest.raises(ValueError):
        gp.GpRegressor(x=np.zeros([3, 1, 1]), y=np.zeros(3))
    with pytest.raises(ValueError):
        gp.GpRegressor(x=np.zeros(3), y=np.zeros(3), y_err=np.zeros(4))
    with pytest.raises(TypeError):
        gp.GpRegressor(x=np.zeros(3), y=np.zeros(3), y_err=0.1)
    with pytest.raises(ValueError):
        gp.GpRegressor(x=np.zeros(3), y=np.zeros(3), y_cov=np.zeros([3, 2]))
    with pytest.raises(ValueError):
        gp.GpRegressor(x=np.zeros(3), y=np.zeros(3), y_cov=np.array([[1.0, 0.5, 0], [0, 1, 0], [0, 0, 1]]))
    with pytest.raises(TypeError):
        gp.GpRegressor(x=np.zeros(3), y=np.zeros(3), y_cov="nope")


def test_mean_function_host_helpers():
    rng = np.random.default_rng(5)
    x = rng.uniform(size=(30, 2))
    th = rng.normal(size=5)
    m = gp.QuadraticMean()
    m.pass_spatial_data(x)
    mu, grads = m.mean_and_gradients(th)
    dx = x - x.mean(axis=0)
    assert np.allclose(mu, th[0] + dx @ th[1:3] + dx**2 @ th[3:5])
    assert len(grads) == 5 and np.allclose(grads[3], dx[:, 0] ** 2)
    assert m(x[3], th) == pytest.approx(mu[3])
    lin = gp.LinearMean()
    lin.pass_spatial_data(x)
    assert np.allclose(lin.build_mean(th[:3]), th[0] + dx @ th[1:3])
    assert gp.ConstantMean()(x[0], th) == th[0]


def test_linear_inverter_host_side():
    """GpLinearInverter (reference inversion.py:55-136): argument checks in the reference's order, hyper-parameter
    bookkeeping, and no engine (hence no GPU) before the first evaluation."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "linv_rq_linear.npz"))
    y, y_err, A, x = g["y"], g["y_err"], g["A"], g["x"]
    with pytest.raises(ValueError, match="'model_matrix' argument must be a 2D"):
        gp.GpLinearInverter(y, y_err, A.ravel(), x)
    with pytest.raises(ValueError, match="of equal size"):
        gp.GpLinearInverter(y, y_err[:-1], A, x)
    with pytest.raises(ValueError, match="first dimension of 'model_matrix'"):
        gp.GpLinearInverter(y[:-1], y_err[:-1], A, x)
    with pytest.raises(ValueError, match="'parameter_spatial_positions' must be a 2D"):
        gp.GpLinearInverter(y, y_err, A, x.ravel())
    with pytest.raises(ValueError, match="second dimension of 'model_matrix'"):
        gp.GpLinearInverter(y, y_err, A, x[:-1])
    with pytest.raises(TypeError):
        gp.GpLinearInverter(y, y_err, A, x, prior_covariance_function=object())
    inv = gp.GpLinearInverter(y, y_err, A, x, prior_covariance_function=gp.RationalQuadratic, prior_mean_function=gp.LinearMean)
    assert inv._engine is None
    assert inv.hyperpar_labels == [str(s) for s in g["labels"]]
    assert inv.n_hyperpars == len(g["thetas"][0]) == inv.mean.n_params + inv.cov.n_params
    assert inv.mean_slice == slice(0, inv.mean.n_params) and inv.cov_slice == slice(inv.mean.n_params, inv.n_hyperpars)
    assert inv.cov.bounds == [(None, None)] * inv.cov.n_params and inv.mean.bounds == [(None, None)] * inv.mean.n_params
    with pytest.raises(ValueError, match="hyper-parameters"):
        inv.optimize_hyperparameters(np.ones(inv.n_hyperpars + 1))


def test_distributed_fit_rejects_options_that_cannot_run_in_lockstep():
    """GpRegressor(distributed=...) without hyperpars runs the optimiser on every rank in lockstep (collective likelihood /
    gradient calls): leave-one-out selection and worker threads are refused before any engine is created."""
    x = np.linspace(0, 1, 20)
    y = np.sin(x)
    for kw in ({"n_processes": 2}, {"cross_val": True}):
        with pytest.raises(ValueError, match="distributed=True optimises"):
            gp.GpRegressor(x, y, distributed=(0, 1, None), **kw)
    # the start points of the lockstep optimiser depend on the replicated targets only
    holder = type("Replica", (), {})()
    holder.y = y
    seed = gp.GpRegressor._lockstep_seed(holder)
    holder.y = y.copy()
    assert gp.GpRegressor._lockstep_seed(holder) == seed and 0 <= seed < 2**31
    holder.y = y + 1e-12
    assert gp.GpRegressor._lockstep_seed(holder) != seed
