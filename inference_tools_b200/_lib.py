"""ctypes binding of libgpb200.so (C ABI: include/gpb200.h).  Fails loudly if the CUDA library is absent."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpb200.so")

COV_SE, COV_RQ, COV_WHITE, COV_HETERO = 0, 1, 2, 3
MEAN_CONST, MEAN_LINEAR, MEAN_QUADRATIC = 0, 1, 2
GET_K_XX, GET_L, GET_ALPHA, GET_MU = 0, 1, 2, 3
EI_VALUE, EI_NEG_LOG, EI_NEG_LOG_GRAD = 0, 1, 2
ACQ_VALUE, ACQ_OPT, ACQ_OPT_GRAD = 0, 1, 2      # the same modes, read for any acquisition kind
ACQ_EI, ACQ_UCB, ACQ_MAXVAR = 0, 1, 2
MAX_DIM, MAX_COMP, MAX_REG = 8, 4, 4

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_ctx_p = C.c_void_p

# name -> (restype, argtypes); every symbol declared in include/gpb200.h
SIGNATURES = {
    "gpb_last_error": (C.c_char_p, []),
    "gpb_device_count": (C.c_int, [_ip]),
    "gpb_launch_count": (C.c_int64, []),
    "gpb_gemm_flops": (C.c_double, []),
    "gpb_gemm_flops_int8": (C.c_double, []),
    "gpb_set_option": (C.c_int, [C.c_char_p, C.c_int64]),
    "gpb_get_option": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64)]),
    "gpb_ctx_create": (C.c_int, [C.c_int, C.POINTER(_ctx_p)]),
    "gpb_ctx_destroy": (None, [_ctx_p]),
    "gpb_set_data": (C.c_int, [_ctx_p, _dp, C.c_int64, C.c_int, _dp, _dp, _dp]),
    "gpb_set_model": (C.c_int, [_ctx_p, _ip, C.c_int, C.c_int]),
    "gpb_set_model_ex": (C.c_int, [_ctx_p, _ip, _ip, _ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gpb_num_hyperpars": (C.c_int, [_ctx_p, _ip, _ip]),
    "gpb_build_covariance": (C.c_int, [_ctx_p, _dp, C.c_int, _dp]),
    "gpb_covariance_and_gradients": (C.c_int, [_ctx_p, _dp, _dp, _dp]),
    "gpb_cross_covariance": (C.c_int, [_ctx_p, _dp, C.c_int64, _dp, C.c_int64, _dp, _dp]),
    "gpb_factor": (C.c_int, [_ctx_p, _dp, _ip]),
    "gpb_get": (C.c_int, [_ctx_p, C.c_int, _dp]),
    "gpb_lml": (C.c_int, [_ctx_p, _dp, _dp, _ip]),
    "gpb_lml_grad": (C.c_int, [_ctx_p, _dp, _dp, _dp, _ip]),
    "gpb_lml_grad_batch": (C.c_int, [C.POINTER(_ctx_p), C.c_int, _dp, C.c_int, C.c_int, _dp, _dp, _ip]),
    "gpb_loo": (C.c_int, [_ctx_p, _dp, _dp, _dp, _ip]),
    "gpb_loo_predictions": (C.c_int, [_ctx_p, _dp, _dp]),
    "gpb_predict": (C.c_int, [_ctx_p, _dp, C.c_int64, _dp, _dp]),
    "gpb_predict_dev": (C.c_int, [_ctx_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "gpb_gradient": (C.c_int, [_ctx_p, _dp, C.c_int64, _dp, _dp]),
    "gpb_spatial_derivatives": (C.c_int, [_ctx_p, _dp, C.c_int64, _dp, _dp]),
    "gpb_posterior": (C.c_int, [_ctx_p, _dp, C.c_int64, _dp, _dp]),
    "gpb_expected_improvement": (C.c_int, [_ctx_p, _dp, C.c_int64, C.c_double, C.c_int, _dp, _dp, C.POINTER(C.c_int64)]),
    "gpb_append_point": (C.c_int, [_ctx_p, _dp, C.c_double, C.c_double, _ip]),
    "gpb_acquisition": (C.c_int, [_ctx_p, C.c_int, C.c_double, _dp, C.c_int64, C.c_int, _dp, _dp, C.POINTER(C.c_int64)]),
    "gpb_dist_unique_id": (C.c_int, [C.c_char_p]),
    "gpb_dist_init": (C.c_int, [_ctx_p, C.c_int, C.c_int, C.c_char_p]),
    "gpb_dist_factor": (C.c_int, [_ctx_p, _dp, C.c_int, _ip, _dp]),
    "gpb_dist_lml": (C.c_int, [_ctx_p, _dp, C.c_int, _dp, _ip, _dp]),
    "gpb_dist_lml_grad": (C.c_int, [_ctx_p, _dp, C.c_int, _dp, _dp, _ip, _dp]),
    "gpb_dist_alpha": (C.c_int, [_ctx_p, _dp]),
    "gpb_dist_predict": (C.c_int, [_ctx_p, _dp, C.c_int64, _dp, _dp]),
    "gpb_dist_finalize": (C.c_int, [_ctx_p]),
    "gpb_dist_plan": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, _ip, _ip, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _ip]),
    "gpb_linv_set_problem": (C.c_int, [_ctx_p, _dp, C.c_int64, _dp, _dp]),
    "gpb_linv_lml": (C.c_int, [_ctx_p, _dp, _dp, _ip]),
    "gpb_linv_lml_grad": (C.c_int, [_ctx_p, _dp, _dp, _dp, _ip]),
    "gpb_linv_posterior": (C.c_int, [_ctx_p, _dp, _dp, _dp, _ip]),
    "gpb_ctx_stat": (C.c_int, [_ctx_p, C.c_char_p, _dp]),
    "gpb_timers": (C.c_int, [_ctx_p, C.c_char_p, C.c_int, _dp, C.c_int, _ip]),
    "gpb_dev_alloc": (C.c_int, [_ctx_p, C.c_int64, C.POINTER(C.c_void_p)]),
    "gpb_dev_free": (C.c_int, [_ctx_p, C.c_void_p]),
    "gpb_dev_upload": (C.c_int, [_ctx_p, C.c_void_p, _dp, C.c_int64]),
    "gpb_dev_download": (C.c_int, [_ctx_p, _dp, C.c_void_p, C.c_int64]),
    "gpb_sync": (C.c_int, [_ctx_p]),
}

_lib = None


class EngineError(RuntimeError):
    pass


def load_library():
    """Load libgpb200.so and attach signatures.  Needs no GPU (symbol check only)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} is missing: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "inference_tools_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_test_lib = None


def load_test_library():
    """libgpb200_test.so: host-buffer hooks onto internal primitives (csrc/api_test.cu) for tests/ and tools/ only."""
    global _test_lib
    if _test_lib is None:
        load_library()  # resolves its libgpb200.so dependency to the already loaded copy
        _test_lib = C.CDLL(os.path.join(_HERE, "libgpb200_test.so"))
    return _test_lib


def set_option(name: str, value: int) -> None:
    """Runtime switch of the library (process-wide): "gemm_i8" 0/1/2, "gemm_i8_min_k", "gemm_i8_pair", "gemm_tile",
    "gemm_tma", "graphs", "i8_fallback", "predict_block" (include/gpb200.h)."""
    lib = load_library()
    if lib.gpb_set_option(name.encode(), int(value)) != 0:
        raise EngineError(lib.gpb_last_error().decode())


def get_option(name: str) -> int:
    lib = load_library()
    v = C.c_int64(0)
    if lib.gpb_get_option(name.encode(), C.byref(v)) != 0:
        raise EngineError(lib.gpb_last_error().decode())
    return v.value


class options:
    """Context manager: ``with _lib.options(gemm_i8=0): ...`` sets options and restores the previous values."""

    def __init__(self, **kw):
        self.new, self.old = kw, {}

    def __enter__(self):
        for k, v in self.new.items():
            self.old[k] = get_option(k)
            set_option(k, v)
        return self

    def __exit__(self, *exc):
        for k, v in self.old.items():
            set_option(k, v)
        return False


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class Engine:
    """One libgpb200 context = one GPU's worth of state for one (x, y, model)."""

    def __init__(self, device: int | None = None):
        self.lib = load_library()
        if device is None:
            device = int(os.environ.get("GPB200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
            cnt = C.c_int(0)
            if self.lib.gpb_device_count(C.byref(cnt)) == 0 and cnt.value > 0:
                device %= cnt.value
        self.device = device
        self._ctx = _ctx_p()
        self._check(self.lib.gpb_ctx_create(device, C.byref(self._ctx)))
        self.n = self.d = 0
        self.n_mean = self.n_cov = 0

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.gpb_last_error().decode()
            if rc == -3:
                raise NotImplementedError(msg)
            raise EngineError(f"libgpb200 error {rc}: {msg}")

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.gpb_ctx_destroy(self._ctx)
            self._ctx = _ctx_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- setup
    def set_data(self, x, y, noise_var=None, y_cov=None):
        x = _f64(x)
        y = _f64(y)
        self.n, self.d = x.shape
        nv = None if noise_var is None else _f64(noise_var)
        yc = None if y_cov is None else _f64(y_cov)
        self._check(self.lib.gpb_set_data(self._ctx, _ptr(x), self.n, self.d, _ptr(y), _ptr(nv), _ptr(yc)))

    def set_model(self, kinds, mean_kind, layout=None):
        """kinds: leaf kernel kinds in sum order.  layout (optional, from CovarianceFunction.layout()): dict with
        theta_offs, regions, n_regions, cp_axis, cp_theta_off, n_params for models containing a ChangePoint."""
        arr = (C.c_int * len(kinds))(*kinds)
        if layout is None:
            self._check(self.lib.gpb_set_model(self._ctx, arr, len(kinds), mean_kind))
        else:
            offs = (C.c_int * len(kinds))(*layout["theta_offs"])
            regs = (C.c_int * len(kinds))(*layout["regions"])
            self._check(self.lib.gpb_set_model_ex(self._ctx, arr, offs, regs, len(kinds), layout["n_regions"],
                                                  layout["cp_axis"], layout["cp_theta_off"], layout["n_params"], mean_kind))
        a, b = C.c_int(0), C.c_int(0)
        self._check(self.lib.gpb_num_hyperpars(self._ctx, C.byref(a), C.byref(b)))
        self.n_mean, self.n_cov = a.value, b.value

    # ---------------------------------------------------------------- covariance API
    def build_covariance(self, theta_cov, add_sig=False):
        th = _f64(theta_cov)
        out = np.empty((self.n, self.n))
        self._check(self.lib.gpb_build_covariance(self._ctx, _ptr(th), int(add_sig), _ptr(out)))
        return out

    def covariance_and_gradients(self, theta_cov):
        th = _f64(theta_cov)
        k = np.empty((self.n, self.n))
        dk = np.empty((self.n_cov, self.n, self.n))
        self._check(self.lib.gpb_covariance_and_gradients(self._ctx, _ptr(th), _ptr(k), _ptr(dk)))
        return k, dk

    def cross_covariance(self, u, v, theta_cov):
        u, v, th = _f64(u), _f64(v), _f64(theta_cov)
        out = np.empty((u.shape[0], v.shape[0]))
        self._check(self.lib.gpb_cross_covariance(self._ctx, _ptr(u), u.shape[0], _ptr(v), v.shape[0], _ptr(th), _ptr(out)))
        return out

    # ---------------------------------------------------------------- fit / objective
    def factor(self, theta):
        th = _f64(theta)
        info = C.c_int(0)
        self._check(self.lib.gpb_factor(self._ctx, _ptr(th), C.byref(info)))
        return info.value

    def get(self, which):
        shape = (self.n, self.n) if which in (GET_K_XX, GET_L) else (self.n,)
        out = np.empty(shape)
        self._check(self.lib.gpb_get(self._ctx, which, _ptr(out)))
        return out

    def lml(self, theta):
        th = _f64(theta)
        val, info = C.c_double(0), C.c_int(0)
        self._check(self.lib.gpb_lml(self._ctx, _ptr(th), C.byref(val), C.byref(info)))
        return val.value, info.value

    def lml_grad(self, theta):
        th = _f64(theta)
        val, info = C.c_double(0), C.c_int(0)
        grad = np.empty(self.n_mean + self.n_cov)
        self._check(self.lib.gpb_lml_grad(self._ctx, _ptr(th), C.byref(val), _ptr(grad), C.byref(info)))
        return val.value, grad, info.value

    def loo(self, theta, want_grad):
        th = _f64(theta)
        val, info = C.c_double(0), C.c_int(0)
        grad = np.empty(self.n_mean + self.n_cov) if want_grad else None
        self._check(self.lib.gpb_loo(self._ctx, _ptr(th), C.byref(val), _ptr(grad), C.byref(info)))
        return val.value, grad, info.value

    def loo_predictions(self):
        mu, sig = np.empty(self.n), np.empty(self.n)
        self._check(self.lib.gpb_loo_predictions(self._ctx, _ptr(mu), _ptr(sig)))
        return mu, sig

    # ---------------------------------------------------------------- prediction family
    def predict(self, q):
        q = _f64(q)
        m = q.shape[0]
        mu, sig = np.empty(m), np.empty(m)
        self._check(self.lib.gpb_predict(self._ctx, _ptr(q), m, _ptr(mu), _ptr(sig)))
        return mu, sig

    def gradient(self, q):
        q = _f64(q)
        m = q.shape[0]
        mean, cov = np.empty((m, self.d)), np.empty((m, self.d, self.d))
        self._check(self.lib.gpb_gradient(self._ctx, _ptr(q), m, _ptr(mean), _ptr(cov)))
        return mean, cov

    def spatial_derivatives(self, q):
        q = _f64(q)
        m = q.shape[0]
        dmu, dvar = np.empty((m, self.d)), np.empty((m, self.d))
        self._check(self.lib.gpb_spatial_derivatives(self._ctx, _ptr(q), m, _ptr(dmu), _ptr(dvar)))
        return dmu, dvar

    def posterior(self, q, mean_only=False):
        q = _f64(q)
        m = q.shape[0]
        mu = np.empty(m)
        sigma = None if mean_only else np.empty((m, m))
        self._check(self.lib.gpb_posterior(self._ctx, _ptr(q), m, _ptr(mu), _ptr(sigma)))
        return mu, sigma

    def expected_improvement(self, q, y_max, mode=EI_VALUE):
        q = _f64(q)
        m = q.shape[0]
        out = np.empty(m)
        grad = np.empty((m, self.d)) if mode == EI_NEG_LOG_GRAD else None
        best = C.c_int64(-1)
        self._check(self.lib.gpb_expected_improvement(self._ctx, _ptr(q), m, float(y_max), mode, _ptr(out), _ptr(grad), C.byref(best)))
        return out, grad, best.value

    def append_point(self, x_new, y_new, noise_var_new=0.0):
        x_new = _f64(x_new).reshape(self.d)
        info = C.c_int(0)
        self._check(self.lib.gpb_append_point(self._ctx, _ptr(x_new), float(y_new), float(noise_var_new), C.byref(info)))
        if info.value == 0:
            self.n += 1
        return info.value

    def acquisition(self, kind, param, q, mode=ACQ_VALUE):
        """(values, gradient or None, index of the best candidate) for ACQ_EI (param = y_max), ACQ_UCB (kappa), ACQ_MAXVAR"""
        q = _f64(q)
        m = q.shape[0]
        out = np.empty(m)
        grad = np.empty((m, self.d)) if mode == ACQ_OPT_GRAD else None
        best = C.c_int64(-1)
        self._check(self.lib.gpb_acquisition(self._ctx, kind, float(param), _ptr(q), m, mode, _ptr(out), _ptr(grad), C.byref(best)))
        return out, grad, best.value

    # ---------------------------------------------------------------- device-resident path + timers
    def dev_alloc(self, n_doubles):
        p = C.c_void_p()
        self._check(self.lib.gpb_dev_alloc(self._ctx, int(n_doubles), C.byref(p)))
        return p

    def dev_free(self, p):
        self._check(self.lib.gpb_dev_free(self._ctx, p))

    def dev_upload(self, p, host):
        host = _f64(host)
        self._check(self.lib.gpb_dev_upload(self._ctx, p, _ptr(host), host.size))

    def dev_download(self, p, n_doubles):
        out = np.empty(int(n_doubles))
        self._check(self.lib.gpb_dev_download(self._ctx, _ptr(out), p, out.size))
        return out

    def predict_dev(self, q_dev, m, mu_dev, sig_dev):
        self._check(self.lib.gpb_predict_dev(self._ctx, q_dev, int(m), mu_dev, sig_dev))

    def sync(self):
        self._check(self.lib.gpb_sync(self._ctx))

    # ---------------------------------------------------------------- distributed Cholesky (one process per GPU)
    def dist_init(self, rank: int, world: int, unique_id: bytes | None):
        uid = C.create_string_buffer(unique_id if unique_id else b"\0" * 128, 128)
        self._check(self.lib.gpb_dist_init(self._ctx, rank, world, uid))

    def dist_lml(self, theta, block: int = 1024):
        th = _f64(theta)
        val, info = C.c_double(0), C.c_int(0)
        secs = (C.c_double * 3)()
        self._check(self.lib.gpb_dist_lml(self._ctx, _ptr(th), block, C.byref(val), C.byref(info), secs))
        return val.value, info.value, {"assemble_s": secs[0], "factor_s": secs[1], "total_s": secs[2]}

    def dist_factor(self, theta, block: int = 1024):
        th = _f64(theta)
        info = C.c_int(0)
        secs = (C.c_double * 3)()
        self._check(self.lib.gpb_dist_factor(self._ctx, _ptr(th), block, C.byref(info), secs))
        return info.value, {"assemble_s": secs[0], "factor_s": secs[1], "total_s": secs[2]}

    def dist_lml_grad(self, theta, block: int = 1024):
        """collective: log marginal likelihood and its gradient on the block-column-cyclic layout"""
        th = _f64(theta)
        val, info = C.c_double(0), C.c_int(0)
        grad = np.zeros(self.n_mean + self.n_cov)
        secs = (C.c_double * 4)()
        self._check(self.lib.gpb_dist_lml_grad(self._ctx, _ptr(th), block, C.byref(val), _ptr(grad), C.byref(info), secs))
        return val.value, grad, info.value, {"assemble_s": secs[0], "factor_s": secs[1], "total_s": secs[2], "gradient_s": secs[3]}

    def dist_alpha(self):
        out = np.empty(self.n)
        self._check(self.lib.gpb_dist_alpha(self._ctx, _ptr(out)))
        return out

    def dist_predict(self, q):
        """collective: this rank's slab of query points (may be empty) against the sharded factor"""
        q = _f64(q).reshape(-1, self.d)
        m = q.shape[0]
        mu, sig = np.empty(m), np.empty(m)
        self._check(self.lib.gpb_dist_predict(self._ctx, _ptr(q) if m else None, m, _ptr(mu) if m else None, _ptr(sig) if m else None))
        return mu, sig

    def dist_finalize(self):
        self._check(self.lib.gpb_dist_finalize(self._ctx))

    # ---------------------------------------------------------------- linear inversion (inversion.py)
    def linv_set_problem(self, A, y, y_err):
        A, y, y_err = _f64(A), _f64(y), _f64(y_err)
        if A.ndim != 2 or A.shape[1] != self.n or y.shape != (A.shape[0],) or y_err.shape != y.shape:
            raise ValueError("linv_set_problem: A must be (m, n) with n the number of positions; y, y_err length m")
        self.linv_m = A.shape[0]
        self._check(self.lib.gpb_linv_set_problem(self._ctx, _ptr(A), A.shape[0], _ptr(y), _ptr(y_err)))

    def linv_lml(self, theta):
        th = _f64(theta)
        val, info = C.c_double(0), C.c_int(0)
        self._check(self.lib.gpb_linv_lml(self._ctx, _ptr(th), C.byref(val), C.byref(info)))
        return val.value, info.value

    def linv_lml_grad(self, theta):
        th = _f64(theta)
        val, info = C.c_double(0), C.c_int(0)
        grad = np.empty(self.n_mean + self.n_cov)
        self._check(self.lib.gpb_linv_lml_grad(self._ctx, _ptr(th), C.byref(val), _ptr(grad), C.byref(info)))
        return val.value, grad, info.value

    def linv_posterior(self, theta, want_cov=True):
        th = _f64(theta)
        info = C.c_int(0)
        mean = np.empty(self.n)
        cov = np.empty((self.n, self.n)) if want_cov else None
        self._check(self.lib.gpb_linv_posterior(self._ctx, _ptr(th), _ptr(mean), _ptr(cov), C.byref(info)))
        return mean, cov, info.value

    def timers(self):
        names = C.create_string_buffer(1024)
        ms = (C.c_double * 64)()
        n = C.c_int(0)
        self._check(self.lib.gpb_timers(self._ctx, names, 1024, ms, 64, C.byref(n)))
        keys = names.value.decode().split(";") if n.value else []
        return {k: ms[i] for i, k in enumerate(keys)}

    def stat(self, name: str) -> float:
        v = C.c_double(0)
        self._check(self.lib.gpb_ctx_stat(self._ctx, name.encode(), C.byref(v)))
        return v.value

    def launch_count(self):
        return int(self.lib.gpb_launch_count())

    def gemm_flops(self):
        return float(self.lib.gpb_gemm_flops())

    def gemm_flops_int8(self):
        return float(self.lib.gpb_gemm_flops_int8())


def lml_grad_batch(engines, thetas, want_grad=True):
    """(lml[R], grad[R, p] or None, info[R]) for R hyper-parameter vectors, sharded over `engines` (one per GPU, same data
    and model), each engine driven by its own thread inside the library."""
    lib = load_library()
    thetas = _f64(thetas)
    if thetas.ndim != 2:
        raise ValueError("thetas must be (R, p)")
    r, p = thetas.shape
    ctxs = (_ctx_p * len(engines))(*[e._ctx for e in engines])
    lml = np.empty(r)
    grad = np.empty((r, p)) if want_grad else None
    info = np.zeros(r, dtype=np.int32)
    rc = lib.gpb_lml_grad_batch(ctxs, len(engines), _ptr(thetas), r, p, _ptr(lml), _ptr(grad), info.ctypes.data_as(_ip))
    if rc != 0:
        raise EngineError(f"libgpb200 error {rc}: {lib.gpb_last_error().decode()}")
    return lml, grad, info


def device_count() -> int:
    lib = load_library()
    cnt = C.c_int(0)
    if lib.gpb_device_count(C.byref(cnt)) != 0:
        raise EngineError(lib.gpb_last_error().decode())
    return cnt.value


def nccl_unique_id() -> bytes:
    """128-byte NCCL unique id (call on rank 0 and share it with the other ranks)."""
    lib = load_library()
    buf = C.create_string_buffer(128)
    if lib.gpb_dist_unique_id(buf) != 0:
        raise EngineError(lib.gpb_last_error().decode())
    return buf.raw


def dist_plan(n: int, block: int, world: int, rank: int) -> dict:
    """Block-column-cyclic layout of the distributed Cholesky for one rank (host-only; no GPU needed)."""
    lib = load_library()
    nb, no = C.c_int(0), C.c_int(0)
    pd, sd = C.c_int64(0), C.c_int64(0)
    npad = (n + 127) // 128 * 128
    owners = (C.c_int * ((npad + block - 1) // block + 1))()
    if lib.gpb_dist_plan(n, block, world, rank, C.byref(nb), C.byref(no), C.byref(pd), C.byref(sd), owners) != 0:
        raise EngineError(lib.gpb_last_error().decode())
    return {"n_blocks": nb.value, "n_owned": no.value, "panel_doubles": pd.value, "staging_doubles": sd.value,
            "owners": list(owners[: nb.value])}
