"""One launch of gemm_i8_kernel per shape the step issues, for `ncu --set full` (never a bench number):
the predict update (28416 x 4096 x 8192, beta = 1), a lower-tile SYRK-shaped update (16384 x 16384 x 2048, GEMM_LOWER)
and the square 8192^3 launch of round 1's capture."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from inference_tools_b200 import _lib
lib = _lib.load_test_library()
lib.gpb_last_error.restype = ctypes.c_char_p
dp = ctypes.POINTER(ctypes.c_double)
P = lambda a: a.ctypes.data_as(dp) if a is not None else None
rng = np.random.default_rng(0)
for (M, N, K, fl) in [(28416, 4096, 8192, 0), (16384, 16384, 2048, 1), (8192, 8192, 8192, 0)]:
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); C = rng.standard_normal((M, N)); D = np.zeros((M, N))
    r = lib.gpb_test_gemm_impl(1, M, N, K, P(A), P(B), P(C), ctypes.c_double(-1.0), ctypes.c_double(1.0), fl, P(D), 0, None)
    if r:
        raise RuntimeError(lib.gpb_last_error().decode())
    print(M, N, K, fl, "ok", flush=True)
