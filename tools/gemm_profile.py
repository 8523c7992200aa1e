"""Large DMMA GEMM launches for ncu (never a bench number): 8192^3 NT, the SYRK shape, the predict shape."""
import ctypes, os, sys
import numpy as np
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "inference_tools_b200", "libgpb200_test.so"))
dp = ctypes.POINTER(ctypes.c_double)
P = lambda a: a.ctypes.data_as(dp) if a is not None else None
rng = np.random.default_rng(0)
for (M, N, K, fl) in [(8192, 8192, 8192, 0), (16384, 16384, 1024, 1), (18944, 128, 16384, 0)]:
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); C = rng.standard_normal((M, N)); D = np.zeros((M, N))
    ms = ctypes.c_double(0)
    rc = lib.gpb_test_gemm(M, N, K, P(A), P(B), P(C), ctypes.c_double(-1.0), ctypes.c_double(1.0), fl, P(D), 2, ctypes.byref(ms))
    print(M, N, K, fl, rc, ms.value, flush=True)
