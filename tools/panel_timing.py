"""Latency of the serial pieces of the distributed sweep's panel step (gpurun): potrf and inverse of an nbd x nbd
diagonal block, through the test hooks of libgpb200_test.so."""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from inference_tools_b200 import _lib
lib = _lib.load_test_library()
dp = C.POINTER(C.c_double)
lib.gpb_test_potrf.argtypes = [C.c_int, dp, dp, C.POINTER(C.c_int), C.c_int, dp]
lib.gpb_test_inverse.argtypes = [C.c_int, dp, dp, dp, C.c_int, dp, dp]
P = lambda a: a.ctypes.data_as(dp)
out = {}
rng = np.random.default_rng(0)
for n in (512, 1024, 2048, 4096):
    x = rng.uniform(0, 1, (n, 2))
    K = np.exp(-0.5 * ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1) / 0.09) + 0.0025 * np.eye(n)
    A = K.copy(); info = C.c_int(0); ms = C.c_double(0)
    lib.gpb_test_potrf(n, P(A), None, C.byref(info), 5, C.byref(ms))
    W = np.zeros((n, n)); Ki = np.zeros((n, n)); ms3 = np.zeros(3)
    lib.gpb_test_inverse(n, P(K), P(W), P(Ki), 0, None, P(ms3))
    out[n] = {"potrf_ms": ms.value, "trtri_ms": ms3[0], "lauum_ms": ms3[1]}
print(json.dumps(out))
json.dump(out, open("gpurun_out/panel_timing.json", "w"), indent=1)
