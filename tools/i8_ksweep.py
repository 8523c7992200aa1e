"""K sweep: INT8-path vs DMMA GEMM time at fixed M, N (gpurun)."""
import ctypes, json, os
import numpy as np
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "inference_tools_b200", "libgpb200_test.so"))
lib.gpb_last_error.restype = ctypes.c_char_p
dp = ctypes.POINTER(ctypes.c_double)
P = lambda a: a.ctypes.data_as(dp) if a is not None else None
os.environ["GPB200_GEMM_I8"] = "0"
rng = np.random.default_rng(1)
res = {}
for (M, N) in [(8192, 8192), (28416, 1024), (2048, 2048)]:
    for K in (256, 512, 1024, 2048, 4096):
        A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); C = rng.standard_normal((M, N))
        row = {}
        for impl in (1, 0):
            D = np.zeros((M, N)); ms = ctypes.c_double(0)
            r = lib.gpb_test_gemm_impl(impl, M, N, K, P(A), P(B), P(C), ctypes.c_double(-1.0), ctypes.c_double(1.0), 0, P(D), 5, ctypes.byref(ms))
            if r: raise RuntimeError(lib.gpb_last_error().decode())
            row["i8_ms" if impl else "dmma_ms"] = ms.value
        row["speedup"] = row["dmma_ms"] / row["i8_ms"]
        res[f"{M}x{N}x{K}"] = row
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/i8_ksweep.json", "w"), indent=1)
