"""Does the INT8 kernel's operand stream depend on the digit planes' row stride (= k extent of the compact planes)?
8192-byte rows are a power of two: every row of a k-block starts in the same L2 slice pattern.  Times gemm_nt_i8 at
k = 8192 against k = 8192 +- 64 / 128 with the timing probes (0 normal, 1 loads only, 2 MMAs only).  gpurun."""
import ctypes, json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from inference_tools_b200 import _lib
lib = _lib.load_test_library()
lib.gpb_last_error.restype = ctypes.c_char_p
dp = ctypes.POINTER(ctypes.c_double)
P = lambda a: a.ctypes.data_as(dp) if a is not None else None
rng = np.random.default_rng(0)
out = {}
for (M, N, K) in [(8192, 8192, 8192), (8192, 8192, 8256), (8192, 8192, 8128), (8192, 8192, 8320), (28416, 4096, 8192), (28416, 4096, 8256)]:
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); D = np.zeros((M, N))
    row = {}
    for dbg in (0, 1, 2):
        _lib.set_option("gemm_i8_debug", dbg)
        ms = ctypes.c_double(0)
        r = lib.gpb_test_gemm_impl(1, M, N, K, P(A), P(B), None, ctypes.c_double(1.0), ctypes.c_double(0.0), 0, P(D), 6, ctypes.byref(ms))
        if r:
            raise RuntimeError(lib.gpb_last_error().decode())
        row[f"dbg{dbg}_ms"] = round(ms.value, 4)
        row[f"dbg{dbg}_ms_per_8192k"] = round(ms.value * 8192 / K, 4)
    _lib.set_option("gemm_i8_debug", 0)
    out[f"{M}x{N}x{K}"] = row
    print(M, N, K, row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/i8_stride_probe.json", "w"), indent=1)
