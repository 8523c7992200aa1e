"""INT8 kernel with an L2 prefetch running ahead of the operand loads (option "gemm_i8_prefetch" = distance in k-blocks):
ms per gemm_nt_i8 call (splitting included) for the normal kernel (dbg 0) and the loads-only timing probe (dbg 1), plus
a check that the result does not change.  gpurun; output gpurun_out/i8_prefetch_probe.json."""
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
dists = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 2, 4, 8, 16]
for (M, N, K, fl) in [(8192, 8192, 8192, 0), (28416, 4096, 8192, 0), (28416, 1024, 16384, 0), (16384, 16384, 2048, 1), (8192, 8192, 1024, 0)]:
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); Cm = rng.standard_normal((M, N))
    row, D0 = {}, None
    for pf in dists:
        _lib.set_option("gemm_i8_prefetch", pf)
        for dbg in (0, 1):
            _lib.set_option("gemm_i8_debug", dbg)
            D = np.zeros((M, N)); ms = ctypes.c_double(0)
            r = lib.gpb_test_gemm_impl(1, M, N, K, P(A), P(B), P(Cm), ctypes.c_double(-1.0), ctypes.c_double(1.0), fl, P(D), 6, ctypes.byref(ms))
            if r:
                raise RuntimeError(lib.gpb_last_error().decode())
            row[f"pf{pf}_dbg{dbg}_ms"] = round(ms.value, 4)
            if dbg == 0:
                if D0 is None:
                    D0 = D
                else:
                    row[f"pf{pf}_bit_identical"] = bool(np.array_equal(D, D0))
    _lib.set_option("gemm_i8_debug", 0)
    out[f"{M}x{N}x{K}" + ("_lower" if fl else "")] = row
    print(M, N, K, fl, row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/i8_prefetch_probe.json", "w"), indent=1)
