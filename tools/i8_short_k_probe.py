"""Where does the INT8 kernel's time go at short k?  Times gemm_nt_i8 (splitting included) per shape with the kernel's
timing probes: debug 0 = normal, 1 = loads + epilogue without the MMAs, 2 = MMAs + epilogue without the loads (results
of 1 and 2 are meaningless), for 8 and 16 epilogue warps.  gpurun; output gpurun_out/i8_short_k_probe.json."""
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
PAIR = int(sys.argv[1]) if len(sys.argv) > 1 else 1     # option "gemm_i8_pair" (1 = uniform slots, 2 = wide layout)
_lib.set_option("gemm_i8_pair", PAIR)
for (M, N, K, fl) in [(8192, 8192, 1024, 0), (16384, 16384, 2048, 1), (8192, 8192, 8192, 0)]:
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); Cm = rng.standard_normal((M, N)); D = np.zeros((M, N))
    row = {}
    for beta in (1.0, 0.0):
        for epi in (8, 16):
            for dbg in (0, 1, 2):
                _lib.set_option("gemm_i8_epi", epi)
                _lib.set_option("gemm_i8_debug", dbg)
                ms = ctypes.c_double(0)
                r = lib.gpb_test_gemm_impl(1, M, N, K, P(A), P(B), P(Cm) if beta else None, ctypes.c_double(-1.0), ctypes.c_double(beta), fl, P(D), 6, ctypes.byref(ms))
                if r:
                    raise RuntimeError(lib.gpb_last_error().decode())
                row[f"beta{int(beta)}_epi{epi}_dbg{dbg}_ms"] = round(ms.value, 4)
    _lib.set_option("gemm_i8_debug", 0)
    out[f"{M}x{N}x{K}" + ("_lower" if fl else "")] = row
    print(M, N, K, fl, row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/i8_short_k_probe_pair{PAIR}.json", "w"), indent=1)
