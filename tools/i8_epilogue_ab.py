"""A/B timing of the INT8 GEMM kernel's epilogue configuration (8 vs 16 epilogue warps) over the shapes the engine issues
(gpurun).  Times include the operand splitting of gemm_nt_i8; ms per call, best of the repetition average."""
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
for (M, N, K) in [(28416, 1024, 1024), (16384, 2048, 2048), (8192, 8192, 1024), (8192, 8192, 2048), (8192, 8192, 8192), (28416, 4096, 8192)]:
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); Cm = rng.standard_normal((M, N)); D = np.zeros((M, N))
    row = {}
    for epi in (8, 16):
        _lib.set_option("gemm_i8_epi", epi)
        ms = ctypes.c_double(0)
        r = lib.gpb_test_gemm_impl(1, M, N, K, P(A), P(B), P(Cm), ctypes.c_double(-1.0), ctypes.c_double(1.0), 0, P(D), 8, ctypes.byref(ms))
        if r:
            raise RuntimeError(lib.gpb_last_error().decode())
        row[f"epi{epi}_ms"] = ms.value
        row[f"epi{epi}_tflops_fp64_equiv"] = 2.0 * M * N * K / ms.value / 1e9
        if epi == 16:
            ref = Cm - A[:256] @ B.T if False else None
    err = np.abs(D[:128] - (Cm[:128] - A[:128] @ B.T)).max() / (np.abs(A[:128]) @ np.abs(B).T).max()
    row["err_vs_numpy_rows0_127"] = float(err)
    out[f"{M}x{N}x{K}"] = row
    print(M, N, K, row, flush=True)
json.dump(out, open("gpurun_out/i8_epilogue_ab.json", "w"), indent=1)
