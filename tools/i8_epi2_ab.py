"""A/B of the INT8 kernel's second-sweep epilogue (option "gemm_i8_epi2": 0 = one pass, chunk by chunk; 1 = two passes: the
accumulator row to registers, TMEM released, then the global-memory part) on short-k shapes (8 epilogue warps).  ms per
call include the operand splitting; D must agree bit for bit.  gpurun."""
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
shapes = [(2048, 2048, 1024, 0), (8192, 8192, 1024, 0), (8192, 8192, 2048, 0), (16384, 16384, 2048, 1), (28416, 1024, 1024, 0),
          (16384, 2048, 2048, 0), (8192, 8192, 704, 0), (8192, 8192, 3584, 0)]
for (M, N, K, fl) in shapes:
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); Cm = rng.standard_normal((M, N))
    row, Ds = {}, {}
    for mode in (0, 1):
        _lib.set_option("gemm_i8_epi2", mode)
        D = np.zeros((M, N)); ms = ctypes.c_double(0)
        r = lib.gpb_test_gemm_impl(1, M, N, K, P(A), P(B), P(Cm), ctypes.c_double(-1.0), ctypes.c_double(1.0), fl, P(D), 6, ctypes.byref(ms))
        if r:
            raise RuntimeError(lib.gpb_last_error().decode())
        Ds[mode] = D
        flops = 2.0 * M * N * K * (0.5 * (1 + 256.0 / M) if fl else 1.0)
        row[f"epi2_{mode}_ms"] = ms.value
        row[f"epi2_{mode}_tflops_fp64_equiv"] = flops / ms.value / 1e9
    rows = slice(0, 256)
    ref = Cm[rows] - A[rows] @ B.T
    mask = np.ones((256, N), bool)
    if fl:
        mask[:, 256:] = False            # lower launches compute only the tiles that touch the lower triangle
    row["err_vs_numpy"] = float(np.abs((Ds[1][rows] - ref) * mask).max() / (np.abs(A[rows]) @ np.abs(B).T).max())
    row["layouts_bit_identical"] = bool(np.array_equal(Ds[0], Ds[1]))
    row["speedup"] = row["epi2_0_ms"] / row["epi2_1_ms"]
    out[f"{M}x{N}x{K}" + ("_lower" if fl else "")] = row
    print(M, N, K, fl, row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/i8_epi2_ab.json", "w"), indent=1)
