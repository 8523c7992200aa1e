"""A/B of the INT8 kernel's shared-memory layout for CTA pairs (option "gemm_i8_pair": 1 = four uniform 56 KB slots,
2 = 3 A slots + 2 half-size B slots) over shapes the engine issues (gpurun).  ms per call include the operand splitting
of gemm_nt_i8.  The two layouts run the same exact integer arithmetic: D must agree bit for bit."""
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
shapes = [(2048, 2048, 1024, 0), (8192, 8192, 8192, 0), (28416, 4096, 8192, 0), (28416, 1024, 16384, 0), (16384, 16384, 2048, 1),
          (8192, 8192, 1024, 0), (8192, 8192, 704, 0), (28416, 1024, 1024, 0)]
for (M, N, K, fl) in shapes:
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); Cm = rng.standard_normal((M, N))
    row, Ds = {}, {}
    for mode in (1, 2):
        _lib.set_option("gemm_i8_pair", mode)
        D = np.zeros((M, N)); ms = ctypes.c_double(0)
        r = lib.gpb_test_gemm_impl(1, M, N, K, P(A), P(B), P(Cm), ctypes.c_double(-1.0), ctypes.c_double(1.0), fl, P(D), 6, ctypes.byref(ms))
        if r:
            raise RuntimeError(lib.gpb_last_error().decode())
        Ds[mode] = D
        flops = 2.0 * M * N * K * (0.5 * (1 + 256.0 / M) if fl else 1.0)
        row[f"pair{mode}_ms"] = ms.value
        row[f"pair{mode}_tflops_fp64_equiv"] = flops / ms.value / 1e9
    rows = slice(0, 256)
    ref = Cm[rows] - A[rows] @ B.T
    mask = np.ones((256, N), bool)
    if fl:
        mask[:, 256:] = False            # lower launches compute only the tiles that touch the lower triangle
    row["err_vs_numpy"] = float(np.abs((Ds[2][rows] - ref) * mask).max() / (np.abs(A[rows]) @ np.abs(B).T).max())
    row["layouts_bit_identical"] = bool(np.array_equal(Ds[1], Ds[2]))
    row["speedup"] = row["pair1_ms"] / row["pair2_ms"]
    out[f"{M}x{N}x{K}" + ("_lower" if fl else "")] = row
    print(M, N, K, fl, row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/i8_layout_ab.json", "w"), indent=1)
