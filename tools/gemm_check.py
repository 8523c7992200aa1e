"""GPU check of the DMMA GEMM primitive against numpy + timing (gpurun only)."""
import ctypes, json, os, sys
import numpy as np
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "inference_tools_b200", "libgpb200_test.so"))
lib.gpb_last_error.restype = ctypes.c_char_p
dp = ctypes.POINTER(ctypes.c_double)
def P(a): return a.ctypes.data_as(dp) if a is not None else None
def gemm(A, B, C, alpha, beta, flags, reps=0):
    M, K = A.shape; N = B.shape[0]
    D = np.zeros((M, N)); ms = ctypes.c_double(0)
    r = lib.gpb_test_gemm(M, N, K, P(A), P(B), P(C), ctypes.c_double(alpha), ctypes.c_double(beta), flags, P(D), reps, ctypes.byref(ms))
    if r: raise RuntimeError(lib.gpb_last_error().decode())
    return D, ms.value
rng = np.random.default_rng(0)
res = {}
for (M, N, K) in [(128, 64, 16), (128, 64, 48), (256, 192, 208), (512, 512, 512)]:
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); C = rng.standard_normal((M, N))
    D, _ = gemm(A, B, C, -1.0, 1.0, 0)
    ref = C - A @ B.T
    res[f"full_{M}x{N}x{K}"] = float(np.abs(D - ref).max())
    D, _ = gemm(A, B, None, 2.0, 0.0, 0)
    res[f"full_b0_{M}x{N}x{K}"] = float(np.abs(D - 2 * A @ B.T).max())
# lower + trik
M = N = K = 512
A = np.triu(rng.standard_normal((M, K))); C = rng.standard_normal((M, N))
D, _ = gemm(A, A, C, 1.0, 1.0, 1 | 2 | 4)
ref = C + A @ A.T
mask = np.tril(np.ones((M, N), bool))      # tile coverage above the diagonal depends on the tile config
res["lower_trik_err"] = float(np.abs((D - ref)[mask]).max())
# tril_b : B lower-triangular
Bl = np.tril(rng.standard_normal((N, K))); A = rng.standard_normal((M, K))
D, _ = gemm(A, Bl, None, 1.0, 0.0, 8)
res["tril_b_err"] = float(np.abs(D - A @ Bl.T).max())
# timing
for (M, N, K, fl) in [(8192, 8192, 8192, 0), (16384, 16384, 512, 0), (16384, 16384, 512, 1), (18944, 128, 16384, 0), (4096, 4096, 4096, 0), (16384, 16384, 1024, 1)]:
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); C = rng.standard_normal((M, N))
    D, ms = gemm(A, B, C, -1.0, 1.0, fl, reps=5)
    tiles = (M // 128) * (M // 128 + 1) if fl & 1 else (M // 128) * (N // 64)
    fl_count = tiles * 2.0 * 128 * 64 * K
    res[f"time_{M}x{N}x{K}_f{fl}"] = {"ms": ms, "tflops": fl_count / ms / 1e9}
    if M <= 8192 and fl == 0:
        i = rng.integers(0, M, 64); j = rng.integers(0, N, 64)
        refs = C[i, j] - np.einsum("ik,ik->i", A[i], B[j])
        res[f"spot_{M}"] = float(np.abs(D[i, j] - refs).max())
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_check.json", "w"), indent=1)
