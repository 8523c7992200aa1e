"""INT8-path GEMM: transposed operands, triangular k ranges, k extents beyond one int32-exact chunk (gpurun)."""
import ctypes, json, os, sys
import numpy as np
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "inference_tools_b200", "libgpb200_test.so"))
lib.gpb_last_error.restype = ctypes.c_char_p
dp = ctypes.POINTER(ctypes.c_double)
P = lambda a: a.ctypes.data_as(dp) if a is not None else None
LOWER, TRIK_A, TRIK_B, TRIL_B, TRIL_A, A_T, B_T = 1, 2, 4, 8, 16, 32, 64
def gemm(impl, A, B, C, alpha, beta, flags, M, N, K, reps=0):
    D = np.zeros((M, N)); ms = ctypes.c_double(0)
    r = lib.gpb_test_gemm_impl(impl, M, N, K, P(np.ascontiguousarray(A)), P(np.ascontiguousarray(B)), P(C), ctypes.c_double(alpha), ctypes.c_double(beta), flags, P(D), reps, ctypes.byref(ms))
    if r: raise RuntimeError(lib.gpb_last_error().decode())
    return D, ms.value
rng = np.random.default_rng(1)
res = {}
M, N, K = 512, 384, 640
A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); C = rng.standard_normal((M, N))
ref = C - A @ B.T; den = np.abs(A) @ np.abs(B).T + np.abs(C)
for name, fl, a, b in [("nt", 0, A, B), ("a_t", A_T, A.T, B), ("b_t", B_T, A, B.T), ("ab_t", A_T | B_T, A.T, B.T)]:
    D, _ = gemm(1, a, b, C, -1.0, 1.0, fl, M, N, K)
    res[name] = float((np.abs(D - ref) / den).max())
# trtri-like: B n-major upper-stored, k >= n ; A lower triangular (k <= m)
n = 1024
W = np.tril(rng.standard_normal((n, n))); L = rng.standard_normal((n, n))
D, _ = gemm(1, L, W, None, 1.0, 0.0, B_T | TRIK_B, n, n, n)       # B(nn,k) = W[k][nn], zero for k < nn
res["b_t_trik_b"] = float(np.abs(D - L @ W).max())
T = rng.standard_normal((n, n))
D, _ = gemm(1, W, T, None, -1.0, 0.0, B_T | TRIL_A, n, n, n)      # A = W lower (k <= m); B(nn,k) = T[k][nn]
res["b_t_tril_a"] = float(np.abs(D + W @ T).max())
# lauum-like: Kinv = W^T W, lower tiles, both transposed, k >= max(i,j); two chunks when n > 16384
D, _ = gemm(1, W, W, None, 1.0, 0.0, A_T | B_T | TRIK_A | TRIK_B | LOWER, n, n, n)
res["lauum_1024"] = float(np.abs(np.tril(D - W.T @ W)).max())
if len(sys.argv) > 1 and sys.argv[1] == "big":
    K2 = 16384 + 4096
    A = rng.standard_normal((256, K2)); B = rng.standard_normal((1280, K2)); C = rng.standard_normal((256, 1280))
    D, _ = gemm(1, A, B, C, -1.0, 1.0, 0, 256, 1280, K2)
    res["k_chunked_20480"] = float((np.abs(D - (C - A @ B.T)) / (np.abs(A) @ np.abs(B).T)).max())
    n = 8192
    W = np.tril(rng.standard_normal((n, n)))
    for impl in (1, 0):
        D, ms = gemm(impl, W, W, None, 1.0, 0.0, A_T | B_T | TRIK_A | TRIK_B | LOWER, n, n, n, reps=2)
        res[f"lauum_8192_impl{impl}_ms"] = ms
    i = rng.integers(0, n, 32); j = np.minimum(rng.integers(0, n, 32), i)
    res["lauum_8192_spot"] = float(np.abs(D[i, j] - np.einsum("ki,ki->i", W[:, i], W[:, j])).max())
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/i8_check2.json", "w"), indent=1)
