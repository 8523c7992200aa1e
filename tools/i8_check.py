"""GPU check of the INT8-tensor-core FP64 GEMM (csrc/gemm_i8.cu) against numpy, plus timing against the DMMA kernel.
Run under gpurun; every process is wrapped in a timeout by the caller."""
import ctypes, json, os, subprocess, sys
import numpy as np

def run(mode):
    os.environ["GPB200_GEMM_I8"] = "2" if mode == "i8" else "0"
    os.environ["GPB200_GEMM_I8_MINK"] = "64"
    lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "inference_tools_b200", "libgpb200_test.so"))
    lib.gpb_last_error.restype = ctypes.c_char_p
    dp = ctypes.POINTER(ctypes.c_double)
    P = lambda a: a.ctypes.data_as(dp) if a is not None else None
    def gemm(A, B, C, alpha, beta, flags, reps=0):
        M, K = A.shape; N = B.shape[0]
        D = np.zeros((M, N)); ms = ctypes.c_double(0)
        r = lib.gpb_test_gemm(M, N, K, P(A), P(B), P(C), ctypes.c_double(alpha), ctypes.c_double(beta), flags, P(D), reps, ctypes.byref(ms))
        if r: raise RuntimeError(lib.gpb_last_error().decode())
        return D, ms.value
    rng = np.random.default_rng(0)
    res = {}
    stage = sys.argv[2] if len(sys.argv) > 2 else "all"
    if stage in ("all", "small"):
        for (M, N, K) in [(128, 64, 64), (128, 64, 128), (256, 192, 320), (1024, 1024, 1024)]:
            A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); C = rng.standard_normal((M, N))
            D, _ = gemm(A, B, C, -1.0, 1.0, 0)
            ref = C - A @ B.T; den = np.abs(A) @ np.abs(B).T + np.abs(C)
            res[f"full_{M}x{N}x{K}"] = float((np.abs(D - ref) / den).max())
            D, _ = gemm(A, B, None, 2.0, 0.0, 0)
            res[f"full_b0_{M}x{N}x{K}"] = float((np.abs(D - 2 * A @ B.T) / den).max())
        # wide dynamic range inside rows
        M, N, K = 512, 512, 2048
        A = rng.standard_normal((M, K)) * np.exp(3 * rng.standard_normal((M, K))); B = rng.standard_normal((N, K)) * np.exp(3 * rng.standard_normal((N, K)))
        D, _ = gemm(A, B, None, 1.0, 0.0, 0)
        ref = A @ B.T
        res["wide_range_err_over_absAabsB"] = float((np.abs(D - ref) / (np.abs(A) @ np.abs(B).T)).max())
        res["wide_range_err_over_rowmax"] = float((np.abs(D - ref) / (K * np.abs(A).max(1)[:, None] * np.abs(B).max(1)[None, :])).max())
        # rows scaled very differently + zero rows
        A = rng.standard_normal((M, K)) * np.exp(20 * rng.standard_normal((M, 1))); A[5] = 0; B = rng.standard_normal((N, K)) * np.exp(20 * rng.standard_normal((N, 1))); B[7] = 0
        D, _ = gemm(A, B, None, 1.0, 0.0, 0)
        res["row_scaled_err"] = float((np.abs(D - A @ B.T) / (np.abs(A) @ np.abs(B).T + 1e-300)).max())
        M = N = K = 1024
        A = np.triu(rng.standard_normal((M, K))); C = rng.standard_normal((M, N))
        D, _ = gemm(A, A, C, 1.0, 1.0, 1 | 2 | 4)
        mask = np.tril(np.ones((M, N), bool))
        res["lower_trik_err"] = float(np.abs((D - (C + A @ A.T))[mask]).max())
        Bl = np.tril(rng.standard_normal((N, K))); A = rng.standard_normal((M, K))
        # garbage above the diagonal blocks must not be read: poison it
        Bp = Bl.copy()
        for j in range(N): Bp[j, (j // 128 + 1) * 128:] = 1e30   # the INT8 kernel's tiles are 128 columns wide
        D, _ = gemm(A, Bp, None, 1.0, 0.0, 8)
        res["tril_b_err"] = float(np.abs(D - A @ Bl.T).max())
    if stage in ("all", "time"):
        for (M, N, K, fl) in [(8192, 8192, 8192, 0), (16384, 16384, 2048, 0), (16384, 16384, 1024, 1), (28416, 2048, 2048, 0), (16384, 16384, 16384, 0)]:
            A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); C = rng.standard_normal((M, N))
            D, ms = gemm(A, B, C, -1.0, 1.0, fl, reps=3)
            tiles = (M // 128) * (M // 128 + 1) if fl & 1 else (M // 128) * (N // 64)
            res[f"time_{M}x{N}x{K}_f{fl}"] = {"ms": ms, "tflops_fp64_equiv": tiles * 2.0 * 128 * 64 * K / ms / 1e9}
            i = rng.integers(0, M, 64); j = np.minimum(rng.integers(0, N, 64), i if fl & 1 else N)
            refs = C[i, j] - np.einsum("ik,ik->i", A[i], B[j])
            res[f"spot_{M}x{N}x{K}"] = float(np.abs(D[i, j] - refs).max())
    print(json.dumps({mode: res}, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/i8_check_{mode}_{stage}.json", "w"), indent=1)

if __name__ == "__main__":
    run(sys.argv[1])
