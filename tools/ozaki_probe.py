"""Feasibility probe for FP64 GEMM emulation on the INT8 tensor cores (Ozaki splitting): library int8 GEMM rate on
this GPU and the accuracy of S-slice splitting on operands shaped like the predict solve (L blocks, solved panels).
Library calls only (torch._int_mm); not part of the product path."""
import json, sys, time
import numpy as np, torch

dev = torch.device("cuda:0")
out = {}

def rate(M, N, K, reps=5):
    a = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
    b = torch.randint(-127, 128, (K, N), dtype=torch.int8, device=dev)
    torch._int_mm(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch._int_mm(a, b)
    e1.record(); torch.cuda.synchronize()
    return 2.0 * M * N * K * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12

out["int_mm_tops"] = {f"{M}x{N}x{K}": rate(M, N, K) for (M, N, K) in [(8192, 8192, 8192), (16384, 16384, 16384), (28416, 16384, 16384), (8192, 8192, 65536)]}
# b given as (K,N) row-major = N-major; also the k-major form via transpose view
a = torch.randint(-127, 128, (8192, 8192), dtype=torch.int8, device=dev)
bt = torch.randint(-127, 128, (8192, 8192), dtype=torch.int8, device=dev)
torch._int_mm(a, bt.t()); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): torch._int_mm(a, bt.t())
e1.record(); torch.cuda.synchronize()
out["int_mm_tops"]["8192^3_kmajorB"] = 2.0 * 8192**3 * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e12

def split(x, S):
    """x (rows, K) float64 -> list of S int8 slices and per-row exponents: x ~ 2^e * sum_s q_s 2^{-7(s+1)}"""
    amax = x.abs().amax(dim=1, keepdim=True).clamp_min(1e-300)
    e = torch.ceil(torch.log2(amax)) + 1  # |x| 2^-e < 0.5
    r = x * torch.exp2(-e)
    qs = []
    for _ in range(S):
        r = r * 128.0
        q = torch.trunc(r)
        r = r - q
        qs.append(q.to(torch.int8))
    return qs, e

def emulated(A, B, S, full=False):
    """A (M,K), B (N,K) float64 -> A B^T"""
    qa, ea = split(A, S)
    qb, eb = split(B, S)
    C = torch.zeros(A.shape[0], B.shape[0], dtype=torch.float64, device=dev)
    G = 2 * S - 2 if full else S - 1
    for g in range(G, -1, -1):
        acc = None
        for s in range(max(0, g - S + 1), min(g, S - 1) + 1):
            p = torch._int_mm(qa[s], qb[g - s].t().contiguous())
            acc = p.to(torch.float64) if acc is None else acc + p.to(torch.float64)
        C += acc * 2.0 ** (-7 * (g + 2))
    return C * torch.exp2(ea) * torch.exp2(eb).t()

torch.manual_seed(0)
n, d, m = 4096, 5, 2048
x = torch.rand(n, d, dtype=torch.float64, device=dev)
q = torch.rand(m, d, dtype=torch.float64, device=dev)
def rq(u, v, a=1.0, alpha=1.5, l=0.4):
    d2 = ((u[:, None, :] - v[None, :, :]) ** 2).sum(-1) / l**2
    return a * a * (1 + 0.5 * d2 / alpha) ** (-alpha)
K = rq(x, x) + torch.eye(n, dtype=torch.float64, device=dev) * 0.05**2
L = torch.linalg.cholesky(K)
Kq = rq(q, x)
h = n // 2
X1 = torch.linalg.solve_triangular(L[:h, :h], Kq[:, :h].t(), upper=False).t().contiguous()  # m x h : solved half
L21 = L[h:, :h].contiguous()
ref = X1 @ L21.t()
den = X1.abs() @ L21.abs().t()
acc = {}
for S in (6, 7, 8, 9):
    C = emulated(X1, L21, S)
    acc[f"S{S}"] = {"max_err_over_absAabsB": float(((C - ref).abs() / den).max()), "max_err_over_max": float((C - ref).abs().max() / ref.abs().max())}
C = emulated(X1, L21, 8, full=True)
acc["S8_full"] = {"max_err_over_absAabsB": float(((C - ref).abs() / den).max())}
# the fp64 GEMM's own rounding noise, against a float128-free proxy: permuted summation order
ref2 = (X1[:, torch.randperm(h, device=dev)[:h]] * 0 + X1).flip(1) @ L21.flip(1).t()
acc["fp64_reorder_noise"] = float(((ref2 - ref).abs() / den).max())
out["accuracy_predict_update_m2048_k2048"] = acc
# random gaussian operands with wide dynamic range
A = torch.randn(1024, 4096, dtype=torch.float64, device=dev) * torch.exp(4 * torch.randn(1024, 4096, dtype=torch.float64, device=dev))
B = torch.randn(1024, 4096, dtype=torch.float64, device=dev) * torch.exp(4 * torch.randn(1024, 4096, dtype=torch.float64, device=dev))
ref = A @ B.t(); den = A.abs() @ B.abs().t()
out["accuracy_wide_dynamic_range"] = {f"S{S}": float(((emulated(A, B, S) - ref).abs() / den).max()) for S in (7, 8, 9, 10)}
print(json.dumps(out, indent=1))
json.dump(out, open("gpurun_out/ozaki_probe.json", "w"), indent=1)
