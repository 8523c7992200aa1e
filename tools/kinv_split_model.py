"""CPU model (numpy, no GPU): where does the INT8 digit-split path lose accuracy in K^-1 = W^T W, and which row-scale rule
would fix it?

The digit split keeps 55 bits below a row's MAXIMUM over the k range it is split for.  Rows of Y = W^T = L^-T carry their
largest entries at and next to the diagonal, so when the scale of an operand row is taken over a k range that contains
its diagonal, the small entries far from the diagonal keep fewer bits -- and those are the entries every off-diagonal
block of K^-1 is built from.  This script quantises Y row by row under different scale rules (no int8 arithmetic is
needed for this part), forms the first-order error of K^-1, adds the digit pairs s + t >= 7 that gemm_i8_kernel drops
(exact integer digit planes, as split_rows_kernel writes them) and reports the error both cause in the gradient traces
1/2 sum K^-1 o dK_p that the marginal-likelihood gradient is made of:

  full      one scale per row over the whole k extent (round 1)
  chunked   scales per k-chunk of N/4 (shipped: lauum_lower, potrf.cu)
  chunk/2   chunks of N/8
  columns   block columns of width nb: the B operand's scale starts AFTER its own diagonal block (what the distributed
            gradient does, dist.cu phase 2); the diagonal blocks themselves use chunked scales
  tiles     as chunked, but a B row's scale in the chunk that holds its diagonal skips its own 128-wide tile; diagonal tiles
            in FP64 (a possible single-GPU variant)

Result (profiles/kinv_split_model_N*_r2.json): the 55-bit quantisation contributes 1e-16 .. 5e-14 to the gradient under every
rule -- it is NOT what limits the INT8 gradient.  The digit pairs s + t >= 7 the kernel drops (2^-56 of the product of the
two rows' chunk maxima each; the anti-diagonals 7 and 8 are summed exactly here) are: 8.9e-11 with one scale per row,
2.8e-12 with chunks of N/4, 8.5e-13 with N/8, 5.5e-13 with the tiles rule at N = 4096.
"""
import json
import os
import sys

import numpy as np

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
TILE = 128
rng = np.random.default_rng(5)
x = rng.uniform(0, 1, (N, 2))
y = np.sin(3 * x).sum(axis=1) + rng.normal(0, 0.05, N)
amp2, l = np.exp(2 * 0.1), 0.3
d2 = ((x[:, None, :] - x[None, :, :]) ** 2)
Ks = amp2 * np.exp(-0.5 * d2.sum(-1) / l**2)
K = Ks + (0.05**2 + amp2 * 1e-12) * np.eye(N)
L = np.linalg.cholesky(K)
Y = np.linalg.solve(L, np.eye(N)).T.copy()            # Y = L^-T, upper triangular: row a = column a of inv(L)
Kinv = np.tril(Y @ Y.T)
dK = [2 * Ks] + [(d2[:, :, k] / l**2) * Ks for k in range(2)]      # SE gradient planes (ln a, ln l_k)
alpha = np.linalg.solve(K, y - y.mean())
grad = [0.5 * float(((np.outer(alpha, alpha) - Kinv) * np.tril(p, -1)).sum() * 2 + ((alpha**2 - np.diag(Kinv)) * np.diag(p)).sum())
        for p in dK]
gscale = max(abs(g) for g in grad)
tile_of = (np.arange(N) // TILE) * TILE


PAIRS = [(s_, t_) for s_ in range(7) for t_ in range(7) if 7 <= s_ + t_ <= 8]   # the leading dropped digit pairs (2^-70, 2^-78)


def split_chunk(rows, c0, c1, skip_tile):
    """The digit split of Y[rows, c0:c1] as gemm_i8.cu does it: per row e with max|x| 2^-e <= 127/128 over the chunk (with
    skip_tile: over the chunk minus the row's own 128-wide diagonal tile, whose entries are left out of this operand),
    m = round(x 2^(55-e)), 7 balanced base-256 digits.  Returns 2^e, the digit planes (float arrays of small integers) and
    the quantisation error q - x."""
    Yc = Y[rows, c0:c1].copy()
    if skip_tile:
        k = np.arange(c0, c1)[None, :]
        Yc = np.where(k >= (tile_of[rows] + TILE)[:, None], Yc, 0.0)
    mx = np.abs(Yc).max(axis=1)
    ok = mx > 0
    e = np.frexp(np.where(ok, mx, 1.0) * 128.0 / 127.0)[1]
    m = np.rint(Yc * np.ldexp(1.0, 55 - e)[:, None]).astype(np.int64)
    dq = np.where(ok[:, None], m.astype(np.float64) * np.ldexp(1.0, e - 55)[:, None] - Yc, 0.0)
    planes = [None] * 7
    for s_ in range(6, 0, -1):
        low = ((m & 0xFF) ^ 0x80) - 0x80
        planes[s_] = low.astype(np.float64)
        m = (m - low) >> 8
    planes[0] = m.astype(np.float64)
    return np.where(ok, np.ldexp(1.0, e), 0.0), planes, dq, Yc


def error_matrix(rows_a, rows_b, k_start, chunk, skip_tile_b=False):
    """error of sum_k Ya[i, k] Yb[j, k] on the INT8 path, chunk by chunk: the quantisation of both operands (first order:
    dA B^T + A dB^T) and the digit pairs the kernel drops (s + t >= 7; the two leading anti-diagonals are summed exactly)"""
    E = np.zeros((rows_a.size, rows_b.size))
    D = np.zeros_like(E)
    for c0 in range(k_start, N, chunk):
        c1 = min(N, c0 + chunk)
        sa, pa, da, Ya = split_chunk(rows_a, c0, c1, False)
        sb, pb, db, Yb = split_chunk(rows_b, c0, c1, skip_tile_b)
        E += da @ Yb.T + Ya @ db.T
        drop = np.zeros_like(E)
        for s_, t_ in PAIRS:
            drop += np.ldexp(pa[s_] @ pb[t_].T, -(14 + 8 * (s_ + t_)))
        D -= (sa[:, None] * sb[None, :]) * drop
    return E, D


def kinv_error(rule, chunk, nb=None):
    """(quantisation part, dropped-pairs part) of the error of the lower triangle of K^-1"""
    allr = np.arange(N)
    if rule == "columns":
        E, D = np.zeros((N, N)), np.zeros((N, N))
        for b0 in range(0, N, nb):
            b1 = min(N, b0 + nb)
            rb = np.arange(b0, b1)
            E[b0:b1, b0:b1], D[b0:b1, b0:b1] = error_matrix(rb, rb, b0, chunk)
            if b1 < N:      # rows below: k from b1 on, chunk boundaries relative to b1 -- B's scale never sees its diagonal block
                E[b1:, b0:b1], D[b1:, b0:b1] = error_matrix(np.arange(b1, N), rb, b1, chunk)
        return np.tril(E), np.tril(D)
    E, D = error_matrix(allr, allr, 0, chunk, skip_tile_b=(rule == "tiles"))
    E, D = np.tril(E), np.tril(D)
    if rule == "tiles":     # diagonal tiles in FP64
        for t0 in range(0, N, TILE):
            E[t0:t0 + TILE, t0:t0 + TILE] = 0.0
            D[t0:t0 + TILE, t0:t0 + TILE] = 0.0
    return E, D


def grad_error(E):
    return max(abs(0.5 * ((E * np.tril(p, -1)).sum() * 2 + (np.diag(E) * np.diag(p)).sum())) for p in dK) / gscale


res = {"N": N, "kernel": "SquaredExponential 2-D, a = e^0.1, l = 0.3, sigma_n = 0.05", "min_Lii": float(np.diag(L).min()),
       "row_dynamic_range_median": float(np.median(np.abs(Y).max(axis=1) / np.maximum(np.median(np.abs(Y) + np.tril(np.full((N, N), np.inf), -1), axis=1), 1e-300))),
       "what": "relative error of the gradient traces (max over parameters / max |grad|) from the INT8 product K^-1 = Y Y^T: row quantisation (first order) and the dropped digit pairs s + t = 7, 8 (exact)",
       "rules": {}}
for name, rule, chunk, nb in [("full", "plain", N, None), ("chunked N/4", "plain", N // 4, None), ("chunked N/8", "plain", N // 8, None),
                              ("chunked N/16", "plain", N // 16, None),
                              ("block columns nb = N/16, chunks N/4", "columns", N // 4, max(TILE, N // 16)),
                              ("block columns nb = N/32, chunks N/4", "columns", N // 4, max(TILE, N // 32)),
                              ("tiles: B scale skips its own tile, chunks N/4", "tiles", N // 4, None)]:
    E, D = kinv_error(rule, chunk, nb)
    res["rules"][name] = {"grad_rel_err_quantisation": float(grad_error(E)), "grad_rel_err_dropped_pairs": float(grad_error(D)),
                          "grad_rel_err_total": float(grad_error(E + D)), "kinv_max_abs_err": float(np.abs(E + D).max())}
    print(name, res["rules"][name], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/kinv_split_model_N{N}.json", "w"), indent=1)
