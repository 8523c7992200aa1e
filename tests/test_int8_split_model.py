"""CPU model of the FP64 -> 7 x int8 digit splitting used by csrc/gemm_i8.cu (split_rows_kernel, row_scale,
put_digits) and of the 28-product reconstruction (gemm_i8_kernel's epilogue), in exact integer arithmetic.
It pins down the error analysis the kernel relies on; the CUDA kernel itself is checked against numpy on the GPU
(tests/test_gpu_parity.py::test_int8_*)."""
import math
from fractions import Fraction

import numpy as np
import pytest

S = 7  # digit planes


def split_row(x):
    """digits (S, K) int64, exponent e: x ~ 2^e sum_s d_s 2^-(7+8s).  Mirrors row_scale + put_digits."""
    amax = float(np.abs(x).max())
    if not (1e-280 < amax < 1e280):
        return np.zeros((S, x.size), dtype=np.int64), None
    e = math.frexp(amax * (128.0 / 127.0))[1]           # ilogb(.) + 1: |x| 2^-e <= 127/128
    m = np.array([int(np.rint(np.ldexp(v, 55 - e))) for v in x], dtype=object)   # round(x 2^(55-e)), |m| < 2^55
    digits = np.zeros((S, x.size), dtype=np.int64)
    for s in range(S - 1, 0, -1):
        low = np.array([((int(v) & 0xFF) ^ 0x80) - 0x80 for v in m], dtype=object)   # signed low byte
        digits[s] = low.astype(np.int64)
        m = np.array([(int(v) - int(l)) >> 8 for v, l in zip(m, low)], dtype=object)
    digits[0] = m.astype(np.int64)
    return digits, e


def reconstruct(digits, e):
    return sum(Fraction(int(d)) * Fraction(2) ** (e - 7 - 8 * s) for s, d in enumerate(digits))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_digits_are_int8_and_reconstruct_to_2_pow_minus_55_of_the_row_maximum(seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(257) * np.exp(6 * rng.standard_normal(257))
    x[3] = 0.0
    x[5] = -np.abs(x).max() * (1 - 2**-40)          # next to the row maximum, negative
    d, e = split_row(x)
    assert d[0].min() >= -127 and d[0].max() <= 127            # top digit: |x| 2^-e <= 127/128 leaves room for the carry
    assert d[1:].min() >= -128 and d[1:].max() <= 127          # balanced base-256 digits fit int8
    amax = np.abs(x).max()
    for k in range(x.size):
        err = abs(reconstruct(d[:, k], e) - Fraction(float(x[k])))
        assert err <= Fraction(2) ** (e - 56)                    # half a unit of the 55-bit integer
        assert err <= Fraction(amax) * Fraction(2) ** -53       # = one FP64 ulp-ish of the row maximum
    assert np.all(d[:, 3] == 0)


def test_rows_that_cannot_be_split():
    z, e = split_row(np.zeros(16))
    assert e is None and not z.any()
    z, e = split_row(np.full(16, 1e-300))                       # underflowing tails count as zero
    assert e is None and not z.any()


def test_truncated_28_products_meet_the_fp64_normwise_bound():
    """sum over s + t <= 6 of 2^-(14 + 8 (s + t)) A_s B_t^T against the exact product of the FP64 inputs."""
    rng = np.random.default_rng(3)
    K = 192
    a = rng.standard_normal(K) * np.exp(2 * rng.standard_normal(K))
    b = rng.standard_normal(K) * np.exp(2 * rng.standard_normal(K))
    da, ea = split_row(a)
    db, eb = split_row(b)
    acc = Fraction(0)
    diag = [0] * S
    for s in range(S):
        for t in range(S - s):
            diag[s + t] += int(np.dot(da[s].astype(object), db[t].astype(object)))
    for g in range(S):
        assert abs(diag[g]) < 2**31                              # what the TMEM accumulators hold
        acc += Fraction(diag[g]) * Fraction(2) ** (ea + eb - 14 - 8 * g)
    exact = sum(Fraction(float(u)) * Fraction(float(v)) for u, v in zip(a, b))
    # worst case per product: two splitting errors of 2^-55 (of the row maxima) plus the six dropped s + t = 7 pairs,
    # 6 * 128^2 * 2^-70 of 2^(ea+eb) <= 4.07 amax bmax -- together below 2^-51 amax bmax
    scale = K * Fraction(float(np.abs(a).max())) * Fraction(float(np.abs(b).max()))
    assert abs(acc - exact) <= scale * Fraction(2) ** -51
    # digits of real data are not worst case: the observed error is well under one FP64 rounding of the row maxima.
    # (It is a NORMWISE statement: against |a|.|b| the error grows with the dynamic range inside a row.)
    assert abs(acc - exact) <= scale * Fraction(2) ** -55


def test_int32_accumulators_cannot_overflow_within_one_chunk():
    """7 products of worst-case digits over the longest k chunk the kernel issues (MAX_K = 16384)."""
    assert 7 * 16384 * 128 * 128 < 2**31
    assert 8 * 16384 * 128 * 128 >= 2**31                        # which is why longer k extents are chunked


def test_kinv_product_error_is_the_dropped_digit_pairs_of_the_diagonal_chunk(tmp_path):
    """tools/kinv_split_model.py (the digit split applied to K^-1 = W^T W of a dense SquaredExponential set, exact integer
    digit planes): the 55-bit quantisation is negligible, the dropped digit pairs s + t >= 7 set the gradient's INT8 error,
    and shorter k-chunks / scales that skip the diagonal tile shrink it -- the reasoning behind lauum_chunk (potrf.cu)."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "kinv_split_model.py"), "768"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1000:]
    res = json.load(open(tmp_path / "gpurun_out" / "kinv_split_model_N768.json"))["rules"]
    full, c4, c8 = res["full"], res["chunked N/4"], res["chunked N/8"]
    for rule in (full, c4, c8):
        assert rule["grad_rel_err_quantisation"] < 1e-14
        assert rule["grad_rel_err_dropped_pairs"] > 20 * rule["grad_rel_err_quantisation"]
    assert c4["grad_rel_err_total"] < full["grad_rel_err_total"]
    assert c8["grad_rel_err_total"] < 0.5 * c4["grad_rel_err_total"]
    assert res["tiles: B scale skips its own tile, chunks N/4"]["grad_rel_err_total"] < 0.5 * c4["grad_rel_err_total"]
