"""CPU restatement of the arithmetic of the tcgen05 chi-squared engine (csrc/chi2_ozaki.cuh): the balanced base-256 digit
recoding of `oz_digits4`, the per-row power-of-two scale of `k_oz_slice_rows`, the level sums the tensor cores accumulate
and the FP64 recombination of the epilogue.  Pure numpy: it pins the algorithm (exactness, digit ranges, int32 head-room,
accuracy per plane count) without a GPU; the CUDA kernels are held to the same numbers by tests/test_gpu_tcgen05.py."""
import numpy as np
import pytest


def frac_bits(S):
    return 6 + 8 * (S - 1)


def slice_rows(X, S):
    """fp64 rows -> (digits[S, rows, n] int64 in [-128, 127], scale[rows] = 2^e with |x| < 2^e)."""
    X = np.asarray(X, dtype=np.float64)
    mx = np.abs(X).max(axis=1)
    e = np.where(mx > 0, np.frexp(mx)[1], 0)                      # mx = m 2^e, m in [0.5, 1)  ==  ilogb(mx) + 1
    v = np.rint(np.ldexp(X, (frac_bits(S) - e)[:, None])).astype(np.int64)
    bias = sum(0x80 << (8 * k) for k in range(S - 1))             # 0x80 at every digit position below the top one
    u = (v + bias) ^ bias                                         # one addition + one xor (oz_digits4)
    digits = np.empty((S,) + X.shape, dtype=np.int64)
    for s in range(S):
        b = S - 1 - s                                             # plane s = byte S-1-s of u
        if s == 0:
            digits[s] = u >> (8 * b)                              # what remains above the biased digits (signed)
        else:
            digits[s] = ((u >> (8 * b)) & 0xFF).astype(np.int8)   # signed byte
    return digits, np.ldexp(1.0, e), v


@pytest.mark.parametrize("S", [5, 6, 7])
def test_digit_recoding_is_exact(S):
    rng = np.random.default_rng(S)
    X = rng.standard_normal((64, 333)) * np.exp(rng.uniform(-20, 20, (64, 1)))
    X[3, 5] = 0.0
    X[7] = 0.0                                                    # an all-zero row keeps scale 1 and digits 0
    d, scale, v = slice_rows(X, S)
    assert d.min() >= -128 and d.max() <= 127 and np.abs(d[0]).max() <= 65
    recon = sum(d[s] << (8 * (S - 1 - s)) for s in range(S))      # sum_s d_s 256^(S-1-s) == the fixed-point integer
    assert np.array_equal(recon, v)
    # the fixed-point number is x rounded to FRAC_BITS bits below the row's power-of-two scale
    err = np.abs(np.ldexp(v.astype(np.float64), -frac_bits(S)) * scale[:, None] - X)
    assert np.all(err <= np.ldexp(scale, -frac_bits(S) - 1)[:, None] * (1 + 1e-12))
    assert np.all(np.abs(X) < scale[:, None]) and scale[7] == 1.0 and not d[:, 7].any()
    # the same digits come out of the sequential definition d = ((v + 128) & 255) - 128, v <- (v + 128) >> 8
    w = v.copy()
    for s in range(S - 1, 0, -1):
        ds = ((w + 128) & 255) - 128
        assert np.array_equal(ds, d[s])
        w = (w + 128) >> 8
    assert np.array_equal(w, d[0])


@pytest.mark.parametrize("S,rel_tol", [(5, 3e-9), (6, 2e-11), (7, 1e-13)])
def test_level_sums_recombine_to_the_fp64_contraction(S, rel_tol):
    """chi2 = |W r|^2 from the S (S + 1) / 2 kept digit-plane products, accumulated per level l = i + j in int32 and
    recombined as y = 2^(eR + eW - 12) sum_l 2^-8l acc_l, against long-double arithmetic."""
    rng = np.random.default_rng(10 + S)
    n, B = 257, 24
    A = rng.standard_normal((n, n)) * 0.05
    L = np.linalg.cholesky(A @ A.T + np.diag(rng.uniform(0.01, 0.05, n)))
    W = np.tril(np.linalg.inv(L))
    R = rng.standard_normal((B, n)) * rng.uniform(0.05, 3.0, (B, 1))
    dR, sR, _ = slice_rows(R, S)
    dW, sW, _ = slice_rows(W, S)
    y = np.zeros((B, n), dtype=np.longdouble)
    for lvl in range(S):
        acc = np.zeros((B, n), dtype=np.int64)
        for i in range(lvl + 1):
            acc += dR[i] @ dW[lvl - i].T
        assert np.abs(acc).max() < 2**31                          # fits the int32 accumulators in TMEM
        y += acc.astype(np.longdouble) * np.longdouble(2.0) ** (-8 * lvl)
    y *= (sR[:, None] * sW[None, :]).astype(np.longdouble) * np.longdouble(2.0) ** -12
    chi2 = (y * y).sum(axis=1)
    ref = ((W.astype(np.longdouble) @ R.T.astype(np.longdouble)) ** 2).sum(axis=0)
    assert np.max(np.abs(chi2 - ref) / ref) < rel_tol


def test_int32_headroom_bound():
    """Worst case of one level: S products of |d_i d_j| <= 2^14 over n terms; the library keeps the tcgen05 engine for
    n_sn <= 16384 (cosmolike.cu) where 7 * 2^14 * n < 2^31."""
    assert 7 * 2**14 * 16384 < 2**31 and 7 * 2**14 * 18725 > 2**31


def _fma(a, b, c):
    """Correctly rounded a * b + c (what DFMA computes), through exact rational arithmetic."""
    from fractions import Fraction
    return float(Fraction(a) * Fraction(b) + Fraction(c))


@pytest.mark.parametrize("S", [5, 6, 7])
def test_epilogue_fold_one_fma_per_level_group(S):
    """The epilogue's fold (csrc/chi2_ozaki.cuh): a group of up to three levels is built as the integer
    g = a_l 2^16 + a_(l+1) 2^8 + a_(l+2) in the low mantissa bits of the double kBias + g; group 0 takes its bias off inside
    its FMA (exact), the later groups are added bias and all, and the exactly representable sum of the carried constants
    comes off together with the column scale: y = fma(h, cs, -C cs).  Against the exact value of sum_l 2^-8l a_l: relative
    2^-52 when the high groups dominate, and never worse than an absolute 2^-40 (units of the level-0 digit product) when
    they cancel - the a-priori bound already allows 2^-39 per TERM of the sum that produced a_l."""
    from fractions import Fraction
    rng = np.random.default_rng(100 + S)
    kBias = 6755401588539392.0                                    # 1.5 * 2^52 + 2^31
    assert kBias == 1.5 * 2.0**52 + 2.0**31
    groups = {7: [(0, 3), (3, 3), (6, 1)], 6: [(0, 3), (3, 3)], 5: [(0, 3), (3, 2)]}[S]
    C = kBias * sum(2.0 ** (-8 * (l0 + cnt - 1)) for l0, cnt in groups[1:])
    assert Fraction(C) == Fraction(kBias) * sum(Fraction(1, 2 ** (8 * (l0 + cnt - 1))) for l0, cnt in groups[1:])   # representable
    n_sn = 2048
    n_abs = 0
    for trial in range(400):
        # level l holds l + 1 products of |d_i d_j| <= 2^14 over n_sn terms; every tenth trial has cancelling high levels
        acc = [int(rng.integers(-(l + 1) * 2**14 * n_sn, (l + 1) * 2**14 * n_sn + 1)) for l in range(S)]
        if trial % 10 == 0:
            acc[0] = acc[1] = acc[2] = 0
            acc[3] = int(rng.integers(-3, 4))
        cs = 2.0 ** int(rng.integers(-30, 10))
        h = None
        for gi, (l0, cnt) in enumerate(groups):
            g = sum(acc[l0 + q] << (8 * (cnt - 1 - q)) for q in range(cnt))
            assert abs(g) < 2**51
            t = kBias + g                                          # exact: the integer sits in the mantissa
            assert Fraction(t) == Fraction(kBias) + g
            wg = 2.0 ** (-8 * (l0 + cnt - 1))
            h = _fma(t, wg, -kBias * wg) if gi == 0 else _fma(t, wg, h)
        y = _fma(h, cs, -C * cs)
        exact = sum(Fraction(acc[l], 2 ** (8 * l)) for l in range(S)) * Fraction(cs)
        err = abs(Fraction(y) - exact)
        rel = float(err / abs(exact)) if exact != 0 else 0.0
        assert rel <= 2.0**-51 or float(err / Fraction(cs)) <= 2.0**-40, (trial, rel, float(err))
        n_abs += rel > 2.0**-51
    assert n_abs <= 40                                             # only the cancelling cases lean on the absolute bound
