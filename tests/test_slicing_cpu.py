"""The arithmetic claims behind the exact int8 slice products of the correlation row, checked on the CPU
with a NumPy mirror of k_corr_intcheck / k_corr_slices8 / k_corr_f32check / k_corr_fslices8 / k_corr_faccum(_sxy) and exact rational
arithmetic: the digits reconstruct every entry exactly, every digit product fits int32, and the recombined sums equal the exact sums
up to the final float64 roundings."""
from fractions import Fraction

import numpy as np
import pytest


def slices_float32_rows(X):
    """(K, hi[m], S[K][m][m] int64 digits in [-127, 127]) for a float64 array holding float32 numbers -- one exponent per row."""
    m = X.shape[0]
    hi = np.zeros(m, dtype=np.int64)
    need = 0
    for i in range(m):
        nz = X[i][X[i] != 0]
        if nz.size == 0:
            continue
        assert np.array_equal(nz.astype(np.float32).astype(np.float64), nz)          # the route's precondition (k_corr_f32check)
        _, e = np.frexp(nz)
        hi[i] = e.max()
        need = max(need, int(e.max() - (e - 24).min()))
    assert need <= 63
    K = max(1, -(-need // 7))
    S = np.zeros((K,) + X.shape, dtype=np.int64)
    for i in range(m):
        q = np.ldexp(np.abs(X[i]), int(7 * K - hi[i]))
        assert np.array_equal(q, np.floor(q)) and q.max(initial=0) < 2.0 ** (7 * K)       # an integer below 2^(7K): nothing truncated
        qi = q.astype(np.uint64)
        for t in range(K):
            S[t, i] = ((qi >> np.uint64(7 * (K - 1 - t))) & np.uint64(127)).astype(np.int64) * np.sign(X[i]).astype(np.int64)
    return K, hi, S


def frac_sum(values):
    return sum((Fraction(float(v)) for v in values), Fraction(0))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_float32_rows_slice_exactly_and_products_are_exact(seed):
    rng = np.random.default_rng(seed)
    m = 24
    X = (rng.lognormal(0.0, 3.0, size=(m, m)) * rng.choice([-1.0, 1.0], size=(m, m))).astype(np.float32).astype(np.float64)
    X[rng.random((m, m)) < 0.4] = 0.0
    X[5] = 0.0                                                # an empty row
    I = (X != 0).astype(np.int64)
    K, hi, S = slices_float32_rows(X)
    assert np.abs(S).max() <= 127
    # the digits reconstruct every entry exactly
    R = sum(np.ldexp(S[t].astype(np.float64), (hi - 7 * (t + 1))[:, None]) for t in range(K))
    assert np.array_equal(R, X)
    # Sx = X0 I^T from K integer products; Sxy = X0 X0^T from the pairs a <= b with T + T^T
    Sx = np.zeros((m, m))
    for t in range(K):
        T = S[t] @ I.T                                        # int64 here; on the device int32: |T| <= 127 m
        assert np.abs(T).max() <= 127 * m
        Sx = Sx + np.ldexp(T.astype(np.float64), (hi - 7 * (t + 1))[:, None])
    Sxy = np.zeros((m, m))
    for a in range(K):
        for b in range(a, K):
            T = S[a] @ S[b].T
            assert np.abs(T).max() <= 127 * 127 * m
            V = (T + T.T if a != b else T).astype(np.float64)
            Sxy = Sxy + np.ldexp(V, hi[:, None] + hi[None, :] - 7 * (a + b + 2))
    for i in range(0, m, 5):
        for j in range(0, m, 3):
            ex = frac_sum(X[i][I[j] != 0])
            assert abs(Fraction(float(Sx[i, j])) - ex) <= abs(ex) * Fraction(1, 10 ** 15)
            exy = sum((Fraction(float(x)) * Fraction(float(y)) for x, y in zip(X[i], X[j])), Fraction(0))
            tol = sum((abs(Fraction(float(x)) * Fraction(float(y))) for x, y in zip(X[i], X[j])), Fraction(0)) * Fraction(K * K, 10 ** 16)
            assert abs(Fraction(float(Sxy[i, j])) - exy) <= tol


def test_integer_maps_slice_exactly():
    rng = np.random.default_rng(7)
    m = 30
    X = rng.integers(0, 5000, size=(m, m)).astype(np.float64)
    X[rng.random((m, m)) < 0.5] = 0.0
    I = (X != 0).astype(np.int64)
    ns = max(1, -(-int(X.max()).bit_length() // 7))
    S = [((X.astype(np.int64) >> (7 * t)) & 127) for t in range(ns)]
    assert np.array_equal(sum(S[t] << (7 * t) for t in range(ns)), X.astype(np.int64))
    Sx = sum((S[t] @ I.T) * float(1 << (7 * t)) for t in range(ns))
    Sxy = np.zeros((m, m))
    for a in range(ns):
        for b in range(a, ns):
            T = S[a] @ S[b].T
            Sxy += (T + T.T if a != b else T) * float(1 << (7 * (a + b)))
    Xi = X.astype(np.int64)
    assert np.array_equal(Sx, (Xi @ I.T).astype(np.float64)) and np.array_equal(Sxy, (Xi @ Xi.T).astype(np.float64))
