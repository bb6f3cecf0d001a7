"""CPU: executable statements of two device algorithms whose exactness the GPU parity tests rely on.  Neither function is product
code -- each restates, in NumPy / exact rational arithmetic, what the kernel named in its docstring does, so that the argument for
bit-exactness can be checked without a GPU (the `-m gpu` tests then check the kernels themselves against the oracle):

* the per-diagonal median select of k_diag_hist<PASS> / diag_select_body (csrc/gcx_kernels.cuh): digits of 9 + 9 + 7 + 7 bits of the
  order-preserving key, early exit after two digits when the candidates' low 14 bits are all zero, the upper middle of an even count
  tracked as min(next occupied bin at the last level, smallest key of the next bucket at the levels above);
* the o/e quotient of diag_apply_quad: x * r corrected by one fma against the tabulated reciprocal r = RN(1 / av), with the exact
  division taken whenever the result could round to a different float32.
"""
from fractions import Fraction

import numpy as np
import pytest

WIDTH = (9, 9, 7, 7)
SHIFT = (23, 14, 7, 0)
NONE = 0xFFFFFFFF


def f2ord(v):
    b = np.asarray(v, np.float32).view(np.uint32).astype(np.uint64)
    return np.where(b >> np.uint64(31), (~b) & np.uint64(0xFFFFFFFF), b | np.uint64(0x80000000))


def ord2f(k):
    k = int(k)
    b = (k & 0x7FFFFFFF) if k >> 31 else (~k) & 0xFFFFFFFF
    return float(np.array([b], np.uint32).view(np.float32)[0])


def select_median(values):
    """(median of the non-zero entries as the kernels compute it, number of digit passes used)."""
    keys = f2ord(values[values != 0])
    if keys.size == 0:
        return 0.0, 1
    pre, nbp, nxt, cnt, kr = 0, NONE, NONE, 0, 0
    for p in range(4):
        sh, w = SHIFT[p], WIDTH[p]
        if p == 0:
            matched = keys
        else:
            hi = keys >> np.uint64(sh + w)
            matched = keys[hi == np.uint64(pre)]
            tracked = keys[hi == np.uint64(nbp)]
            if tracked.size:
                nxt = min(nxt, int(tracked.min()))
        hist = np.bincount(((matched >> np.uint64(sh)) & np.uint64((1 << w) - 1)).astype(np.int64), minlength=512)
        if p == 0:
            cnt = int(hist.sum())
            kr = (cnt - 1) // 2
        cum = np.cumsum(hist)
        b = int(np.searchsorted(cum, kr, side="right"))
        run, hb = int(cum[b] - hist[b]), int(hist[b])
        kin = kr - run
        above = np.nonzero(hist[b + 1:])[0]
        nb = b + 1 + int(above[0]) if above.size else -1
        sel = (pre << w) | b
        nbsel = ((pre << w) | nb) if nb >= 0 else NONE
        clean = p == 1 and not np.any(matched & np.uint64(0x3FFF))
        if p == 3 or clean:
            key = sel << sh
            if nb >= 0:
                nxt = min(nxt, nbsel << sh)
            lo = ord2f(key)
            need = cnt % 2 == 0 and kin + 1 >= hb
            hi_v = ord2f(nxt) if need else lo
            return ((lo + hi_v) / 2.0 if cnt % 2 == 0 else lo), p + 1
        pre, kr, nbp = sel, kin, nbsel
    raise AssertionError("unreachable")


POOLS = [np.array([1, 2, 3, 4, 7, 8], np.float32), np.array([1, 100, 1000, 1023, 1024, 1025], np.float32),
         np.array([5000, 65537, 3, 16777215], np.float32), np.array([0.5, 1e-3, 2.0, 1536.0, 1536.25], np.float32),
         np.array([-3, -1, 2, 1024, -1025.5], np.float32)]


@pytest.mark.parametrize("seed", range(4))
def test_radix_select_specification_matches_numpy_median(seed):
    rng = np.random.default_rng(seed)
    passes = []
    for t in range(3000):
        pool = POOLS[rng.integers(len(POOLS))] if t % 2 else np.concatenate(POOLS)
        v = rng.choice(pool, size=int(rng.integers(1, 14))).astype(np.float32)
        if t % 5 == 0:
            v = rng.standard_normal(v.size).astype(np.float32)
        if t % 11 == 0:
            v[rng.random(v.size) < 0.5] = 0.0
        got, used = select_median(v)
        nz = v[v != 0]
        exp = float(np.median(nz.astype(np.float64))) if nz.size else 0.0
        assert got == exp, (v, got, exp)
        passes.append(used)
    assert min(passes) <= 2 and max(passes) == 4          # both the early exit and the full descent were exercised


def test_radix_select_counts_need_two_passes():
    """Integer counts up to 1023 are fixed by sign + exponent + nine mantissa bits: the select ends after the second digit."""
    rng = np.random.default_rng(9)
    for _ in range(500):
        v = rng.integers(0, 1024, size=int(rng.integers(1, 40))).astype(np.float32)
        got, used = select_median(v)
        nz = v[v != 0]
        assert used <= 2 and got == (float(np.median(nz.astype(np.float64))) if nz.size else 0.0)


def _rn(fr):
    return float(fr)            # Fraction -> float is correctly rounded


def quotient_as_kernel(x, av):
    """diag_apply_quad, divide mode, for one live entry: returns (float32 result, took_exact_division)."""
    r = 1.0 / av                                             # k_diag_recip
    q0 = x * r
    e = _rn(Fraction(x) - Fraction(q0) * Fraction(av))       # fma(-q0, av, x)
    q = _rn(Fraction(q0) + Fraction(e) * Fraction(r))        # fma(e, r, q0)
    bits = int(np.float64(q).view(np.uint64))
    lo, ex = bits & 0xFFFFFFFF, (bits >> 52) & 0x7FF
    risky = (((lo & 0x1FFFFFFF) - 0x0FFFFFFF) & 0xFFFFFFFF) <= 2 or ((ex - 899) & 0xFFFFFFFF) >= (0x7FF - 899)
    if risky:
        q = x / av
    with np.errstate(over="ignore"):
        return np.float32(q), risky


def test_reciprocal_quotient_is_the_reference_quotient():
    """f32(RN64(x / av)) (lib/normalizeCore.pyx:105-117) == the kernel's value on adversarial pairs -- quotients within an ulp of the
    midpoint between two floats, where a last-bit error of the float64 quotient would change the float32 -- and on ordinary ones.
    Also checks the bound the guard relies on: the corrected product is never more than one float64 ulp from the true quotient."""
    rng = np.random.default_rng(77)
    m = 1500
    x = np.ldexp(rng.integers(1 << 23, 1 << 24, size=m).astype(np.float64), rng.integers(-30, 8, size=m)).astype(np.float32)
    f = np.ldexp(rng.integers(1 << 23, 1 << 24, size=m).astype(np.float64), rng.integers(-24, 0, size=m)).astype(np.float32)
    mid = f.astype(np.float64) + 0.5 * np.spacing(f).astype(np.float64)
    av = x.astype(np.float64) / mid
    xs = list(x.astype(np.float64)) + list(rng.integers(1, 5000, size=4000).astype(np.float64)) + [1.0, 3.0, 1e-30, 2.5e38]
    avs = list(av) + list(rng.uniform(0.01, 900.0, size=4000)) + [1e-300, 1e300, 7.0, 1e-3]
    fallbacks = 0
    for xi, ai in zip(xs, avs):
        got, risky = quotient_as_kernel(float(xi), float(ai))
        with np.errstate(over="ignore"):
            exp = np.float32(np.float64(xi) / np.float64(ai))
        assert got.view(np.uint32) == exp.view(np.uint32) or (np.isnan(got) and np.isnan(exp)), (xi, ai, got, exp)
        fallbacks += risky
        # the bound: |corrected product - RN64(x / av)| <= 1 ulp wherever the fast path is taken
        if not risky:
            r = 1.0 / ai
            q0 = xi * r
            q = _rn(Fraction(q0) + Fraction(_rn(Fraction(xi) - Fraction(q0) * Fraction(ai))) * Fraction(r))
            assert abs(int(np.float64(q).view(np.int64)) - int(np.float64(xi / ai).view(np.int64))) <= 1
    assert fallbacks >= m                                     # every adversarial pair went through the exact division
