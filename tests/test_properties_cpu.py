"""CPU property tests (hypothesis) for the host-side logic that decides masks and partitions."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from gcmapexplorer_b200 import ccmapHelpers as cmh
from gcmapexplorer_b200 import dist as gd
from gcmapexplorer_b200 import util
from oracle import oracle_np as onp


@settings(max_examples=60, deadline=None)
@given(n=st.integers(4, 120), seed=st.integers(0, 10_000), pct=st.one_of(st.none(), st.integers(1, 100)),
       occ=st.one_of(st.none(), st.floats(0.01, 0.99)))
def test_mask_rules_match_oracle_on_random_statistics(n, seed, pct, occ):
    """mask_from_row_stats == oracle.mask_from_stats for arbitrary (consistent) row statistics, incl. the error cases."""
    rng = np.random.default_rng(seed)
    zc = rng.integers(0, n + 1, size=n)
    empty = zc == n
    zc = np.where(rng.random(n) < 0.1, n, zc)          # a few empty rows
    empty = zc == n
    rowsum = np.where(empty, 0.0, rng.random(n) + 1.0)

    def run(f, *a):
        try:
            return f(*a), None
        except (AssertionError, ValueError) as e:
            return None, type(e)
    exp, eerr = run(onp.mask_from_stats, empty, np.where(empty, 0, zc), pct, occ)
    got, gerr = run(cmh.mask_from_row_stats, rowsum, zc, n, pct, occ)
    assert eerr == gerr
    if exp is not None:
        assert np.array_equal(exp, got)


@settings(max_examples=200, deadline=None)
@given(n=st.integers(1, 300_000), world=st.integers(1, 16))
def test_row_partitions_tile_the_rows(n, world):
    for part in (gd.row_partition, gd.row_partition_aligned):
        pos = 0
        for r in range(world):
            row0, nloc = part(n, world, r)
            assert nloc >= 0 and (nloc == 0 or row0 == pos)
            pos += nloc
        assert pos == n
    if n >= 2048 * world:
        sizes = [gd.row_partition_aligned(n, world, r)[1] for r in range(world)]
        assert max(sizes) - min(sizes) <= 2048                      # cuts are rounded to 1024 rows
        assert all(gd.row_partition_aligned(n, world, r)[0] % 1024 == 0 for r in range(world))


@settings(max_examples=200, deadline=None)
@given(v=st.floats(1e-12, 1e6))
def test_locate_significant_digit(v):
    k = util.locate_significant_digit_after_decimal(v)
    assert k == onp.locate_significant_digit_after_decimal(v)
    assert int(abs(v) * 10 ** k) >= 1 and (k == 0 or int(abs(v) * 10 ** (k - 1)) == 0)


@settings(max_examples=30, deadline=None)
@given(n=st.integers(2, 60), level=st.integers(1, 9), seed=st.integers(0, 1000))
def test_downsample_oracle_shape_and_mass(n, level, seed):
    """Oracle restatement of downSample2DMap: shape rule and conservation of the summed block mass (row/col 0 are dropped)."""
    rng = np.random.default_rng(seed)
    A = rng.integers(0, 20, size=(n, n)).astype(np.float32)
    out = onp.downsample_2d(A, level, "sum")
    assert out.shape == onp.downsample_shape(A.shape, level)
    assert not out[0].any() and not out[:, 0].any()
    assert out.sum() == A[1:, 1:].sum()


# ---- host-side helpers of the widened rows (lib/util.py, lib/ccmap.py:845-870) -- no GPU needed ----
def test_resolution_helpers_known_answers():
    """The examples of the reference's docstrings (lib/util.py:176-286)."""
    from gcmapexplorer_b200 import util
    for binsize, text in [(1, "1b"), (10, "10b"), (10000, "10kb"), (100000, "100kb"), (125500, "125.5kb"), (1000000, "1mb"), (1634300, "1.6343mb")]:
        assert util.binsizeToResolution(binsize) == text
    for text, binsize in [("1b", 1), ("10b", 10), ("10kb", 10000), ("16kb", 16000), ("1.23kb", 1230), ("1.6mb", 1600000), ("1.457mb", 1457000)]:
        assert util.resolutionToBinsize(text) == binsize
    with pytest.raises(ValueError):
        util.resolutionToBinsize("10 parsec")


def test_outlier_helpers_and_downsample1d():
    from gcmapexplorer_b200 import ccmap as cmp, util
    x = np.array([1.0, 1.1, 0.9, 1.05, 0.95, 1.0, 25.0, 1.02])
    assert list(np.nonzero(util.detectOutliers1D(x))[0]) == [6]
    mx = np.ma.masked_equal(np.array([0.0, 1.0, 1.1, 0.9, 0.0, 30.0, 1.05, 0.95]), 0.0)
    out = util.detectOutliersMasked1D(mx, thresh=3.5)
    assert np.array_equal(np.ma.getmaskarray(out), np.array([1, 0, 0, 0, 1, 1, 0, 0], dtype=bool))      # input mask kept, outlier added
    assert np.array_equal(out.filled(0.0), np.array([0, 1.0, 1.1, 0.9, 0, 0, 1.05, 0.95]))
    t = np.array([9.0, 1.0, 0.0, 3.0, 5.0, 0.0, 0.0, 7.0])                 # element 0 stays apart; 7 values -> 4 groups of 2, zero-padded
    assert np.array_equal(cmp.downSample1D(t, level=2, func="sum"), [0.0, 1.0, 8.0, 0.0, 7.0])
    assert np.array_equal(cmp.downSample1D(t, level=2, func="max"), [0.0, 1.0, 5.0, 0.0, 7.0])
    m = cmp.downSample1D(t, level=2, func="mean")                         # mean over the non-zero entries of a group
    assert m[1] == 1.0 and m[2] == 4.0 and m[4] == 7.0


@settings(max_examples=25, deadline=None)
@given(st.integers(min_value=1, max_value=10 ** 9))
def test_resolution_roundtrip(binsize):
    from gcmapexplorer_b200 import util
    text = util.binsizeToResolution(binsize)
    assert abs(util.resolutionToBinsize(text) - binsize) <= max(1, binsize // 10 ** 9)
