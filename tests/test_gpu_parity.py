"""GPU parity tests (pytest -m gpu): every call goes through the C ABI of libgcxnorm.so and is compared with
the oracle (oracle/oracle_np.py) and the golden vectors produced by the real reference.

Tolerances (north_star: bias 1e-6, fp32 matrices 1e-4 -- asserted much tighter here):
  BIAS_RTOL   bias / KR x vectors vs fp64 oracle
  MAT_RTOL    fp32 matrices vs fp64 oracle rounded to fp32 (a couple of fp32 ulps)
  ASIS_RTOL   fp32 matrices vs the as-is fp32 reference IC (its own noise floor)
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden, rel_err
from oracle import oracle_np as onp
from oracle import ref_loader as rl

pytestmark = pytest.mark.gpu

BIAS_RTOL = 1e-9
MAT_RTOL = 3e-7
ASIS_RTOL = 5e-5


@pytest.fixture(scope="module")
def gx():
    from gcmapexplorer_b200 import device
    assert device._lib.device_count() >= 1
    return device


def _ic_device(gx, A, tol, iteration, keep):
    with gx.DeviceMap.from_host(A) as dm:
        dm.set_mask(keep)
        r = dm.ic(tol, iteration)
        mn, mx = dm.apply_sym(None, masked_mode=1, in_place=False)
        r["matrix"] = dm.download(1)
        r["min"], r["max"] = mn, mx
    return r


@pytest.mark.parametrize("n", [5, 33, 97, 256, 1000, 1025])
def test_upload_download_roundtrip_and_checksum(gx, n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)).astype(np.float32)
    with gx.DeviceMap.from_host(A) as dm:
        assert np.array_equal(dm.download(0), A)
        c0 = dm.checksum(0)
        # strided (non-contiguous rows) host view
        big = np.zeros((n, n + 7), dtype=np.float32)
        big[:, :n] = A
        dm.upload(big[:, :n])
        assert np.array_equal(dm.download(0), A)
        assert dm.checksum(0) == c0


@pytest.mark.parametrize("n,seed", [(64, 1), (97, 2), (482, 3), (1500, 4)])
def test_row_stats_and_mask(gx, n, seed):
    A = onp.synth_map(n, seed=seed, scale=30.0, alpha=1.3, empty_frac=0.03)
    from gcmapexplorer_b200 import ccmapHelpers as cmh
    with gx.DeviceMap.from_host(A) as dm:
        rs, zc = dm.row_stats()
    np.testing.assert_array_equal(rs, A.astype(np.float64).sum(axis=1))      # integer counts: exact
    np.testing.assert_array_equal(zc, (A == 0).sum(axis=1))
    for kw in (dict(), dict(threshold_percentile=90), dict(threshold_percentile=50), dict(threshold_data_occup=0.3),
               dict(threshold_data_occup=0.05)):
        try:
            exp = onp.get_nonzeros_index(A, kw.get("threshold_percentile"), kw.get("threshold_data_occup"))
        except ValueError:
            with pytest.raises(ValueError):            # same error when the filter leaves nothing
                cmh.get_nonzeros_index(A, **kw)
            continue
        got = cmh.get_nonzeros_index(A, **kw)
        assert got.dtype == bool and np.array_equal(got, exp)


@pytest.mark.parametrize("name", golden_names("MASK"))
def test_mask_golden(gx, name):
    from gcmapexplorer_b200 import ccmapHelpers as cmh
    g = load_golden(name)
    keep = cmh.get_nonzeros_index(g["input"], threshold_percentile=g["kwargs"].get("percentile_threshold_no_data"),
                                  threshold_data_occup=g["kwargs"].get("threshold_data_occup"))
    assert np.array_equal(keep, g["keep"])


def test_mask_errors(gx):
    from gcmapexplorer_b200 import ccmapHelpers as cmh
    A = onp.synth_map(40, seed=3)
    with pytest.raises(AssertionError):
        cmh.get_nonzeros_index(A, 90, 0.5)
    with pytest.raises(ValueError):
        cmh.get_nonzeros_index(A, None, 1.5)


@pytest.mark.parametrize("n,seed", [(5, 0), (31, 1), (300, 2), (1031, 3), (2048, 4)])
def test_ic_vs_fp64_oracle(gx, n, seed):
    A = onp.synth_map(n, seed=seed) if n > 5 else load_golden("kat5_ic")["input"]
    M64, B, passes, l1 = onp.ic_fp64(A, 1e-4, 500)
    r = _ic_device(gx, A, 1e-4, 500, np.ones(n, dtype=bool))
    assert r["passes"] == passes and r["matvecs"] == passes + 1
    assert r["products_launched"] == passes + (2 if r["symmetric"] else 1)
    np.testing.assert_allclose(r["bias"], B, rtol=BIAS_RTOL)
    assert rel_err(r["matrix"], M64.astype(np.float32)) < MAT_RTOL
    assert abs(r["l1"] - l1[-1]) <= 1e-6 * max(l1[-1], 1e-30) + 1e-15
    nzv = M64.astype(np.float32)
    assert r["max"] == pytest.approx(float(nzv.max()), rel=MAT_RTOL)
    assert r["min"] == pytest.approx(float(nzv[nzv != 0].min()), rel=MAT_RTOL)


@pytest.mark.parametrize("name", golden_names("IC"))
def test_ic_golden_dropin(gx, name):
    """normalizeCCMapByIC through CCMAP objects vs (a) the real reference's as-is output (loose, its own
    fp32 noise) and (b) the fp64 oracle (tight), incl. mask, quirks, min/max and pass count."""
    from gcmapexplorer_b200 import ccmap as cmp, normalizer as nz
    g = load_golden(name)
    c = cmp.CCMAP()
    c.gen_from_matrix(g["input"], 100000, xlabel="chrS", ylabel="chrS")
    c.make_unreadable()
    out = nz.normalizeCCMapByIC(c, **g["kwargs"])
    assert out is not None and out.matrix is not None and c.matrix is None      # states: lib/normalizer.py:583
    got = np.array(out.matrix)
    assert np.array_equal(out.bNoData, g["bNoData"])                             # bit-exact mask
    assert rel_err(got, g["matrix"]) < ASIS_RTOL
    ref = onp.normalize_ic(g["input"], fp64=True, **g["kwargs"])
    assert rel_err(got, ref["matrix"]) < MAT_RTOL
    assert nz.last_run_info["passes"] == ref["passes"]
    np.testing.assert_allclose(nz.last_run_info["bias"][ref["keep"]], ref["bias"], rtol=BIAS_RTOL)
    assert out.minvalue == pytest.approx(ref["minvalue"], rel=MAT_RTOL)
    assert out.maxvalue == pytest.approx(ref["maxvalue"], rel=MAT_RTOL)
    assert out.minvalue == pytest.approx(float(g["minvalue"]), rel=ASIS_RTOL)


@pytest.mark.parametrize("name", golden_names("KR"))
def test_kr_golden_dropin(gx, name):
    from gcmapexplorer_b200 import ccmap as cmp, normalizer as nz
    g = load_golden(name)
    c = cmp.CCMAP()
    c.gen_from_matrix(g["input"], 100000, xlabel="chrS", ylabel="chrS")
    c.make_unreadable()
    out = nz.normalizeCCMapByKR(c, **g["kwargs"])
    assert out.matrix is None and c.matrix is None                               # lib/normalizer.py:359-360
    out.make_readable()
    got = np.array(out.matrix)
    assert np.array_equal(out.bNoData, g["bNoData"])
    assert (nz.last_run_info["outer"], nz.last_run_info["MVP"]) == (int(g["outer"]), int(g["MVP"]))
    assert rel_err(got, g["matrix"]) < MAT_RTOL                                  # vs the REAL reference output
    assert out.minvalue == pytest.approx(float(g["minvalue"]), rel=MAT_RTOL)
    assert out.maxvalue == pytest.approx(float(g["maxvalue"]), rel=MAT_RTOL)
    kept = ~g["bNoData"]
    assert np.all(got[~kept] == 0) and np.all(got[:, ~kept] == 0)


@pytest.mark.parametrize("name", golden_names("KR"))
def test_kr_x_vs_reference(gx, name):
    g = load_golden(name)
    T = onp.clamp(g["input"], g["kwargs"].get("vmin"), g["kwargs"].get("vmax"))
    keep = ~g["bNoData"]
    with gx.DeviceMap.from_host(T) as dm:
        dm.set_mask(keep)
        r = dm.kr(g["kwargs"].get("tol", 1e-12), 0.1, 3)
    np.testing.assert_allclose(r["x"][keep], g["x"], rtol=BIAS_RTOL)
    assert np.all(r["x"][~keep] == 0)
    assert (r["outer"], r["MVP"]) == (int(g["outer"]), int(g["MVP"]))


@pytest.mark.parametrize("name", golden_names("VC"))
def test_vc_golden_dropin(gx, name):
    from gcmapexplorer_b200 import ccmap as cmp, normalizer as nz
    g = load_golden(name)
    c = cmp.CCMAP()
    c.gen_from_matrix(g["input"], 100000, xlabel="chrS", ylabel="chrS")
    c.make_unreadable()
    out = nz.normalizeCCMapByVCNorm(c, **g["kwargs"])
    assert out.matrix is not None
    got = np.array(out.matrix)
    assert np.array_equal(out.bNoData, g["bNoData"])
    assert np.array_equal(got, g["matrix"])                                      # integer counts: bit-exact fp32
    assert np.float32(out.minvalue) == np.float32(g["minvalue"])
    assert np.float32(out.maxvalue) == np.float32(g["maxvalue"])


def test_inner_seam_functions(gx):
    """The normalizeCore / normalizeKnightRuiz shims keep the reference's calling conventions."""
    from gcmapexplorer_b200 import normalizeCore as core, normalizeKnightRuiz as krn
    A = onp.synth_map(211, seed=5)
    keep = onp.get_nonzeros_index(A)
    sub = np.ascontiguousarray(A[keep][:, keep])
    M = sub.copy()
    assert core.performIterativeCorrection(M, 1e-4, 500) is None                 # in place, returns None
    M64, B, passes, _ = onp.ic_fp64(sub, 1e-4, 500)
    assert rel_err(M, M64.astype(np.float32)) < MAT_RTOL and core.last_ic_info["passes"] == passes
    out = A.copy()
    assert core.performVCNormalization(A, Out=out, sqroot=False, bNoData=~keep) is None
    assert np.array_equal(out, onp.vc(A, bNoData=~keep, out=A.copy())[0])
    ret = core.performVCNormalization(A, sqroot=True, bNoData=~keep)
    exp = onp.vc(A, bNoData=~keep, sqroot=True, out=np.zeros_like(A))[0]
    assert np.array_equal(ret, exp)
    K = krn.KnightRuizNorm(sub, tol=1e-12, delta=0.1, Delta=3)
    O = np.zeros_like(A)
    mn, mx = K.run(sub, 0, O, ~keep)
    r = onp.kr(sub)
    assert (K.i, K.MVP) == (r["outer"], r["MVP"]) and K.x.shape == (sub.shape[0], 1)
    np.testing.assert_allclose(K.x[:, 0], r["x"], rtol=BIAS_RTOL)
    exp = onp.scatter_symmetric(onp.kr_matrix(sub, r["x"]), ~keep, np.zeros_like(A))
    assert rel_err(O, exp) < MAT_RTOL
    with pytest.raises(ValueError):
        krn.KnightRuizNorm(sub, memory="TAPE")
    # remove_zeros: device mask + device compaction into a float64 memmap (lib/ccmapHelpers.pyx:241-301)
    from gcmapexplorer_b200 import ccmapHelpers as cmh
    for kw in (dict(), dict(threshold_percentile=95)):
        Z, bNo = cmh.remove_zeros(A, **kw)
        k2 = onp.get_nonzeros_index(A, kw.get("threshold_percentile"), None)
        assert np.array_equal(bNo, ~k2) and Z.arr.dtype == np.float64
        assert np.array_equal(np.asarray(Z.arr), A[k2][:, k2].astype(np.float64))


@pytest.mark.skipif(not rl.have_cores(), reason="oracle/_ref not built")
def test_ic_kr_vs_live_reference_cores(gx):
    """n=1200 through the reference's own compiled cores (oracle/_ref) on the same map."""
    core, krmod, cmh = rl.load_cores()
    A = onp.synth_map(1200, seed=12)
    keep = onp.get_nonzeros_index(A)
    sub = np.ascontiguousarray(A[keep][:, keep])
    M = sub.copy()
    core.performIterativeCorrection(M, 1e-4, 500)
    r = _ic_device(gx, sub, 1e-4, 500, np.ones(sub.shape[0], dtype=bool))
    assert rel_err(r["matrix"], M) < ASIS_RTOL

    class F(np.ndarray):
        def flush(self):
            pass
    K = krmod.KnightRuizNorm(sub, tol=1e-12, delta=0.1, Delta=3)
    O = np.zeros_like(A).view(F)
    K.run(sub, 0, O, ~keep)
    with gx.DeviceMap.from_host(A) as dm:
        dm.set_mask(keep)
        k = dm.kr(1e-12, 0.1, 3)
        dm.apply_sym(None, masked_mode=0, in_place=False)
        got = dm.download(1)
    assert (k["outer"], k["MVP"]) == (K.i, K.MVP)
    np.testing.assert_allclose(k["x"][keep], np.asarray(K.x)[:, 0], rtol=BIAS_RTOL)
    assert rel_err(got, np.asarray(O)) < MAT_RTOL


@pytest.mark.parametrize("variant", [0, 1, 2, 6])
def test_matvec_variants_agree(gx, variant, monkeypatch):
    """Every compiled (rows-per-band, occupancy) variant of k_matvec gives the same IC result."""
    monkeypatch.setenv("GCX_MATVEC_VARIANT", str(variant))
    A = onp.synth_map(777, seed=21)
    M64, B, passes, _ = onp.ic_fp64(A, 1e-4, 500)
    r = _ic_device(gx, A, 1e-4, 500, np.ones(777, dtype=bool))
    assert r["passes"] == passes
    np.testing.assert_allclose(r["bias"], B, rtol=BIAS_RTOL)


def test_device_synth_properties_and_full_size_invariants(gx):
    """Device-generated map at a size the oracle cannot reach quickly (n=12000): symmetry, empty bins,
    and size-independent invariants of the normalizations."""
    n = 12000
    with gx.DeviceMap(n) as dm:
        dm.synth(seed=110, scale=200.0, alpha=1.0, empty_frac=0.01)
        rs, zc = dm.row_stats()
        A_rows = dm.download(0)
        assert np.array_equal(A_rows[:200, :], A_rows[:, :200].T)                 # symmetric
        empty = rs == 0
        assert 0.005 * n < empty.sum() < 0.02 * n
        keep = ~empty
        dm.set_mask(np.ones(n, dtype=bool))
        r = dm.ic(1e-4, 500)
        assert 10 < r["passes"] < 200
        dm.apply_sym(None, masked_mode=1, in_place=False)
        dm.swap()
        rs_ic, _ = dm.row_stats()
        # IC invariants: total conserved, kept row sums all equal (lib/normalizeCore.pyx:36,53)
        assert abs(rs_ic.sum() - rs.sum()) < 1e-6 * rs.sum()
        k = rs_ic[keep]
        assert (k.max() - k.min()) / k.mean() < 1e-3
        dm.swap()
        dm.set_mask(keep)
        kk = dm.kr(1e-12, 0.1, 3)
        assert kk["rout"] <= 1e-24 and kk["MVP"] < 200
        dm.apply_sym(None, masked_mode=0, in_place=False)
        dm.swap()
        rs_kr, _ = dm.row_stats()
        assert np.max(np.abs(rs_kr[keep] - 1.0)) < 1e-5 and np.all(rs_kr[~keep] == 0)   # fp32 storage
        # MCFS invariants on the KR-balanced (non-integer) map: after o/e scaling every diagonal's statistic is 1
        for stats in ("mean", "median"):
            avg = dm.diag_stats(stats)
            dm.diag_apply(None, subtract=False, in_place=False)
            dm.swap()
            after = dm.diag_stats(stats)
            dm.swap()
            has = avg != 0
            assert has.sum() > n // 2
            assert np.max(np.abs(after[has] - 1.0)) < (1e-6 if stats == "mean" else 2e-7)
            assert np.all(after[~has] == 0)


# --------------------------------------------------------------------------------------------------
# distance-frequency (MCFS)
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,seed,dense", [(5, 0, True), (64, 1, True), (97, 2, False), (300, 3, True), (1111, 4, False)])
@pytest.mark.parametrize("stats", ["median", "mean"])
def test_diag_stats_vs_oracle(gx, n, seed, dense, stats):
    if n == 5:
        A = load_golden("kat5_ic")["input"]
    else:
        A = onp.synth_map(n, seed=seed) if dense else onp.synth_map(n, seed=seed, scale=30.0, alpha=1.3, empty_frac=0.03)
    exp = onp.avg_contact_by_distance(A, stats)
    with gx.DeviceMap.from_host(A) as dm:
        got = dm.diag_stats(stats)
    if stats == "median":
        assert np.array_equal(got, exp)                      # selection: bit-exact
    else:
        np.testing.assert_allclose(got, exp, rtol=1e-13, atol=0)


def test_diag_median_non_integer_and_negative(gx):
    """Median select on arbitrary floats (already-normalized maps, negative o-e values, duplicates)."""
    rng = np.random.default_rng(5)
    n = 257
    L = rng.standard_normal((n, n)).astype(np.float32)
    L[rng.random((n, n)) < 0.3] = 0.0
    L[rng.random((n, n)) < 0.2] = np.float32(0.5)          # many duplicates
    A = np.tril(L) + np.tril(L, -1).T
    exp = onp.avg_contact_by_distance(A, "median")
    with gx.DeviceMap.from_host(A) as dm:
        got = dm.diag_stats("median")
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_diag_median_early_exit_levels(gx, seed):
    """The radix select ends after two passes on diagonals whose candidates are short integers and runs all four on the others;
    the upper middle of an even count comes from the selected bin, the next bin of the same histogram, or the next bucket one,
    two or three levels up.  Few entries per diagonal and values spread over many exponents hit every one of those cases."""
    rng = np.random.default_rng(100 + seed)
    n = 96 if seed < 2 else 333
    pools = [np.array([1, 2, 3, 4, 7, 8], np.float32), np.array([1, 100, 1000, 1023, 1024, 1025], np.float32),
             np.array([5000, 65537, 3, 16777215], np.float32), np.array([0.5, 1e-3, 2.0, 1536.0, 1536.25], np.float32),
             np.array([-3, -1, 2, 1024, -1025.5], np.float32)]
    L = np.zeros((n, n), np.float32)
    for d in range(n):
        pool = pools[rng.integers(len(pools))] if seed % 2 else np.concatenate(pools)
        vals = rng.choice(pool, size=n - d)
        vals[rng.random(n - d) < 0.35] = 0.0
        L[np.arange(d, n), np.arange(0, n - d)] = vals
    A = np.tril(L) + np.tril(L, -1).T
    exp = onp.avg_contact_by_distance(A, "median")
    with gx.DeviceMap.from_host(A) as dm:
        got = dm.diag_stats("median")
    assert np.array_equal(got, exp)
    # the same through the batched and the pooled entry points (one word of "pending" state for the whole batch)
    B = onp.synth_map(200, seed=seed)
    with gx.DeviceMap.from_host(A) as da, gx.DeviceMap.from_host(B) as db:
        gx.MapBatch([da, db]).mcfs("median", subtract=False)
        assert np.array_equal(da.diag_avg(), exp)
        assert np.array_equal(db.diag_avg(), onp.avg_contact_by_distance(B, "median"))


@pytest.mark.parametrize("subtract", [False, True])
def test_diag_apply_exactness(gx, subtract):
    """o/e divides in float64 and stores float32 (lib/normalizeCore.pyx:105-117); the kernel reaches the same bits through a tabulated
    reciprocal and one correction step, falling back to the division where the two could round differently.  Adversarial input:
    per distance d a value x_d and a divisor chosen so that x_d / av_d lands within an ulp of the MIDPOINT between two floats
    (the only place a last-bit error of the float64 quotient can change the float32 result), besides random values, huge and tiny
    quotients and zero divisors."""
    rng = np.random.default_rng(77)
    n = 1536
    x = np.ldexp(rng.integers(1 << 23, 1 << 24, size=n).astype(np.float64), rng.integers(-30, 8, size=n)).astype(np.float32)
    f = np.ldexp(rng.integers(1 << 23, 1 << 24, size=n).astype(np.float64), rng.integers(-24, 0, size=n)).astype(np.float32)
    mid = f.astype(np.float64) + 0.5 * np.spacing(f).astype(np.float64)          # exactly representable in float64
    av = x.astype(np.float64) / mid
    av[::7] = rng.uniform(0.1, 900.0, size=av[::7].size)                           # ordinary divisors
    av[5::97] = 0.0                                                                # distances without data
    av[11::101] = 1e-300                                                           # reciprocal overflows
    av[13::103] = 1e300                                                            # quotient below the float32 range
    i, j = np.indices((n, n))
    d = np.abs(i - j)
    A = x[d].copy()
    flip = rng.random((n, n)) < 0.3
    A[flip] = rng.integers(1, 2000, size=int(flip.sum())).astype(np.float32)
    A[rng.random((n, n)) < 0.4] = 0.0
    A = np.tril(A) + np.tril(A, -1).T
    with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
        if subtract:
            t = A.astype(np.float64) - av[d]
            t[t < 0.0] = 1.0
            exp = t.astype(np.float32)
        else:
            exp = (A.astype(np.float64) / av[d]).astype(np.float32)
    keep = (A == 0) | (av[d] == 0.0)
    exp[keep] = A[keep]
    with gx.DeviceMap.from_host(A) as dm:
        dm.diag_apply(av, subtract=subtract, in_place=False)
        got = dm.download(1)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))
    if not subtract:
        q = A.astype(np.float64) / np.where(av[d] == 0.0, 1.0, av[d])
        near = np.abs((q.view(np.uint64) & np.uint64(0x1fffffff)).astype(np.int64) - 0x10000000) <= 2
        assert near[~keep].sum() > 1000                     # the fallback really ran


@pytest.mark.parametrize("name", golden_names("MCFS"))
def test_mcfs_golden_dropin(gx, name):
    from gcmapexplorer_b200 import ccmap as cmp, normalizer as nz
    g = load_golden(name)
    c = cmp.CCMAP()
    c.gen_from_matrix(g["input"], 100000, xlabel="chrS", ylabel="chrS")
    c.make_unreadable()
    out = nz.normalizeCCMapByMCFS(c, **g["kwargs"])
    assert out is not None and out.matrix is not None and c.matrix is not None     # lib/normalizer.py:851-857
    got = np.array(out.matrix)
    assert np.array_equal(out.bNoData, g["bNoData"])
    if "avg" in g:
        if g["kwargs"].get("stats", "median") == "median":
            assert np.array_equal(nz.last_run_info["avg"], g["avg"])
        else:
            np.testing.assert_allclose(nz.last_run_info["avg"], g["avg"], rtol=1e-13)
    assert rel_err(got, g["matrix"]) < MAT_RTOL
    assert out.minvalue == pytest.approx(float(g["minvalue"]), rel=MAT_RTOL)
    assert out.maxvalue == pytest.approx(float(g["maxvalue"]), rel=MAT_RTOL)


def test_mcfs_bad_stype_and_inner_seam(gx):
    from gcmapexplorer_b200 import ccmap as cmp, normalizer as nz, normalizeCore as core, cmstats
    A = onp.synth_map(150, seed=9)
    c = cmp.CCMAP()
    c.gen_from_matrix(A, 100000)
    assert nz.normalizeCCMapByMCFS(c, stype="bogus") is None
    c.make_readable()
    avg = cmstats.getAvgContactByDistance(c, stats="mean")
    np.testing.assert_allclose(avg, onp.avg_contact_by_distance(A, "mean"), rtol=1e-13)
    assert rel_err(core.normalizeByAvgContactByDivision(A, avg), onp.mcfs_divide(A, avg)) < MAT_RTOL
    out = np.zeros_like(A)
    assert core.normalizeByAvgContactBySubtraction(A, avg, Out=out) is None
    assert rel_err(out, onp.mcfs_subtract(A, avg)) < MAT_RTOL
    with pytest.raises(NotImplementedError):
        cmstats.getAvgContactByDistance(c, removeOutliers=True)
    # pooled list (lib/cmstats.py:572-578): maps of different sizes, the d-th diagonals pooled
    B = onp.synth_map(A.shape[0] + 37, seed=77, scale=30.0)
    c2 = cmp.CCMAP()
    c2.gen_from_matrix(B, 100000)
    c2.make_readable()
    for st in ("median", "mean"):
        got = cmstats.getAvgContactByDistance([c, c2, c], stats=st)
        exp = np.zeros(B.shape[0])
        for d in range(B.shape[0]):
            vals = np.concatenate([np.diagonal(M, -d)[np.diagonal(M, -d) != 0] for M in (A, B, A) if d < M.shape[0]]).astype(np.float64)
            if vals.size:
                exp[d] = np.median(vals) if st == "median" else np.mean(vals)
        if st == "median":
            assert np.array_equal(got, exp)
        else:
            np.testing.assert_allclose(got, exp, rtol=1e-13)
    with pytest.raises(TypeError):
        cmstats.getAvgContactByDistance(A)


def test_ccmap_save_load_roundtrip_with_outfile(gx, tmp_path):
    """outFile= path: result saved as .ccmap + gz npbin (lib/normalizer.py:621-622), loads back identically."""
    from gcmapexplorer_b200 import ccmap as cmp, normalizer as nz
    A = onp.synth_map(120, seed=4)
    c = cmp.CCMAP()
    c.gen_from_matrix(A, 50000, xlabel="chr9", ylabel="chr9", workDir=str(tmp_path))
    c.make_unreadable()
    assert nz.normalizeCCMapByIC(c, outFile=str(tmp_path / "n_ic")) is None
    direct = nz.normalizeCCMapByIC(c)
    loaded = cmp.load_ccmap(str(tmp_path / "n_ic.ccmap"), workDir=str(tmp_path))
    loaded.make_readable()
    assert np.array_equal(np.array(loaded.matrix), np.array(direct.matrix))
    assert np.array_equal(loaded.bNoData, direct.bNoData) and loaded.binsize == 50000 and loaded.xlabel == "chr9"
    # a path instead of an object, and a missing file (lib/ccmap.py:644-647 -> TypeError at the unpack)
    again = nz.normalizeCCMapByKR(str(tmp_path / "n_ic.ccmap"), workDir=str(tmp_path))
    assert again is not None
    with pytest.raises(TypeError):
        nz.normalizeCCMapByIC(str(tmp_path / "missing.ccmap"))


# --------------------------------------------------------------------------------------------------
# edge cases: tiny / ragged sizes, error conventions, residency variants
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 31, 32, 33, 255, 257])
def test_tiny_and_ragged_sizes(gx, n):
    rng = np.random.default_rng(100 + n)
    U = np.triu(rng.integers(1, 50, size=(n, n)).astype(np.float32))
    A = U + np.triu(U, 1).T                     # dense, positive: every pad / tail path, no empty rows
    keep = np.ones(n, dtype=bool)
    r = _ic_device(gx, A, 1e-6, 500, keep)
    M64, B, passes, _ = onp.ic_fp64(A, 1e-6, 500)
    assert r["passes"] == passes
    np.testing.assert_allclose(r["bias"], B, rtol=BIAS_RTOL)
    assert rel_err(r["matrix"], M64.astype(np.float32)) < MAT_RTOL
    with gx.DeviceMap.from_host(A) as dm:
        dm.set_mask(keep)
        k = dm.kr(1e-12, 0.1, 3)
        dm.apply_sym(None, 0, False)
        D = dm.download(1)
        mn, mx, csum = dm.vc(False, False)
        V = dm.download(1)
        med = dm.diag_stats("median")
        mean = dm.diag_stats("mean")
    ko = onp.kr(A)
    assert (k["outer"], k["MVP"]) == (ko["outer"], ko["MVP"])
    np.testing.assert_allclose(k["x"], ko["x"], rtol=BIAS_RTOL)
    assert rel_err(D, onp.scatter_symmetric(onp.kr_matrix(A, ko["x"]), ~keep, np.zeros_like(A))) < MAT_RTOL
    assert np.array_equal(V, onp.vc(A)[0])
    assert np.array_equal(med, onp.avg_contact_by_distance(A, "median"))
    np.testing.assert_allclose(mean, onp.avg_contact_by_distance(A, "mean"), rtol=1e-13)


def test_error_conventions(gx, caplog):
    """IC swallows non-device errors and returns None (lib/normalizer.py:642-647); KR logs and re-raises (:390-395)."""
    from gcmapexplorer_b200 import ccmap as cmp, normalizer as nz
    A = onp.synth_map(60, seed=8)
    c = cmp.CCMAP()
    c.gen_from_matrix(A, 100000)
    c.make_unreadable()
    assert nz.normalizeCCMapByIC(c, percentile_threshold_no_data=90, threshold_data_occup=0.5) is None     # AssertionError inside
    assert nz.normalizeCCMapByIC(c, threshold_data_occup=1.5) is None                                     # ValueError inside
    with pytest.raises(AssertionError):
        nz.normalizeCCMapByKR(c, percentile_threshold_no_data=90, threshold_data_occup=0.5)
    with pytest.raises(ValueError):
        nz.normalizeCCMapByKR(c, memory="TAPE")
    assert nz.normalizeCCMapByVCNorm(c, threshold_data_occup=1.5) is None
    assert nz.normalizeCCMapByMCFS(c, threshold_data_occup=1.5) is None
    with pytest.raises((ImportError, OSError)):
        nz.normalizeGCMapByIC("in.gcmap", "out.gcmap")        # HDF5 container: needs h5py at call time (absent here) / an existing file


def test_signed_map_float32_emptiness_rule(gx):
    """lib/ccmapHelpers.pyx:192 calls a row empty when its FLOAT32 sum is 0.0.  On a signed map (o-e output, correlation map) a
    row can cancel in float32 but not in float64 (2^24 + 1 - 2^24): the reference masks it; so does the drop-in (the device flags rows
    with negative entries, the host repeats NumPy's float32 sum on those); and a row that cancels exactly in both is masked too."""
    from gcmapexplorer_b200 import ccmapHelpers as cmh
    n = 64
    A = onp.synth_map(n, seed=2, scale=50.0, empty_frac=0.0)
    A[7, :] = 0; A[:, 7] = 0
    A[7, 0], A[7, 8], A[7, 16] = 16777216.0, 1.0, -16777216.0         # NumPy's float32 sum (8 interleaved accumulators): 0.0 ; float64 sum: 1.0
    A[9, :] = 0; A[:, 9] = 0
    A[9, 4], A[9, 5] = 3.0, -3.0                                       # cancels exactly in any precision
    exp = onp.get_nonzeros_index(A)                                    # NumPy float32 row sums, like the reference
    assert not exp[7] and not exp[9]
    got = cmh.get_nonzeros_index(A)
    assert np.array_equal(got, exp)
    with gx.DeviceMap.from_host(A) as dm:
        rs, _ = dm.row_stats()
        flags = dm.row_negative_flags()
        assert rs[7] == 1.0 and flags[7] == 1 and flags[9] == 1 and flags.sum() == 2
        assert np.array_equal(cmh.device_mask(dm, host_rows=lambda i: A[i]), exp)
        assert cmh.device_mask(dm)[7]                                  # without the host copy the float64 rule stands (documented)


def test_device_resident_twin_adopt_and_export(gx):
    """when the map is already in HBM (a torch tensor) it is adopted / exported device to device -- same results as
    the host route, no PCIe trip."""
    import torch
    n = 1500
    A = onp.synth_map(n, seed=6)
    t = torch.from_numpy(A).cuda()
    with gx.DeviceMap(n) as dm, gx.DeviceMap.from_host(A) as ref:
        dm.adopt_tensor(t)
        assert dm.checksum(0) == ref.checksum(0)
        keep = onp.get_nonzeros_index(A)
        for d in (dm, ref):
            d.set_mask(keep)
            d.kr(1e-12, 0.1, 3.0)
            d.apply_sym(None, masked_mode=0, in_place=False)
        out = dm.export_tensor(1)
        assert out.is_cuda and np.array_equal(out.cpu().numpy(), ref.download(1))
        pitched = torch.zeros((n, n + 12), dtype=torch.float32, device="cuda")          # strided views work too
        pitched[:, :n] = t
        dm.adopt_tensor(pitched[:, :n])
        assert dm.checksum(0) == ref.checksum(0)


@pytest.mark.parametrize("mb", [256, 768])
def test_own_int8_tensor_core_product_is_exact(gx, mb):
    """k_i8gemm_nt (tcgen05 kind::i8, TMEM accumulators, TMA-staged swizzled tiles) against NumPy integer arithmetic: every entry
    of T = L R^T exact, full int8 range, and the lower-triangle-only mode of a symmetric product."""
    rng = np.random.default_rng(mb)
    L = rng.integers(-128, 128, size=(mb, mb), dtype=np.int8)
    R = rng.integers(-128, 128, size=(mb, mb), dtype=np.int8)
    with gx.DeviceMap(8) as dm:
        T, _ = dm.i8gemm(L, R)
        assert np.array_equal(T, L.astype(np.int32) @ R.astype(np.int32).T)
        S, _ = dm.i8gemm(L, L, symmetric=True)
        il = np.tril_indices(mb)
        assert np.array_equal(S[il], (L.astype(np.int32) @ L.astype(np.int32).T)[il])


def test_correlation_own_product_equals_cublas_route(gx):
    """The correlation matrix with the int8 products on the library's own kernel (default) is bit-identical to the cuBLAS route."""
    n = 700
    A = onp.synth_map(n, seed=12)
    keep = onp.get_nonzeros_index(A)
    outs = []
    for own in (1, 0):
        with gx.DeviceMap(n, options={"corr_own_gemm": own}) as dm:
            dm.upload(A)
            dm.set_mask(keep)
            dm.corr(0.0)
            outs.append((dm.download(1), dm.corr_last_route()))
            dm.scale(0.37)                                   # float32-valued map: the per-row-exponent slices
            dm.corr(0.0)
            outs.append((dm.download(1), dm.corr_last_route()))
    assert outs[0][1][0] > 0 and outs[1][1][0] < 0           # integer slices / float slices were taken
    assert np.array_equal(outs[0][0], outs[2][0]) and np.array_equal(outs[1][0], outs[3][0])
    assert outs[0][1] == outs[2][1] and outs[1][1] == outs[3][1]


def test_hic_norm_vector_application(gx):
    """SURVEY.md 8f row N4, second half: the norm-vector arithmetic of the .hic importer (lib/hic2gcmap.py:82-98) on a resident map,
    bit-identical to the reference's per-record Python arithmetic incl. inf on zero norms and NaN vectors entries."""
    from gcmapexplorer_b200 import normalizeCore as core
    n = 131
    A = onp.synth_map(n, seed=8, scale=25.0)
    rng = np.random.default_rng(3)
    c1 = rng.uniform(0.3, 2.5, n)
    c2 = rng.uniform(0.3, 2.5, n)
    c1[[5, 40]] = 0.0                      # zero norm -> inf for the stored counts of that bin
    c2[17] = np.nan                        # .hic files mark unusable bins with NaN
    with np.errstate(invalid="ignore"):
        exp = onp.hic_norm_apply(A, c1, c2)
    got = core.applyNormVectors(A, c1, c2)
    assert np.array_equal(got, exp, equal_nan=True)
    assert np.isinf(got[5][A[5] != 0]).all() and np.isnan(got[:, 17][A[:, 17] != 0]).all() and (got[A == 0] == 0).all()
    out = np.zeros_like(A)
    assert core.applyNormVectors(A, c1, c1, Out=out) is None            # symmetric use: KR / VC vectors of an intra-chromosomal map
    with np.errstate(invalid="ignore"):
        assert np.array_equal(out, onp.hic_norm_apply(A, c1, c1), equal_nan=True)


def test_streaming_residency_file_rows_and_gzip(gx, tmp_path):
    """the three streaming routes into HBM (raw .npbin through parallel pread, row chunks, a gzip stream) leave
    the same map and the same row statistics as a plain upload, the statistics arrive with the upload (no extra pass), and
    gcx_map_download_file writes the bytes np.memmap would."""
    from gcmapexplorer_b200 import ccmap as cmp, normalizer as nz
    n = 1237
    A = onp.synth_map(n, seed=5)
    raw = tmp_path / "a.npbin"
    A.tofile(raw)
    with gx.DeviceMap.from_host(A) as ref:
        rs0, zc0 = ref.row_stats()
        c0 = ref.checksum(0)
    with gx.DeviceMap(n) as dm:
        dm.upload_file(str(raw))
        l0 = dm.launch_count()
        rs, zc = dm.row_stats()
        assert dm.launch_count() == l0                      # statistics came with the upload
        assert dm.checksum(0) == c0 and np.array_equal(rs, rs0) and np.array_equal(zc, zc0)
        dm.download_file(0, str(tmp_path / "b.npbin"))
        assert np.array_equal(np.fromfile(tmp_path / "b.npbin", dtype=np.float32).reshape(n, n), A)
        for r0 in range(0, n, 300):                           # row chunks, as a gzip / HDF5 reader hands them over
            dm.upload_rows(A[r0:r0 + 300] * 2.0, r0)
        dm.upload_finish()
        rs2, zc2 = dm.row_stats()
        assert np.array_equal(rs2, 2.0 * rs0) and np.array_equal(zc2, zc0)
    # a compressed .ccmap given by file name is inflated straight into the device: same result as the CCMAP object route
    c = cmp.CCMAP()
    c.gen_from_matrix(A, 10000, xlabel="chrZ", ylabel="chrZ", workDir=str(tmp_path))
    cmp.save_ccmap(c, str(tmp_path / "z.ccmap"), compress=True)
    lazy = cmp.load_ccmap(str(tmp_path / "z.ccmap"), workDir=str(tmp_path), lazy_gz=True)
    assert lazy._gz_source.endswith("z.npbin.gz") and lazy.path2matrix == ""
    a = nz.normalizeCCMapByIC(str(tmp_path / "z.ccmap"), workDir=str(tmp_path))
    b = nz.normalizeCCMapByIC(c)
    assert np.array_equal(np.array(a.matrix), np.array(b.matrix)) and np.array_equal(a.bNoData, b.bNoData)
    lazy.make_readable()                                      # and a lazily opened map can still be looked at
    assert np.array_equal(np.array(lazy.matrix), A)


def test_batch_driver_over_a_map_set(gx, tmp_path):
    """normalizeGCMapByIC / ByKR / ByMCFS over a map-set directory (the HDF5-free carrier of a .gcmap, mapset.py): maps are
    processed smallest first (lib/gcmap.py:219-262), each result equals the per-map oracle, IC / KR results are stored with their
    x2 'sum' pyramid down to < 500 bins (lib/gcmap.py:582-604), MCFS normalizes every stored resolution without coarsening."""
    from gcmapexplorer_b200 import ccmap as cmp, normalizer as nz, mapset
    inputs = {"chrB": onp.synth_map(482, seed=3), "chrA": onp.synth_map(1101, seed=4), "chrC": load_golden("s97_ic")["input"]}
    src = tmp_path / "in"
    for name, A in inputs.items():
        c = cmp.CCMAP()
        c.gen_from_matrix(A, 10000, xlabel=name, ylabel=name, workDir=str(tmp_path))
        c.bNoData = ~onp.get_nonzeros_index(A)
        mapset.add_map(c, str(src), generateCoarse=(name == "chrA"), compress_files=(name != "chrB"))
    ms = mapset.MapSet(str(src))
    assert ms.names_smallest_first() == ["chrC", "chrB", "chrA"]
    assert ms.resolutions("chrA") == ["10kb", "20kb", "40kb"] and ms.resolutions("chrB") == ["10kb"]
    order = []
    real_add = mapset.add_map

    def spy(cm, dest, **kw):
        order.append(cm.xlabel)
        return real_add(cm, dest, **kw)
    mapset.add_map = spy
    try:
        nz.normalizeGCMapByIC(str(src), str(tmp_path / "out_ic"), tol=1e-4, workDir=str(tmp_path))
        nz.normalizeGCMapByKR(str(src), str(tmp_path / "out_kr"), workDir=str(tmp_path))
        nz.normalizeGCMapByMCFS(str(src), str(tmp_path / "out_mcfs"), workDir=str(tmp_path))
    finally:
        mapset.add_map = real_add
    assert order[:3] == ["chrC", "chrB", "chrA"] and order[3:6] == ["chrC", "chrB", "chrA"]
    assert order[6:] == ["chrC", "chrB", "chrA", "chrA", "chrA"]             # MCFS: every stored resolution of chrA
    out_ic, out_kr, out_m = (mapset.MapSet(str(tmp_path / d)) for d in ("out_ic", "out_kr", "out_mcfs"))
    for name, A in inputs.items():
        got = out_ic.load(name, workDir=str(tmp_path))
        got.make_readable()
        ref = onp.normalize_ic(A, tol=1e-4, fp64=True)
        assert rel_err(np.array(got.matrix), ref["matrix"]) < MAT_RTOL and np.array_equal(got.bNoData, ref["bNoData"])
        gk = out_kr.load(name, workDir=str(tmp_path))
        gk.make_readable()
        rk = onp.normalize_kr(A)
        assert rel_err(np.array(gk.matrix), rk["matrix"]) < MAT_RTOL and np.array_equal(gk.bNoData, rk["bNoData"])
    # pyramid of the largest map: 1101 -> 551 -> 276 bins, each level the 'sum' coarsening of the level before
    assert out_ic.resolutions("chrA") == ["10kb", "20kb", "40kb"] and out_ic.resolutions("chrB") == ["10kb"]
    fine = out_ic.load("chrA", workDir=str(tmp_path))
    fine.make_readable()
    lvl = np.array(fine.matrix)
    for res in ("20kb", "40kb"):
        cm = out_ic.load("chrA", resolution=res, workDir=str(tmp_path))
        cm.make_readable()
        lvl = onp.downsample_2d(lvl, 2, "sum")
        assert cm.shape == lvl.shape and np.array_equal(np.array(cm.matrix), lvl) and cm.binsize == util_res(res)
    # MCFS walked the stored resolutions of chrA itself (no coarsening of the result)
    for res in ("10kb", "20kb", "40kb"):
        a_in = ms.load("chrA", resolution=res, workDir=str(tmp_path))
        a_in.make_readable()
        m = out_m.load("chrA", resolution=res, workDir=str(tmp_path))
        m.make_readable()
        refm = onp.normalize_mcfs(np.array(a_in.matrix))
        assert rel_err(np.array(m.matrix), refm["matrix"]) < MAT_RTOL


def util_res(res):
    from gcmapexplorer_b200 import util
    return util.resolutionToBinsize(res)


def test_reupload_invalidates_cached_row_statistics(gx):
    A = onp.synth_map(200, seed=1)
    B = onp.synth_map(200, seed=2)
    with gx.DeviceMap.from_host(A) as dm:
        dm.set_mask(onp.get_nonzeros_index(A))
        dm.vc(False, False)
        dm.upload(B)                                   # same context, new map
        keepB = onp.get_nonzeros_index(B)
        dm.set_mask(keepB)
        dm.vc(False, False)
        assert np.array_equal(dm.download(1), onp.vc(B, bNoData=~keepB, out=B.copy())[0])
        with pytest.raises(gx.GcxError):
            dm.apply_sym(None, 0, False)               # no solver has run on the new map


def test_failed_map_allocation_leaves_the_context_usable(gx):
    """A map that cannot fit (n = 600,000 -> 1.4 TB) fails with a message, frees everything, and the same context
    then serves a normal map; misuse of the mask entry point is an error, not a crash."""
    import ctypes as C
    from gcmapexplorer_b200 import _lib
    A = onp.synth_map(150, seed=4)
    with gx.DeviceMap(150) as dm:
        rc = _lib.lib().gcx_map_create(dm._ctx, 600000, 0, 600000)
        assert rc != 0 and b"cudaMalloc" in _lib.lib().gcx_last_error()
        with pytest.raises(gx.GcxError):
            dm.row_stats()                                            # no resident map any more
        _lib.check(_lib.lib().gcx_map_create(dm._ctx, 150, 0, 150))
        dm.upload(A)
        assert _lib.lib().gcx_set_mask(dm._ctx, None) != 0
        keep = onp.get_nonzeros_index(A)
        dm.set_mask(keep)
        assert dm.ic(1e-4, 500)["passes"] == onp.ic_fp64(A[np.ix_(keep, keep)].astype(np.float64), 1e-4, 500)[2]


def test_non_float32_and_non_contiguous_inputs(gx):
    A = onp.synth_map(90, seed=2)
    with gx.DeviceMap.from_host(A.astype(np.float64)) as dm:         # converted on the host
        assert np.array_equal(dm.download(0), A)
    F = np.asfortranarray(A)
    with gx.DeviceMap.from_host(F) as dm:
        assert np.array_equal(dm.download(0), A)


def test_in_place_equals_out_of_place_and_async_download(gx):
    A = onp.synth_map(500, seed=6)
    keep = onp.get_nonzeros_index(A)
    with gx.DeviceMap.from_host(A) as dm:
        dm.set_mask(keep)
        dm.kr(1e-12, 0.1, 3)
        m1 = dm.apply_sym(None, 0, in_place=False)
        out = dm.download(1)
        pin = gx.PinnedArray((500, 500))
        dm.download_async(1, pin.array)
        dm.vc(False, in_place=False)              # overwrites Out: must wait for the async copy on the device
        dm.download_wait()
        assert np.array_equal(pin.array, out)
        m2 = dm.apply_sym(None, 0, in_place=True)
        assert m1 == m2 and np.array_equal(dm.download(0), out)
        with pytest.raises(gx.GcxError):
            dm.download_async(0, np.empty((500, 500), dtype=np.float32))     # pageable memory is refused
        pin.free()


@pytest.mark.parametrize("stats", ["median", "mean"])
@pytest.mark.parametrize("subtract", [False, True])
def test_mcfs_batch_matches_per_map_calls(gx, stats, subtract):
    """gcx_mcfs_batch (one grid per stage over many maps, BASELINE configs[3]) against diag_stats + diag_apply map by map:
    statistics, output maps (checksums) and min/max bit-identical; ragged sizes incl. tiny maps."""
    sizes = [5, 97, 300, 257, 1111, 2500, 64, 1926]
    maps, single = [], []
    try:
        for k, n in enumerate(sizes):
            A = onp.synth_map(n, seed=100 + k, scale=40.0, alpha=1.1, empty_frac=0.02) if n > 5 else load_golden("kat5_ic")["input"]
            maps.append(gx.DeviceMap.from_host(A))
            single.append(gx.DeviceMap.from_host(A))
        mm = gx.MapBatch(maps).mcfs(stats, subtract=subtract)
        for k, (b, o) in enumerate(zip(maps, single)):
            avg = o.diag_stats(stats)
            mn, mx = o.diag_apply(None, subtract=subtract, in_place=False)
            assert np.array_equal(b.diag_avg(), avg), sizes[k]
            assert b.checksum(1) == o.checksum(1), sizes[k]
            assert np.array_equal(b.download(1), o.download(1))
            assert (np.float32(mm[k][0]) == np.float32(mn) or (np.isnan(mm[k][0]) and np.isnan(mn))) and np.float32(mm[k][1]) == np.float32(mx)
        # the batch leaves every map usable on its own stream afterwards
        again = maps[3].diag_stats(stats)
        assert np.array_equal(again, single[3].diag_stats(stats))
    finally:
        for d in maps + single:
            d.close()


def test_fused_apply_vc_matches_separate_passes(gx):
    """gcx_apply_sym_vc: the IC-scaled map and the VC map from one read of the input, bit-identical to gcx_apply_sym_scale + gcx_vc_run."""
    for n, seed in ((97, 1), (1031, 2), (2100, 3)):
        A = onp.synth_map(n, seed=seed)
        keep = onp.get_nonzeros_index(A)
        for solve_mask, mode in ((np.ones(n, dtype=bool), 1), (keep, 1), (keep, 0)):
            with gx.DeviceMap.from_host(A) as dm:
                dm.set_mask(solve_mask)
                dm.ic(1e-4, 500)
                (mn_s, mx_s), (mn_v, mx_v) = dm.apply_sym_vc(keep, masked_mode=mode, sqroot=False)
                S, V = dm.download(1), dm.download(2)
                a = dm.apply_sym(None, masked_mode=mode, in_place=False)
                S0 = dm.download(1)
                dm.set_mask(keep)
                b = dm.vc(False, in_place=False)
                V0 = dm.download(1)
            assert np.array_equal(S, S0) and np.array_equal(V, V0)
            assert (np.float32(mn_s), np.float32(mx_s)) == (np.float32(a[0]), np.float32(a[1]))
            assert (np.float32(mn_v), np.float32(mx_v)) == (np.float32(b[0]), np.float32(b[1]))


# --------------------------------------------------------------------------------------------------
# upper-triangle (symmetric) passes
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [2048, 3000, 4099, 6145])
def test_symmetric_pass_matches_dense_pass(gx, n, monkeypatch):
    """IC and KR through the upper-triangle matvec vs the dense matvec on the same device-generated map."""
    res = {}
    for sym in ("1", "0"):
        monkeypatch.setenv("GCX_SYM", sym)
        with gx.DeviceMap(n) as dm:
            dm.synth(seed=77, scale=150.0, alpha=1.0, empty_frac=0.01)
            rs, _ = dm.row_stats()
            keep = rs != 0
            dm.set_mask(np.ones(n, dtype=bool))
            ic = dm.ic(1e-4, 500)
            dm.set_mask(keep)
            kr = dm.kr(1e-12, 0.1, 3)
            res[sym] = (ic, kr, keep)
    (ic1, kr1, keep), (ic0, kr0, _) = res["1"], res["0"]
    assert ic1["symmetric"] and kr1["symmetric"] and not ic0["symmetric"]
    assert ic1["pass_bytes"] < (0.5 + 0.5 * 1024.0 / n + 0.03) * ic0["pass_bytes"]      # half + the 1024-column diagonal panels
    assert ic1["passes"] == ic0["passes"] and (kr1["outer"], kr1["MVP"]) == (kr0["outer"], kr0["MVP"])
    np.testing.assert_allclose(ic1["bias"], ic0["bias"], rtol=1e-12)
    np.testing.assert_allclose(kr1["x"][keep], kr0["x"][keep], rtol=1e-11)


def test_symmetric_pass_vs_oracle_and_asymmetric_fallback(gx):
    n = 2100
    A = onp.synth_map(n, seed=31)
    M64, B, passes, _ = onp.ic_fp64(A, 1e-4, 500)
    r = _ic_device(gx, A, 1e-4, 500, np.ones(n, dtype=bool))
    assert r["symmetric"] and r["passes"] == passes
    np.testing.assert_allclose(r["bias"], B, rtol=BIAS_RTOL)
    assert rel_err(r["matrix"], M64.astype(np.float32)) < MAT_RTOL
    A2 = A.copy()
    A2[5, 1900] += 1.0                       # break the symmetry in one entry
    r2 = _ic_device(gx, A2, 1e-4, 500, np.ones(n, dtype=bool))
    assert not r2["symmetric"]               # dense pass: row sums of the stored matrix, like every other dense run


def test_dropin_accepts_foreign_ccmap_objects_and_float64_maps(gx, tmp_path):
    """The normalizers take any object with the CCMAP attribute set (the reference's own class is one) and maps whose dtype is
    not float32: results equal the float32 route, the output keeps the input's dtype (no AttributeError / silent None)."""
    from gcmapexplorer_b200 import ccmap as cmp, normalizer as nz, statDist as sd
    A = onp.synth_map(150, seed=12)
    c = cmp.CCMAP()
    c.gen_from_matrix(A, 100000, xlabel="chrS", ylabel="chrS", workDir=str(tmp_path))
    c.make_unreadable()

    class Duck:
        def make_readable(self):
            self.matrix = np.memmap(self.path2matrix, dtype=self.dtype, mode="r", shape=tuple(self.shape))

        def make_unreadable(self):
            self.matrix = None
    d = Duck()
    for k in ("path2matrix", "xticks", "yticks", "binsize", "title", "xlabel", "ylabel", "shape", "minvalue", "maxvalue", "bNoData", "bLog", "dtype", "state"):
        setattr(d, k, getattr(c, k))
    d.matrix = None
    for fn in (nz.normalizeCCMapByKR, nz.normalizeCCMapByIC, nz.normalizeCCMapByVCNorm, nz.normalizeCCMapByMCFS, sd.transitionProbabilityMatrixForCCMap):
        a, b = fn(c), fn(d)
        assert a is not None and b is not None and isinstance(b, cmp.CCMAP)
        a.make_readable(); b.make_readable()
        assert np.array_equal(np.array(a.matrix), np.array(b.matrix)) and np.array_equal(a.bNoData, b.bNoData)
    c64 = cmp.CCMAP()
    c64.gen_from_matrix(A.astype(np.float64), 100000, xlabel="chrS", ylabel="chrS", workDir=str(tmp_path))
    o64, o32 = nz.normalizeCCMapByIC(c64), nz.normalizeCCMapByIC(c)
    assert o64 is not None and o64.matrix.dtype == np.float64
    o32.make_readable()
    assert np.array_equal(np.array(o64.matrix), np.array(o32.matrix).astype(np.float64))


@pytest.mark.parametrize("n", [2100, 4099, 6145])
def test_deferred_scalar_pass_matches_literal_order(gx, n):
    """The default upper-triangle IC pass defers the global scalars (k_ic_pass_symd: v_next = kappa / t, one grid barrier);
    option ic_defer=0 runs the literal order of lib/normalizeCore.pyx:42-53 (k_ic_pass_sym_tma).  Same pass count, same last
    L1, bias equal to rounding; the deferred pass launches exactly one more product."""
    res = {}
    for defer in (1, 0):
        with gx.DeviceMap(n, options={"ic_defer": defer}) as dm:
            dm.synth(seed=91, scale=150.0, alpha=1.0, empty_frac=0.01)
            rs, _ = dm.row_stats()
            for mask in (np.ones(n, dtype=bool), rs != 0):       # unfiltered (empty bins kept: the Phi branch) and filtered
                dm.set_mask(mask)
                res[(defer, bool(mask.all()))] = dm.ic(1e-4, 500)
            dm.set_mask(np.ones(n, dtype=bool))
            res[(defer, "cap")] = dm.ic(1e-4, 3)                 # iteration cap: 5 passes (lib/normalizeCore.pyx:55-60)
            res[(defer, "cap0")] = dm.ic(1e-4, 0)
    for key in (True, False, "cap", "cap0"):
        a, b = res[(1, key)], res[(0, key)]
        assert a["symmetric"] and b["symmetric"]
        assert a["passes"] == b["passes"] and a["matvecs"] == b["matvecs"] == a["passes"] + 1
        assert a["products_launched"] == a["passes"] + 2 and b["products_launched"] == b["passes"] + 1
        np.testing.assert_allclose(a["bias"], b["bias"], rtol=1e-12)
        assert abs(a["l1"] - b["l1"]) <= 1e-7 * b["l1"] + 1e-18      # a sum of differences of nearly equal numbers: rounding is amplified
    assert res[(1, "cap")]["passes"] == 5 and res[(1, "cap0")]["passes"] == 2


def test_deferred_scalar_pass_falls_back_on_irregular_bins(gx):
    """A bin whose product is exactly zero in one pass and non-zero in the next (here: a row holding only +1 and -1, which
    cancel under the all-ones start vector) is outside the deferred recurrence; the run must notice and give the
    literal-order result."""
    n = 2100
    A = onp.synth_map(n, seed=33)
    i, j, k = 700, 40, 1500
    A[i, :] = 0; A[:, i] = 0
    A[i, j] = A[j, i] = 1.0
    A[i, k] = A[k, i] = -1.0
    keep = np.ones(n, dtype=bool)
    with gx.DeviceMap.from_host(A) as dm:
        dm.set_mask(keep)
        a = dm.ic(1e-4, 40)
    with gx.DeviceMap(n, options={"ic_defer": 0}) as dm:
        dm.upload(A)
        dm.set_mask(keep)
        b = dm.ic(1e-4, 40)
    assert a["symmetric"] and b["symmetric"]
    assert a["passes"] == b["passes"] and a["products_launched"] == b["products_launched"]      # the rerun took the literal kernel
    assert np.array_equal(a["bias"], b["bias"])


LARGE_DIR = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden", "large")


def _large(name):
    import os
    p = os.path.join(LARGE_DIR, name)
    if not os.path.isfile(p):
        pytest.skip("fixture %s not generated yet (tests/golden/gen_golden_large.py, needs a GPU box)" % name)
    return np.load(p, allow_pickle=False)


@pytest.mark.parametrize("sym", [1, 0])
def test_at_size_ic_vs_fp64_oracle(gx, sym):
    """BASELINE configs[1] at full size (n = 24,926, the map bench.py times): bias and pass count of the CUDA path -- the
    upper-triangle deferred pass with its 296-CTA static + dynamic-pool schedule, and the dense TMA pass -- against the fp64
    CPU oracle (oracle_np.ic_fp64_matvec = lib/normalizeCore.pyx:41-60) run on the host copy of the same device-generated map."""
    g = _large("ic_n24926_s110.npz")
    n = int(g["n"])
    with gx.DeviceMap(n, options={"sym": sym}) as dm:
        dm.synth(seed=int(g["seed"]), scale=200.0, alpha=1.0, empty_frac=0.01)
        assert dm.checksum(0) == int(g["checksum"])                  # the same map, bit for bit
        dm.set_mask(np.ones(n, dtype=bool))
        r = dm.ic(float(g["tol"]), int(g["iteration"]))
    assert r["symmetric"] == bool(sym)
    assert r["passes"] == int(g["passes"])
    np.testing.assert_allclose(r["bias"], g["bias"], rtol=BIAS_RTOL)
    assert abs(r["l1"] - float(g["l1"][-1])) <= 1e-6 * float(g["l1"][-1])


@pytest.mark.parametrize("name,sym", [("kr_n24926_s110.npz", 1), ("kr_n24926_s110.npz", 0), ("kr_n49851_s110.npz", 1), ("kr_n49851_s110.npz", 0)])
def test_at_size_kr_vs_oracle(gx, name, sym):
    """KR at n = 24,926 and at BASELINE configs[2] (n = 49,851): x, outer iterations and matrix-vector products against
    oracle_np.kr (lib/normalizeKnightRuiz.pyx:106-191, 271-286) on the host copy of the same map."""
    g = _large(name)
    n = int(g["n"])
    with gx.DeviceMap(n, options={"sym": sym}) as dm:
        dm.synth(seed=int(g["seed"]), scale=200.0, alpha=1.0, empty_frac=0.01)
        assert dm.checksum(0) == int(g["checksum"])
        rs, _ = dm.row_stats()
        keep = rs != 0
        dm.set_mask(keep)
        r = dm.kr(float(g["tol"]), 0.1, 3.0)
    assert r["symmetric"] == bool(sym)
    assert (r["outer"], r["MVP"]) == (int(g["outer"]), int(g["MVP"]))
    np.testing.assert_allclose(r["x"][keep], g["x"][keep], rtol=BIAS_RTOL)
    assert np.all(r["x"][~keep] == 0)


# --------------------------------------------------------------------------------------------------
# coarsening
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names("DS"))
def test_downsample_golden_bit_exact(gx, name):
    """ccmap.downSampleCCMap vs the real reference's output: matrix bit-exact, bNoData, binsize, min/max."""
    from gcmapexplorer_b200 import ccmap as cmp
    g = load_golden(name)
    c = cmp.CCMAP()
    c.gen_from_matrix(g["input"], 1000, xlabel="chrS", ylabel="chrS")
    if g["bNoData_in"].size:
        c.bNoData = g["bNoData_in"]
    out = cmp.downSampleCCMap(c, **g["kwargs"])
    assert out.shape == g["matrix"].shape == cmp.getOutputShapeFor2DMapDownsampling(g["input"].shape, g["kwargs"]["level"])
    assert np.array_equal(np.array(out.matrix), g["matrix"])
    if g["bNoData_in"].size:
        assert np.array_equal(out.bNoData, g["bNoData"])
    assert out.binsize == int(g["binsize"])
    assert np.float32(out.minvalue) == np.float32(g["minvalue"]) and np.float32(out.maxvalue) == np.float32(g["maxvalue"])


@pytest.mark.parametrize("n,level,method", [(1000, 2, "sum"), (1001, 4, "mean"), (2049, 8, "sum"), (1500, 13, "mean"), (3001, 100, "sum"), (777, 5, "max")])
def test_downsample_vs_oracle(gx, n, level, method):
    from gcmapexplorer_b200 import ccmap as cmp
    A = onp.normalize_vc(onp.synth_map(n, seed=n))["matrix"]        # non-integer values: rounding order matters
    assert np.array_equal(cmp.downSample2DMap(A, level=level, method=method), onp.downsample_2d(A, level, method))
    out = np.zeros(onp.downsample_shape(A.shape, level), dtype=np.float32)
    assert cmp.downSample2DMap(A, outMatrix=out, level=level, method=method) is None
    assert np.array_equal(out, onp.downsample_2d(A, level, method))


# ---- transition probabilities + stationary distribution (lib/statDist.py) ----
SD_RTOL = 1e-9          # stationary vectors vs the reference's eigendecomposition


def _sd_rel(a, b):
    return float(np.max(np.abs(a - b) / b))


@pytest.mark.parametrize("name", golden_names("SD"))
def test_transition_matrix_golden(gx, name):
    from gcmapexplorer_b200 import ccmap as cmp, statDist as sdm
    g = load_golden(name)
    exact = name != "sd_s97ic"           # count data: float32 row sums are exact in any order -> bit-identical output
    P = sdm.calculateTransitionProbabilityMatrix(g["input"], bNoData=g["bNoData"])
    assert P.dtype == np.float32
    assert np.array_equal(P, g["matrix"]) if exact else rel_err(P, g["matrix"]) < MAT_RTOL
    out = np.zeros_like(g["input"])
    assert sdm.calculateTransitionProbabilityMatrix(g["input"], Out=out, bNoData=g["bNoData"]) is None
    assert np.array_equal(out, P)
    # the public entry point on a CCMAP, thresholds included
    c = cmp.CCMAP()
    c.gen_from_matrix(g["input"], 100000, xlabel="chrS", ylabel="chrS")
    c.make_unreadable()
    tp = sdm.transitionProbabilityMatrixForCCMap(c, **g["kwargs"])
    assert tp.matrix is None and c.matrix is not None                  # output unreadable, input readable (:129, :155)
    tp.make_readable()
    assert np.array_equal(tp.bNoData, g["bNoData"])
    assert np.array_equal(tp.matrix, g["matrix"]) if exact else rel_err(tp.matrix, g["matrix"]) < MAT_RTOL
    tol = 0 if exact else MAT_RTOL
    assert abs(tp.minvalue - g["minvalue"]) <= tol * g["minvalue"] and abs(tp.maxvalue - g["maxvalue"]) <= tol * g["maxvalue"]


@pytest.mark.parametrize("name", golden_names("SD"))
def test_stationary_distribution_golden(gx, name):
    from gcmapexplorer_b200 import ccmap as cmp, statDist as sdm
    g = load_golden(name)
    core = sdm.stationaryDistributionByEigenDecomp(g["matrix"], g["bNoData"])
    assert core.dtype == np.float64 and abs(core.sum() - 1.0) < 1e-12
    assert _sd_rel(core, g["sdist_core"]) < SD_RTOL
    _, iters = onp.stationary_power(g["matrix"], g["bNoData"])
    assert abs(sdm.last_run_info["iterations"] - iters) <= 2
    c = cmp.CCMAP()
    c.gen_from_matrix(g["matrix"], 100000, xlabel="chrS", ylabel="chrS")
    c.bNoData = g["bNoData"]
    c.make_unreadable()
    final = sdm.statDistrByEigenDecompForCCMap(c)
    assert c.matrix is None                                              # :318
    assert np.array_equal(final == 0, g["sdist"] == 0)
    nz = g["sdist"] != 0
    assert _sd_rel(final[nz], g["sdist"][nz]) < SD_RTOL


@pytest.mark.parametrize("n,seed", [(1031, 3), (2500, 4)])
def test_stationary_vs_power_oracle_and_zero_fill(gx, n, seed):
    """Larger maps against the float64 restatement of the same iteration; the second run feeds the raw count map,
    whose kept block contains zeros (lib/statDist.py:466-469 replaces them with the smallest non-zero entry)."""
    A = onp.synth_map(n, seed=seed, scale=40.0, alpha=1.2, empty_frac=0.02)
    b = ~onp.get_nonzeros_index(A)
    P, _ = onp.transition_matrix(A, b)
    for M in (P, A):
        exp, iters = onp.stationary_power(M, b)
        for variant in (1, 0):                   # TMA-bulk transposed product (default) and the per-thread-load one
            with gx.DeviceMap.from_host(M) as dm:
                dm.set_option("sd_variant", variant)
                dm.set_mask(~b)
                r = dm.stationary()
            got = r["x"]
            assert abs(r["iterations"] - iters) <= 2 and r["delta"] < 1e-11
            assert np.all(got[b] == 0) and abs(got.sum() - 1.0) < 1e-12
            ref = exp[~b] / exp[~b].sum()
            assert _sd_rel(got[~b], ref) < 1e-10
    if n <= 1100:
        assert _sd_rel(onp.stationary_eig(P, b), onp.stationary_power(P, b)[0]) < SD_RTOL


def test_transition_in_place_and_large_against_oracle(gx):
    A = onp.synth_map(2049, seed=9, scale=25.0, alpha=1.3, empty_frac=0.02)
    keep = onp.get_nonzeros_index(A, threshold_data_occup=0.02)          # drops 11 bins that do hold data
    exp, mv = onp.transition_matrix(A, ~keep)
    with gx.DeviceMap.from_host(A) as dm:
        dm.set_mask(keep)
        mn, mx = dm.transition(in_place=False)
        assert np.array_equal(dm.download(1), exp) and mn == float(mv) and mx == float(exp.max())
        assert np.array_equal(dm.download(0), A)                          # input untouched
        dm.transition(in_place=True)
        assert np.array_equal(dm.download(0), exp)


def test_transition_with_negative_entries_takes_the_exact_path(gx):
    """Signed integer-valued map (exact float32 row sums), one row with a negative sum: the per-entry division path of the
    minimum pass must give the reference's fill value bit for bit."""
    rng = np.random.default_rng(12)
    A = rng.integers(-3, 10, size=(300, 300)).astype(np.float32)
    A = np.triu(A) + np.triu(A, 1).T
    A[np.arange(300), np.arange(300)] = 50.0
    A[7, 7] = -4000.0
    b = np.zeros(300, dtype=bool)
    b[[3, 150]] = True
    exp, mv = onp.transition_matrix(A, b)
    assert mv < 0
    with gx.DeviceMap.from_host(A) as dm:
        dm.set_mask(~b)
        mn, mx = dm.transition(in_place=False)
        assert np.array_equal(dm.download(1), exp) and mn == float(mv)


def test_statdist_hdf5_handle_and_errors(gx):
    from gcmapexplorer_b200 import ccmap as cmp, statDist as sdm, util
    g = load_golden("sd_s400")

    class Handle:
        def __init__(self):
            self.calls, self.opened, self.closed = [], 0, 0

        def open(self):
            self.opened += 1

        def close(self):
            self.closed += 1

        def addDataByArray(self, chrom, resolution, key, arr, compression="lzf"):
            self.calls.append((chrom, resolution, key, np.array(arr), compression))

    c = cmp.CCMAP()
    c.gen_from_matrix(g["matrix"], 100000, xlabel="chrS", ylabel="chrS")
    c.bNoData = g["bNoData"]
    h = Handle()
    assert sdm.statDistrByEigenDecompForCCMap(c, hdf5Handle=h) is None and not h.calls       # no chrom: skipped (:305-307)
    assert sdm.statDistrByEigenDecompForCCMap(c, chrom="chrS", hdf5Handle=h, compression="gzip") is None
    assert h.opened == 1 and h.closed == 1
    res = [r for (_, r, k, _, _) in h.calls if k == "sum"]
    assert res == ["100kb", "200kb"]                                     # 400 bins -> 201 coarse bins (< 250: stop, :355-356)
    full = h.calls[0][3]
    assert np.max(np.abs(full[g["sdist"] != 0] - g["sdist"][g["sdist"] != 0]) / g["sdist"][g["sdist"] != 0]) < SD_RTOL
    by = {(r, k): a for (_, r, k, a, _) in h.calls}
    assert np.array_equal(by[("200kb", "maximum")], cmp.downSample1D(full, level=2, func="max"))
    assert all(cmpr == "gzip" for (*_, cmpr) in h.calls) and util.resolutionToBinsize("200kb") == 200000
    with pytest.raises(TypeError):
        sdm.calculateTransitionProbabilityMatrix(g["input"])             # bNoData=None fails in the reference too (:90)
    with pytest.raises(NotImplementedError):
        sdm.statDistrByEigenDecompForGCMap("in.gcmap", "out.h5", "100kb")


# ---- correlation matrix (lib/corrMatrix.py, lib/corrMatrixCoreSRC.c) ----
CORR_ATOL32 = 3e-7          # float32 maps of O(1) coefficients vs the reference's output (float64 agreement ~1e-15, then one rounding)
CORR_ATOL64 = 1e-10


@pytest.mark.parametrize("name", golden_names("CORR"))
def test_corr_matrix_golden(gx, name):
    from gcmapexplorer_b200 import ccmap as cmp, corrMatrix as cm
    g = load_golden(name)
    c = cmp.CCMAP()
    c.gen_from_matrix(g["input"], 100000, xlabel="chrS", ylabel="chrS")
    c.bNoData = g["bNoData"]
    c.make_unreadable()
    out = cm.calculateCorrMatrixForCCMap(c, **g["kwargs"])
    assert out.minvalue == -1 and out.maxvalue == 1 and np.array_equal(out.bNoData, g["bNoData"])
    M = np.array(out.matrix)
    assert M.dtype == np.float32 and np.abs(M.astype(np.float64) - g["matrix"]).max() < CORR_ATOL32
    assert np.array_equal(M, M.T) and np.all(np.diag(M) == 0) and np.all(M[g["bNoData"]] == 0) and np.all(M[:, g["bNoData"]] == 0)
    assert (M != g["matrix"]).mean() < 1e-3                               # all but a few rounding flips are bit-identical


@pytest.mark.parametrize("name", golden_names("CORRCORE"))
def test_corr_core_golden(gx, name):
    from gcmapexplorer_b200 import corrMatrix as cm
    g = load_golden(name)
    mv = g["kwargs"]["maskvalue"]
    corr = cm.calculateCorrMatrix(g["input"], maskvalue=mv)
    assert corr.dtype == np.float64 and np.abs(corr - g["corr"]).max() < CORR_ATOL64
    assert np.array_equal(corr == 0, g["corr"] == 0)                     # constant row, <= 3 common entries, diagonal: the same zeros
    out = np.full_like(g["input"], 7.0)
    assert cm.calculateCovMatrix(g["input"], out=out, maskvalue=mv) is None
    assert np.abs(out - g["cov"]).max() < CORR_ATOL64 * np.abs(g["cov"]).max()
    assert np.array_equal(np.diag(out), np.diag(g["cov"]))               # variances: the reference's own arithmetic, bit for bit


def test_corr_vs_compiled_reference_core(gx):
    """n = 1500 against the reference's C core compiled as it lies (oracle/_ref), raw counts and log space."""
    if not rl.have_corr():
        pytest.skip("oracle/_ref/libcorrMatrixCore.so not built")
    from gcmapexplorer_b200 import corrMatrix as cm
    A = onp.synth_map(1500, seed=15, scale=40.0, alpha=1.2, empty_frac=0.02)
    keep = onp.get_nonzeros_index(A)
    X = A[keep][:, keep].T.astype(np.float64, order="C")
    ref = rl.load_corr().calculateCorrMatrix(X, maskvalue=0.0)
    got = cm.calculateCorrMatrix(X, maskvalue=0.0)
    assert np.abs(got - ref).max() < CORR_ATOL64
    with np.errstate(divide="ignore"):
        L = np.log10(X)
    L[np.isinf(L)] = 0
    refl = rl.load_corr().calculateCorrMatrix(L, maskvalue=0.0)
    refl[np.isnan(refl)] = 0
    raw = {}
    # (variant, integer path): counts on the bf16 tensor cores + DSYRK + exact 8-bit-slice products for Sx (the default on this
    # integer-valued map) / the same with the float64 GEMM for Sx / three DGEMMs
    for variant, int_path in ((1, 1), (1, 0), (0, 0)):
        with gx.DeviceMap.from_host(A) as dm:
            dm.set_option("corr_variant", variant)
            dm.set_option("corr_int_path", int_path)
            dm.set_mask(keep)
            dm.corr(0.0, logspace=True)               # log space: never integer-valued -> float64 GEMM in every variant
            M = dm.download(1)
            assert np.abs(M[np.ix_(keep, keep)].astype(np.float64) - refl).max() < CORR_ATOL32 and np.all(M[~keep] == 0)
            assert np.array_equal(M, M.T)
            assert dm.corr_last_route() == (0, False)            # log space: float64 products
            raw[(variant, int_path)] = dm.corr_f64(X, 0.0)
            assert np.abs(raw[(variant, int_path)] - ref).max() < CORR_ATOL64
            slices, sxy = dm.corr_last_route()
            assert (slices > 0 and sxy) if (variant, int_path) == (1, 1) else (slices == 0 and not sxy)
    # on integer data every product is exact whichever unit computes it: bit-identical results
    assert np.array_equal(raw[(1, 1)], raw[(1, 0)]) and np.array_equal(raw[(1, 1)], raw[(0, 0)])
    # non-integer float32-valued data: Sx from exact int8 slice products (per-row exponents) vs the float64 GEMM
    Xf = (X.astype(np.float32) / np.float32(3.0)).astype(np.float64)
    reff = rl.load_corr().calculateCorrMatrix(Xf, maskvalue=0.0)
    got = {}
    for int_path in (1, 0):
        with gx.DeviceMap(32) as dm:
            dm.set_option("corr_int_path", int_path)
            got[int_path] = dm.corr_f64(Xf, 0.0)
            assert np.abs(got[int_path] - reff).max() < CORR_ATOL64
            slices, sxy = dm.corr_last_route()
            assert (slices < 0 and sxy) if int_path else slices == 0       # counts / 3 in float32: 5 slices cover every row
    assert np.abs(got[1] - got[0]).max() < 1e-12


def test_corr_errors(gx):
    from gcmapexplorer_b200 import ccmap as cmp, corrMatrix as cm
    A = onp.synth_map(60, seed=2)
    c = cmp.CCMAP()
    c.gen_from_matrix(A, 100000)
    assert cm.calculateCorrMatrixForCCMap(c) is None                    # no bNoData: warning + None, like the reference (:145-151)
    with pytest.raises(ValueError):
        cm.calculateCorrMatrix(np.zeros((4, 5)))
    with pytest.raises(NotImplementedError):
        cm.calculateCorrMatrixForGCMaps("in.gcmap", "out.gcmap")


# ---- size-independent properties of the widened rows on larger, device-generated maps ----
def test_statdist_fixed_point_property_large(gx):
    """n = 8000: the transition matrix is positive with every kept row summing to 1 up to the fill, and the stationary vector is a
    fixed point of x -> x P' (checked with a NumPy float64 product on the downloaded matrix)."""
    n = 8000
    with gx.DeviceMap(n) as dm:
        dm.synth(seed=77, scale=60.0, alpha=1.1, empty_frac=0.01)
        rs, _ = dm.row_stats()
        keep = rs != 0
        dm.set_mask(keep)
        mn, mx = dm.transition(in_place=True)
        P = dm.download(0)
        r = dm.stationary()
    assert mn > 0 and P.min() == np.float32(mn) and P.max() == np.float32(mx)
    assert np.all(P[~keep] == np.float32(mn)) and np.all(P[:, ~keep] == np.float32(mn))
    x = r["x"]
    assert np.all(x[~keep] == 0) and np.all(x[keep] > 0) and abs(x.sum() - 1.0) < 1e-12
    Pk = P[np.ix_(keep, keep)].astype(np.float64)
    y = x[keep] @ Pk
    y /= y.sum()
    assert np.abs(y - x[keep]).sum() < 1e-11 and r["iterations"] < 200


def test_corr_exact_int8_route_equals_float64_route_large(gx):
    """n = 5000: halving the map is exact in floating point and leaves every correlation unchanged, but turns the integer counts into
    non-integers -- so the same coefficients come once from the exact int8 slice products and once from the float64 GEMM + SYRK."""
    n = 5000
    out = []
    for factor in (1.0, 0.5):
        with gx.DeviceMap(n) as dm:
            dm.synth(seed=78, scale=80.0, alpha=1.1, empty_frac=0.01)
            if factor != 1.0:
                dm.scale(factor)
            rs, _ = dm.row_stats()
            keep = rs != 0
            dm.set_mask(keep)
            dm.corr(0.0)
            out.append(dm.download(1))
    a, b = out
    assert np.array_equal(a, a.T) and np.all(np.diag(a) == 0) and np.all(a[~keep] == 0)
    assert np.abs(a.astype(np.float64) - b).max() < CORR_ATOL32 and (a != b).mean() < 1e-3
    kept = a[np.ix_(keep, keep)]
    assert kept.max() > 0.5 and kept.min() < 0          # a real correlation structure, not zeros
