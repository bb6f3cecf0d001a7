"""CPU: the NumPy oracle (oracle/oracle_np.py) against golden vectors produced by the REAL reference
(tests/golden/gen_golden.py) and, where oracle/_ref is built, against the compiled reference cores live."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, rel_err
from oracle import oracle_np as onp
from oracle import ref_loader as rl

# as-is IC is fp32-in-place; the fp64 restatement sits inside its noise floor
IC_ASIS_VS_FP64_RTOL = 5e-5


@pytest.mark.parametrize("name", golden_names("MASK"))
def test_mask_golden(name):
    g = load_golden(name)
    keep = onp.get_nonzeros_index(g["input"], g["kwargs"].get("percentile_threshold_no_data"), g["kwargs"].get("threshold_data_occup"))
    assert np.array_equal(keep, g["keep"])


@pytest.mark.parametrize("name", golden_names("IC"))
def test_ic_golden(name):
    g = load_golden(name)
    kw = g["kwargs"]
    asis = onp.normalize_ic(g["input"], fp64=False, **kw)
    # the as-is restatement calls the same NumPy primitives in the same order -> bit-exact
    assert np.array_equal(asis["bNoData"], g["bNoData"])
    assert np.array_equal(asis["matrix"], g["matrix"])
    assert asis["passes"] == int(g["passes"])
    assert np.float32(asis["minvalue"]) == np.float32(g["minvalue"])
    assert np.float32(asis["maxvalue"]) == np.float32(g["maxvalue"])
    # the fp64 restatement agrees with the reference to the reference's own fp32 noise
    tight = onp.normalize_ic(g["input"], fp64=True, **kw)
    assert np.array_equal(tight["bNoData"], g["bNoData"])
    assert rel_err(tight["matrix"], g["matrix"]) < IC_ASIS_VS_FP64_RTOL
    if kw.get("iteration", 500) >= 100:      # converged runs: pass counts agree to within the fp32 stall
        assert abs(tight["passes"] - int(g["passes"])) <= 6
    # invariant of the recurrence (lib/normalizeCore.pyx:36,53): total is conserved on the kept block
    keep = tight["keep"]
    sub_in = onp.clamp(g["input"], kw.get("vmin"), kw.get("vmax"))[keep][:, keep].astype(np.float64)
    M64 = sub_in / np.outer(tight["bias"], tight["bias"])
    assert abs(M64.sum() - sub_in.sum()) < 1e-9 * sub_in.sum()


@pytest.mark.parametrize("name", golden_names("KR"))
def test_kr_golden(name):
    g = load_golden(name)
    r = onp.normalize_kr(g["input"], **g["kwargs"])
    assert np.array_equal(r["bNoData"], g["bNoData"])
    assert r["outer"] == int(g["outer"]) and r["MVP"] == int(g["MVP"])
    np.testing.assert_allclose(r["x"], g["x"], rtol=1e-12)
    assert rel_err(r["matrix"], g["matrix"]) < 2e-7          # fp32 storage, <= 1 ulp
    assert abs(r["minvalue"] - float(g["minvalue"])) <= 2e-7 * abs(float(g["minvalue"]))
    assert abs(r["maxvalue"] - float(g["maxvalue"])) <= 2e-7 * abs(float(g["maxvalue"]))
    kept = ~g["bNoData"]
    rs = r["matrix"][kept][:, kept].astype(np.float64).sum(axis=1)
    tol = g["kwargs"].get("tol", 1e-12)
    assert np.max(np.abs(rs - 1.0)) < max(1e-5, 10 * np.sqrt(tol))


@pytest.mark.parametrize("name", golden_names("VC"))
def test_vc_golden(name):
    g = load_golden(name)
    r = onp.normalize_vc(g["input"], **g["kwargs"])
    assert np.array_equal(r["bNoData"], g["bNoData"])
    assert np.array_equal(r["matrix"], g["matrix"])              # same fp32 primitives -> bit-exact
    assert np.float32(r["minvalue"]) == np.float32(g["minvalue"])
    assert np.float32(r["maxvalue"]) == np.float32(g["maxvalue"])


@pytest.mark.parametrize("name", golden_names("MCFS"))
def test_mcfs_golden(name):
    g = load_golden(name)
    kw = dict(g["kwargs"])
    if kw.get("scaleUpInput"):
        nzv = g["input"][g["input"] != 0]
        kw["minvalue_attr"] = float(nzv.min())                   # CCMAP.gen_from_matrix sets minvalue so (lib/ccmap.py:219-221)
    r = onp.normalize_mcfs(g["input"], **kw)
    assert np.array_equal(r["bNoData"], g["bNoData"])
    if "avg" in g:
        if kw.get("stats", "median") == "median":
            assert np.array_equal(r["avg"], g["avg"])
        else:
            np.testing.assert_allclose(r["avg"], g["avg"], rtol=1e-14)
    assert rel_err(r["matrix"], g["matrix"]) < 2e-7
    assert abs(r["minvalue"] - float(g["minvalue"])) <= 2e-7 * abs(float(g["minvalue"]))


def test_mcfs_bad_stype_returns_none():
    assert onp.normalize_mcfs(np.eye(4, dtype=np.float32), stype="bogus") is None


def test_mask_errors():
    A = onp.synth_map(40, seed=3)
    with pytest.raises(AssertionError):
        onp.get_nonzeros_index(A, 90, 0.5)
    with pytest.raises(ValueError):
        onp.get_nonzeros_index(A, None, 1.5)


@pytest.mark.skipif(not rl.have_cores(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n,seed", [(64, 1), (257, 2), (600, 3)])
def test_oracle_vs_live_reference_cores(n, seed):
    """Fresh random maps through the compiled reference cores (oracle/_ref) vs the restatement."""
    core, krmod, cmh = rl.load_cores()
    A = onp.synth_map(n, seed=seed)
    for kw in (dict(), dict(threshold_percentile=95), dict(threshold_data_occup=0.6)):
        assert np.array_equal(cmh.get_nonzeros_index(A, **kw),
                              onp.get_nonzeros_index(A, kw.get("threshold_percentile"), kw.get("threshold_data_occup")))
    keep = onp.get_nonzeros_index(A)
    sub = A[keep][:, keep]
    # IC as-is: bit-exact incl. pass count
    M = np.array(sub, copy=True)          # order="K": keep the layout fancy-indexing produced, as the reference does
    rl.ic_pass_counter(reset=True)
    core.performIterativeCorrection(M, 1e-4, 500)
    Mo, Bo, passes = onp.ic_asis(sub, 1e-4, 500)
    assert np.array_equal(M, Mo) and passes == rl.ic_pass_counter()
    # IC fp64 restatement inside the as-is noise floor
    M64, B64, p64, _ = onp.ic_fp64(sub, 1e-4, 500)
    assert rel_err(M64.astype(np.float32), M) < IC_ASIS_VS_FP64_RTOL
    # KR: counters equal, x to 1e-12
    K = krmod.KnightRuizNorm(sub, tol=1e-12, delta=0.1, Delta=3)
    class F(np.ndarray):
        def flush(self):
            pass
    K.run(sub, 0, np.zeros_like(A).view(F), ~keep)
    r = onp.kr(sub)
    assert (r["outer"], r["MVP"]) == (K.i, K.MVP)
    np.testing.assert_allclose(r["x"], np.asarray(K.x)[:, 0], rtol=1e-12)
    # VC: bit-exact
    out = A.copy()
    core.performVCNormalization(A, Out=out, sqroot=False, bNoData=~keep)
    assert np.array_equal(out, onp.vc(A, bNoData=~keep, out=A.copy())[0])
    # MCFS apply: <= 1 ulp
    avg = onp.avg_contact_by_distance(A, "mean")
    o1 = np.zeros_like(A); core.normalizeByAvgContactByDivision(A, avg, Out=o1)
    o2 = np.zeros_like(A); core.normalizeByAvgContactBySubtraction(A, avg, Out=o2)
    assert np.array_equal(o1, onp.mcfs_divide(A, avg))
    assert np.array_equal(o2, onp.mcfs_subtract(A, avg))


@pytest.mark.parametrize("name", golden_names("DS"))
def test_downsample_golden(name):
    g = load_golden(name)
    out = onp.downsample_2d(g["input"], g["kwargs"]["level"], g["kwargs"]["method"])
    assert np.array_equal(out, g["matrix"])
    if g["bNoData_in"].size:
        assert np.array_equal(onp.downsample_nodata(g["bNoData_in"], out.shape[0], g["kwargs"]["level"]), g["bNoData"])


# ---- transition probabilities + stationary distribution (lib/statDist.py) ----
@pytest.mark.parametrize("name", golden_names("SD"))
def test_statdist_golden(name):
    g = load_golden(name)
    b = g["bNoData"]
    P, mv = onp.transition_matrix(g["input"], b)
    assert np.array_equal(P, g["matrix"])                                  # float32 divisions: bit-exact
    assert float(mv) == g["minvalue"] and float(P.max()) == g["maxvalue"]
    core = onp.stationary_eig(g["matrix"], b)
    assert np.max(np.abs(core - g["sdist_core"]) / g["sdist_core"]) < 1e-10       # LAPACK build differences only
    power, iters = onp.stationary_power(g["matrix"], b)
    assert np.max(np.abs(power - g["sdist_core"]) / g["sdist_core"]) < 1e-9 and iters < 1000
    final = onp.stationary_postprocess(g["sdist_core"])
    np.testing.assert_allclose(final, g["sdist"], rtol=1e-13, atol=0)
    assert np.array_equal(final == 0, g["sdist"] == 0)


def test_stationary_replaces_zeros_of_the_kept_block():
    """lib/statDist.py:466-469: a raw count map (zeros inside the kept block) still gives a positive vector."""
    A = onp.synth_map(131, seed=131, scale=30.0, alpha=1.3, empty_frac=0.03)
    b = ~onp.get_nonzeros_index(A)
    assert (A[~b][:, ~b] == 0).any()
    e = onp.stationary_eig(A, b)
    p, _ = onp.stationary_power(A, b)
    assert (e > 0).all() and abs(e.sum() - 1) < 1e-12 and np.max(np.abs(p - e) / e) < 1e-9


# ---- correlation matrix (lib/corrMatrix.py, lib/corrMatrixCoreSRC.c) ----
@pytest.mark.parametrize("name", golden_names("CORR"))
def test_corr_golden(name):
    g = load_golden(name)
    exp = onp.corr_for_map(g["input"], g["bNoData"], **g["kwargs"])
    assert exp.dtype == np.float32 and np.abs(exp.astype(np.float64) - g["matrix"]).max() < 2e-7       # float32 rounding of ~1e-16 differences
    assert g["minvalue"] == -1 and g["maxvalue"] == 1
    assert np.all(g["matrix"][g["bNoData"]] == 0) and np.all(np.diag(g["matrix"]) == 0)


@pytest.mark.parametrize("name", golden_names("CORRCORE"))
def test_corr_core_golden(name):
    g = load_golden(name)
    mv = g["kwargs"]["maskvalue"]
    assert np.abs(onp.corr_matrix(g["input"], mv) - g["corr"]).max() < 1e-12
    cov = onp.corr_matrix(g["input"], mv, covariance=True)
    assert np.abs(cov - g["cov"]).max() < 1e-12 * np.abs(g["cov"]).max()
    assert np.all(g["corr"][7] == 0) and np.all(np.diag(g["corr"]) == 0)       # constant row / untouched diagonal
