"""Generate tests/golden/*.npz from the REAL reference (rjdkmr/gcMapExplorer v1.0.30).

Run in the build container only (needs /root/reference and `make -C oracle ref`):
    python tests/golden/gen_golden.py

Each fixture stores the float32 input map, the call's keyword arguments and what the reference's own
public API (lib/normalizer.py: normalizeCCMapByIC/KR/VCNorm/MCFS) returned: output matrix, bNoData,
minvalue, maxvalue; plus, for KR, the x / outer-iteration / MVP counters read from a
KnightRuizNorm object driven exactly like lib/normalizer.py:335-336 does, and for IC the number of
passes the compiled core ran (counted through the np.sqrt call at lib/normalizeCore.pyx:52).
The reference has no golden vectors of its own; these files are the pin.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader as rl          # noqa: E402
from oracle import oracle_np as onp          # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
nz = rl.load_full()
core, krmod, cmh = rl.load_cores()
cmp = sys.modules["gcMapExplorer.lib.ccmap"]
cmstats = sys.modules["gcMapExplorer.lib.cmstats"]

KAT5 = np.array([[10, 4, 0, 2, 1], [4, 8, 0, 3, 0], [0, 0, 0, 0, 0], [2, 3, 0, 12, 5], [1, 0, 0, 5, 9]], dtype=np.float32)


class _Flushable(np.ndarray):
    """ndarray standing in for the np.memmap OutMatrix (GenNormMatrixRAM calls OutMatrix.flush(), :257)."""
    def flush(self):
        pass


def make_ccmap(A, binsize=100000):
    c = cmp.CCMAP()
    c.gen_from_matrix(A, binsize, xlabel="chrS", ylabel="chrS")
    c.make_unreadable()
    return c


def result(cc):
    cc.make_readable()
    r = dict(matrix=np.array(cc.matrix), bNoData=np.asarray(cc.bNoData, dtype=bool),
             minvalue=np.float64(cc.minvalue), maxvalue=np.float64(cc.maxvalue))
    cc.make_unreadable()
    return r


def run_case(name, A, method, **kw):
    c = make_ccmap(A)
    extra = {}
    if method == "IC":
        rl.ic_pass_counter(reset=True)
        out = nz.normalizeCCMapByIC(c, **kw)
        extra["passes"] = np.int64(rl.ic_pass_counter())
    elif method == "KR":
        out = nz.normalizeCCMapByKR(c, **kw)
        # counters: drive the class the way lib/normalizer.py:327-336 does
        T = onp.clamp(A, kw.get("vmin"), kw.get("vmax"))
        keep = cmh.get_nonzeros_index(T, threshold_percentile=kw.get("percentile_threshold_no_data"),
                                      threshold_data_occup=kw.get("threshold_data_occup"))
        sub = T[keep, :][:, keep]
        K = krmod.KnightRuizNorm(sub, tol=kw.get("tol", 1e-12), delta=0.1, Delta=3)
        dump = np.zeros_like(T).view(_Flushable)
        K.run(sub, 0, dump, ~keep)
        extra.update(x=np.asarray(K.x)[:, 0].copy(), outer=np.int64(K.i), MVP=np.int64(K.MVP),
                     rout=np.float64(np.asarray(K.rout).item()))
    elif method == "VC":
        out = nz.normalizeCCMapByVCNorm(c, **kw)
    elif method == "MCFS":
        out = nz.normalizeCCMapByMCFS(c, **kw)
        c.make_readable()
        T = onp.clamp(A, kw.get("vmin"), kw.get("vmax"))
        if not kw.get("scaleUpInput"):
            tc = make_ccmap(T)
            tc.make_readable()
            extra["avg"] = np.asarray(cmstats.getAvgContactByDistance(tc, stats=kw.get("stats", "median")), dtype=np.float64)
            tc.make_unreadable()
        c.make_unreadable()
    elif method == "MASK":
        keep = cmh.get_nonzeros_index(A, threshold_percentile=kw.get("percentile_threshold_no_data"),
                                      threshold_data_occup=kw.get("threshold_data_occup"))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), input=A, method=method, kwargs=json.dumps(kw), keep=np.asarray(keep, dtype=bool))
        print(f"{name:34s} MASK kept={int(keep.sum())}/{A.shape[0]}")
        return
    r = result(out)
    r.update(extra)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), input=A, method=method, kwargs=json.dumps(kw), **r)
    msg = " ".join(f"{k}={int(v)}" for k, v in extra.items() if np.ndim(v) == 0 and k != "rout")
    print(f"{name:34s} {method:4s} n={A.shape[0]:4d} min={r['minvalue']:.6g} max={r['maxvalue']:.6g} nodata={int(r['bNoData'].sum())} {msg}")


def sparse_map(n, seed):
    """Far-field-sparse map (steeper decay, low scale) so the percentile / occupancy filters bite."""
    return onp.synth_map(n, seed=seed, scale=30.0, alpha=1.3, empty_frac=0.03)


if __name__ == "__main__":
    maps = {
        "kat5": KAT5,
        "s97": onp.synth_map(97, seed=97),
        "sp131": sparse_map(131, seed=131),
        "c1_482": onp.synth_map(482, seed=21100, scale=2000.0),      # BASELINE config 1 shape (chr21 @100 kb)
    }
    for mname, A in maps.items():
        big = mname == "c1_482"
        run_case(f"{mname}_ic", A, "IC", tol=1e-4)
        run_case(f"{mname}_kr", A, "KR", tol=1e-12)
        run_case(f"{mname}_vc", A, "VC")
        run_case(f"{mname}_mcfs_median_oe", A, "MCFS")
        if big:
            continue
        run_case(f"{mname}_vc_sqrt", A, "VC", sqroot=True)
        run_case(f"{mname}_mcfs_mean_oe", A, "MCFS", stats="mean")
        run_case(f"{mname}_mcfs_mean_ome", A, "MCFS", stats="mean", stype="o-e")
        run_case(f"{mname}_mcfs_median_ome", A, "MCFS", stats="median", stype="o-e")
    # filters, clamps, iteration cap, scale-up
    A = maps["sp131"]
    run_case("sp131_mask_p90", A, "MASK", percentile_threshold_no_data=90)
    run_case("sp131_mask_occ", A, "MASK", threshold_data_occup=0.33)
    run_case("sp131_ic_p90", A, "IC", tol=1e-4, percentile_threshold_no_data=90)
    run_case("sp131_ic_occ", A, "IC", tol=1e-4, threshold_data_occup=0.33)
    run_case("sp131_kr_p90", A, "KR", percentile_threshold_no_data=90)
    run_case("sp131_vc_p90", A, "VC", percentile_threshold_no_data=90)
    run_case("sp131_mcfs_p90", A, "MCFS", percentile_threshold_no_data=90)
    A = maps["s97"]
    run_case("s97_ic_clamp", A, "IC", tol=1e-4, vmin=1.0, vmax=150.0)
    run_case("s97_kr_clamp", A, "KR", vmin=1.0, vmax=150.0)
    run_case("s97_vc_clamp", A, "VC", vmin=1.0, vmax=150.0)
    run_case("s97_mcfs_clamp", A, "MCFS", vmin=1.0, vmax=150.0)
    run_case("s97_ic_cap3", A, "IC", tol=1e-12, iteration=3)
    run_case("s97_kr_tol6", A, "KR", tol=1e-6)
    # MCFS on an already-normalized (non-integer) map incl. scaleUpInput
    r = onp.normalize_vc(A)
    run_case("s97vc_mcfs_scaleup", r["matrix"], "MCFS", scaleUpInput=True)
    run_case("s97vc_mcfs_median", r["matrix"], "MCFS")
