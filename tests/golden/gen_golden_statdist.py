"""Generate tests/golden/sd_*.npz from the REAL reference (rjdkmr/gcMapExplorer, lib/statDist.py) -- SURVEY.md 8f row N3.

Run in the build container only (needs /root/reference and `make -C oracle ref`):
    python tests/golden/gen_golden_statdist.py

Each fixture stores the float32 input map, the keyword arguments, and what the reference's public API returned:
`transitionProbabilityMatrixForCCMap` (matrix, bNoData, minvalue, maxvalue), `stationaryDistributionByEigenDecomp`
on that matrix (sdist_core) and `statDistrByEigenDecompForCCMap` on the transition CCMAP (sdist: outliers removed,
renormalised).
"""
import importlib
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader as rl          # noqa: E402
from oracle import oracle_np as onp          # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
nz = rl.load_full()
cmp = sys.modules["gcMapExplorer.lib.ccmap"]
sd = importlib.import_module("gcMapExplorer.lib.statDist")

KAT5 = np.array([[10, 4, 0, 2, 1], [4, 8, 0, 3, 0], [0, 0, 0, 0, 0], [2, 3, 0, 12, 5], [1, 0, 0, 5, 9]], dtype=np.float32)


def make_ccmap(A, binsize=100000):
    c = cmp.CCMAP()
    c.gen_from_matrix(A, binsize, xlabel="chrS", ylabel="chrS")
    c.make_unreadable()
    return c


def case(name, A, **kw):
    c = make_ccmap(A)
    tp = sd.transitionProbabilityMatrixForCCMap(c, **kw)
    tp.make_readable()
    P = np.array(tp.matrix)
    b = np.asarray(tp.bNoData, dtype=bool)
    tp.make_unreadable()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")               # ComplexWarning when LAPACK returns a complex eigenvector array
        core = sd.stationaryDistributionByEigenDecomp(P, b)
        final = sd.statDistrByEigenDecompForCCMap(tp)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), input=A, method="SD", kwargs=json.dumps(kw), matrix=P, bNoData=b,
                        minvalue=np.float64(tp.minvalue), maxvalue=np.float64(tp.maxvalue), sdist_core=np.asarray(core, dtype=np.float64),
                        sdist=np.asarray(final, dtype=np.float64))
    Po, _ = onp.transition_matrix(A, b)
    print(name, A.shape, "masked", int(b.sum()), "oracle TP bit-exact:", np.array_equal(Po, P), "oracle eig rel:",
          float(np.max(np.abs(onp.stationary_eig(P, b) - core) / core)), "power rel:",
          float(np.max(np.abs(onp.stationary_power(P, b)[0] - core) / core)),
          "final equal:", np.array_equal(onp.stationary_postprocess(core), final))


A97 = onp.synth_map(97, seed=97)
A131 = onp.synth_map(131, seed=131, scale=30.0, alpha=1.3, empty_frac=0.03)
A400 = onp.synth_map(400, seed=400, scale=60.0, alpha=1.1, empty_frac=0.02)
ic97 = onp.normalize_ic(A97)["matrix"]                 # non-integer input: fp32 row sums depend on the summation order
case("sd_kat5", KAT5)
case("sd_s97", A97)
case("sd_s97ic", ic97)
case("sd_sp131", A131)
case("sd_sp131_p95", A131, percentile_threshold_no_data=95)
case("sd_sp131_occ", A131, threshold_data_occup=0.33)
case("sd_s400", A400)
