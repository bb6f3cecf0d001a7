"""Generate tests/golden/corr_*.npz from the REAL reference -- SURVEY.md 8f row N4 (lib/corrMatrix.py + lib/corrMatrixCoreSRC.c).

Run in the build container only (needs /root/reference and `make -C oracle ref`):
    python tests/golden/gen_golden_corr.py

The reference's C core is compiled as it lies (oracle/Makefile); its Cython wrapper does not cythonize against NumPy 2.x, so
oracle/ref_loader.py registers a ctypes stand-in for it and lib/corrMatrix.py runs unmodified on top.  Each fixture stores the
float32 input map, its bNoData, the keyword arguments and the matrix `calculateCorrMatrixForCCMap` returned; the `core_*`
fixtures store float64 inputs and what the C entry points returned for `calculateCorrMatrix` / `calculateCovMatrix`.
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
cm = rl.load_corr_module()
core = rl.load_corr()
cmp = sys.modules["gcMapExplorer.lib.ccmap"]


def case(name, A, **kw):
    c = cmp.CCMAP()
    c.gen_from_matrix(A, 100000, xlabel="chrS", ylabel="chrS")
    b = ~onp.get_nonzeros_index(A)
    c.bNoData = b
    out = cm.calculateCorrMatrixForCCMap(c, **kw)
    out.make_readable()
    M = np.array(out.matrix)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), input=A, method="CORR", kwargs=json.dumps(kw), matrix=M, bNoData=b,
                        minvalue=np.float64(out.minvalue), maxvalue=np.float64(out.maxvalue))
    exp = onp.corr_for_map(A, b, **kw)
    print(name, A.shape, "range", float(M.min()), float(M.max()), "oracle max abs diff", float(np.abs(exp.astype(np.float64) - M).max()))


def core_case(name, X, maskvalue):
    corr = core.calculateCorrMatrix(X, maskvalue=maskvalue)
    cov = core.calculateCovMatrix(X, maskvalue=maskvalue)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), input=X, method="CORRCORE",
                        kwargs=json.dumps(dict(maskvalue=maskvalue)), corr=corr, cov=cov)
    print(name, X.shape, "oracle corr diff", float(np.abs(onp.corr_matrix(X, maskvalue) - corr).max()),
          "cov diff rel", float(np.abs(onp.corr_matrix(X, maskvalue, covariance=True) - cov).max() / np.abs(cov).max()))


A97 = onp.synth_map(97, seed=97)
A131 = onp.synth_map(131, seed=131, scale=30.0, alpha=1.3, empty_frac=0.03)
A400 = onp.synth_map(400, seed=400, scale=60.0, alpha=1.1, empty_frac=0.02)
ic97 = onp.normalize_ic(A97)["matrix"]
case("corr_s97", A97)
case("corr_s97_log", A97, logspace=True)
case("corr_s97ic_log", ic97, logspace=True)
case("corr_sp131", A131)
case("corr_sp131_clamp", A131, vmin=1.0, vmax=25.0)
case("corr_sp131_mask1", A131, maskvalue=1.0)
case("corr_s400", A400)
case("corr_s97_nomask", A97, maskvalue=None)            # no pairwise masking at all: plain Pearson of the kept block
rng = np.random.default_rng(5)
X = rng.standard_normal((60, 60))
X[rng.random((60, 60)) < 0.3] = 0.0
X[7] = 2.5                                   # constant row: variance exactly 0 -> correlations 0 (:121-124)
X[9, 5:] = 0.0                               # 5 entries only; pairs with <= 3 common entries give 0 (:171)
core_case("corr_core_masked", X, 0.0)
core_case("corr_core_nomask", X, None)
