#!/usr/bin/env python
"""At-size oracle vectors for the maps bench.py times (run ONCE on a GPU box; the fixtures are committed).

    python tests/golden/gen_golden_large.py [--n 24926] [--kr-n 49851] [--seed 110]

The benchmark maps are generated ON THE DEVICE (k_synth: counter-based Poisson draws with the GPU's fast-math intrinsics),
so the host cannot reproduce them bit for bit -- the map is generated on the GPU, downloaded, and the CPU oracle
(oracle/oracle_np.py: ic_fp64_matvec = lib/normalizeCore.pyx:41-60 in fp64, kr = lib/normalizeKnightRuiz.pyx:106-191,
271-286) runs on the host copy.  Only O(n) results are stored (a few hundred KB):

    tests/golden/large/ic_n<N>_s<seed>.npz   bias[n], passes, l1 history, map checksum, row sums / zero counts digest
    tests/golden/large/kr_n<N>_s<seed>.npz   x[n], outer, MVP, rout, keep digest, map checksum

tests/test_gpu_parity.py::test_at_size_* regenerate the same map on the device (equal checksum), run the CUDA path -- the
upper-triangle passes and the dense passes -- and compare: bias / x to 1e-9, equal pass / outer / MVP counts.
"""
import argparse
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gcmapexplorer_b200 import device as gx      # noqa: E402
from oracle import oracle_np as onp              # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "large")
SYNTH = dict(scale=200.0, alpha=1.0, empty_frac=0.01)       # bench.py's map law


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def host_map(n, seed):
    with gx.DeviceMap(n) as dm:
        dm.synth(seed=seed, **SYNTH)
        cs = dm.checksum(0)
        rs, zc = dm.row_stats()
        A = dm.download(0)
    return A, cs, rs, zc


class F32Dot:
    """Duck type for oracle_np.kr: A_kept.dot(x) in float64 on the float32 host map, without compacting it and without a float64
    copy (n = 49,851 would need 20 GB): row blocks are converted and multiplied by a pool of host threads (NumPy releases the GIL)."""
    def __init__(self, A32, kept, block=1024, threads=None):
        from concurrent.futures import ThreadPoolExecutor
        self.A, self.kept, self.block = A32, kept, block
        self.shape = (kept.size, kept.size)
        self.pool = ThreadPoolExecutor(max_workers=threads or min(32, os.cpu_count() or 1))

    def dot(self, xk):
        n = self.A.shape[0]
        x = np.zeros(n)
        x[self.kept] = np.asarray(xk).ravel()
        out = np.empty(n)

        def work(r0):
            out[r0:r0 + self.block] = self.A[r0:r0 + self.block].astype(np.float64).dot(x)
        list(self.pool.map(work, range(0, n, self.block)))
        return out[self.kept].reshape(np.shape(xk))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=24926)
    ap.add_argument("--kr-n", type=int, default=49851)
    ap.add_argument("--seed", type=int, default=110)
    ap.add_argument("--skip-kr-large", action="store_true")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)

    # ---- IC (unfiltered: the whole map, lib/normalizer.py:592-594) and KR (kept block) at n ----------------
    t0 = time.time()
    A, cs, rs, zc = host_map(args.n, args.seed)
    assert np.array_equal(A, A.T), "the synthetic map must be symmetric"
    print("map n=%d generated + downloaded in %.1f s, checksum %016x" % (args.n, time.time() - t0, cs), flush=True)
    t0 = time.time()
    A64 = A.astype(np.float64)
    B, passes, l1 = onp.ic_fp64_matvec(A, 1e-4, 500, matvec=lambda w: A64.T.dot(w))
    print("IC oracle: %d passes, last L1 %.3e, %.1f s" % (passes, l1[-1], time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(OUT, "ic_n%d_s%d.npz" % (args.n, args.seed)), n=args.n, seed=args.seed, checksum=np.uint64(cs),
                        bias=B, passes=passes, l1=np.array(l1), rowsum_sha=digest(rs), zero_count_sha=digest(zc), tol=1e-4, iteration=500)
    keep = rs != 0.0            # lib/ccmapHelpers.pyx:192, no percentile / occupancy filter
    t0 = time.time()
    kept = np.nonzero(keep)[0]
    Ak = A64[np.ix_(kept, kept)]
    r = onp.kr(Ak, 1e-12, 0.1, 3)
    x = np.zeros(args.n)
    x[kept] = r["x"]
    print("KR oracle: outer %d, MVP %d, rout %.3e, %.1f s" % (r["outer"], r["MVP"], r["rout"], time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(OUT, "kr_n%d_s%d.npz" % (args.n, args.seed)), n=args.n, seed=args.seed, checksum=np.uint64(cs), x=x,
                        outer=r["outer"], MVP=r["MVP"], rout=r["rout"], keep_sha=digest(keep), tol=1e-12)
    del A64, Ak, A

    # ---- KR at BASELINE configs[2] (n = 49,851) -------------------------------------------------------------
    if not args.skip_kr_large and args.kr_n:
        t0 = time.time()
        A, cs, rs, zc = host_map(args.kr_n, args.seed)
        keep = rs != 0.0
        kept = np.nonzero(keep)[0]
        print("map n=%d generated + downloaded in %.1f s" % (args.kr_n, time.time() - t0), flush=True)
        t0 = time.time()
        r = onp.kr(F32Dot(A, kept), 1e-12, 0.1, 3)
        x = np.zeros(args.kr_n)
        x[kept] = r["x"]
        print("KR oracle n=%d: outer %d, MVP %d, rout %.3e, %.1f s" % (args.kr_n, r["outer"], r["MVP"], r["rout"], time.time() - t0), flush=True)
        np.savez_compressed(os.path.join(OUT, "kr_n%d_s%d.npz" % (args.kr_n, args.seed)), n=args.kr_n, seed=args.seed, checksum=np.uint64(cs), x=x,
                            outer=r["outer"], MVP=r["MVP"], rout=r["rout"], keep_sha=digest(keep), tol=1e-12)


if __name__ == "__main__":
    main()
