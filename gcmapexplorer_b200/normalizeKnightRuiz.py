"""Mirror of ``gcMapExplorer/lib/normalizeKnightRuiz.pyx`` (class KnightRuizNorm).

The Newton / conjugate-gradient iteration of normalizeKnightRuiz.pyx:106-191, 271-286 runs device-resident
(kernels k_matvec + k_kr_step) with the reference's exact control flow, so ``.i`` (outer iterations) and
``.MVP`` (matrix-vector products) match.  ``memory='HDD'`` is accepted and runs the same GPU path -- the
reference's disk mode exists only because the map did not fit in RAM; here that role is played by
row-sharding over GPUs.
"""
import numpy as np

from .device import DeviceMap

dtype_npBINarray = "float32"


class KnightRuizNorm:
    def __init__(self, A, memory="RAM", tol=1e-12, delta=0.1, Delta=3, workDir=None):
        if memory not in ("RAM", "HDD"):
            raise ValueError("This memory={0} is not is not understandable.. Please use 'RAM' or 'HDD'.".format(memory))
        self.memory = memory
        self.workDir = workDir
        self.tol, self.delta, self.Delta = tol, delta, Delta
        self.n = A.shape[0]
        self.MVP = 0
        self.i = 0
        self.x = None
        self.rout = None
        self._dm = DeviceMap.from_host(np.asarray(A) if not isinstance(A, np.ndarray) else A)
        self._dm.set_mask(np.ones(self.n, dtype=bool))

    def run(self, A, fl, OutMatrix, bNoData):
        """Balance, then write diag(x) A diag(x) into OutMatrix where ~bNoData (normalizeKnightRuiz.pyx:193-269).
        Returns (minvalue, maxvalue) of the balanced block."""
        dm = self._dm
        r = dm.kr(self.tol, self.delta, self.Delta)
        self.x = r["x"].reshape(self.n, 1)
        self.i, self.MVP, self.rout = r["outer"], r["MVP"], r["rout"]
        minvalue, maxvalue = dm.apply_sym(None, masked_mode=0, in_place=False)
        D = dm.download(1)
        kept = np.nonzero(~np.asarray(bNoData, dtype=bool))[0]
        if kept.size != self.n:
            raise ValueError("bNoData does not match the size of A")
        OutMatrix[np.ix_(kept, kept)] = D
        if hasattr(OutMatrix, "flush"):
            OutMatrix.flush()
        return minvalue, maxvalue

    def close(self):
        if self._dm is not None:
            self._dm.close()
            self._dm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
