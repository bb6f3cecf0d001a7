"""ShardedDeviceMap: ONE map row-sharded over several B200s of this node from ONE process -- what the drop-in normalizers use
when a map does not fit a single GPU (or when GCX_GPUS asks for it).  The reference's answer to "too large for RAM" is
memory='HDD' (lib/normalizer.py:338-351, 479-482); here it is row sharding (BASELINE north_star).

One libgcxnorm context per GPU, each driven by its own host thread (ctypes releases the GIL; the library is re-entrant per
context).  The contexts are wired with NCCL (unique id passed inside the process) for the small collectives and -- for the
iterative-correction pass -- exchange their partial vectors by direct NVLink stores into each other's buffers (same-process
peer access).  The partition is the library's aligned one (gcx_partition_rows: near-equal row blocks cut on 1024 rows), so a
symmetric map runs the upper-triangle passes with the off-diagonal blocks dealt out in a circulant pattern; symmetry is TESTED
across the ranks (k_sym_hash), not assumed, and an asymmetric map takes the dense passes.

The class mirrors the DeviceMap methods the normalizers call; every collective method runs on all ranks at once and returns
rank 0's (identical) result.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib
from .device import DeviceMap, dist_unique_id
from .dist import row_partition_aligned


def visible_gpus():
    return _lib.device_count()


class ShardedDeviceMap:
    def __init__(self, n, devices=None, options=None):
        self.n = int(n)
        if devices is None:
            devices = list(range(visible_gpus()))
        self.devices = [int(d) for d in devices]
        self.world = len(self.devices)
        if self.world < 2:
            raise ValueError("ShardedDeviceMap needs at least two GPUs")
        self.parts = [row_partition_aligned(self.n, self.world, r) for r in range(self.world)]
        if any(nl <= 0 for _, nl in self.parts):
            raise ValueError("more GPUs than 1024-row blocks: use fewer devices for a map of %d bins" % self.n)
        self._pool = ThreadPoolExecutor(max_workers=self.world)
        uid = dist_unique_id()
        opts = {"exchange_allreduce": 1}
        opts.update(options or {})

        def make(r):
            row0, nloc = self.parts[r]
            return DeviceMap(self.n, row0, nloc, device=self.devices[r], dist=(r, self.world, uid), options=opts)
        self.maps = [None] * self.world
        try:
            self.maps = list(self._pool.map(make, range(self.world)))
        except Exception:
            self.close()
            raise

    # ---- plumbing -------------------------------------------------------------------------------
    def _all(self, fn):
        """fn(rank, DeviceMap) on every rank concurrently; list of results in rank order."""
        return list(self._pool.map(lambda r: fn(r, self.maps[r]), range(self.world)))

    def close(self):
        maps, self.maps = getattr(self, "maps", []), []
        for m in maps:
            if m is not None:
                m.close()
        pool = getattr(self, "_pool", None)
        if pool is not None:
            pool.shutdown(wait=True)
            self._pool = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- residency ------------------------------------------------------------------------------
    @classmethod
    def from_host(cls, matrix, devices=None, options=None):
        if matrix.shape[0] != matrix.shape[1]:
            raise ValueError("contact map must be square")
        sm = cls(matrix.shape[0], devices=devices, options=options)
        sm.upload(matrix)
        return sm

    def upload(self, matrix):
        """Every rank streams its own row block from the host array / memmap (row blocks of a C-contiguous map are contiguous)."""
        if tuple(matrix.shape) != (self.n, self.n):
            raise ValueError(f"expected shape {(self.n, self.n)}, got {tuple(matrix.shape)}")
        self._all(lambda r, m: m.upload(matrix[self.parts[r][0]:self.parts[r][0] + self.parts[r][1]]))

    def upload_file(self, path, offset=0, with_stats=True):
        """Every rank reads its own row block of the raw float32 file (parallel pread per rank)."""
        self._all(lambda r, m: m.upload_file(path, offset, with_stats))

    def upload_rows(self, rows, row_begin, with_stats=True):
        """A chunk of GLOBAL rows [row_begin, row_begin + len(rows)): every rank takes the part inside its block."""
        def part(r, m):
            r0, nl = self.parts[r]
            a, b = max(row_begin, r0), min(row_begin + rows.shape[0], r0 + nl)
            if a < b:
                m.upload_rows(rows[a - row_begin:b - row_begin], a - r0, with_stats)
        self._all(part)

    def upload_finish(self):
        self._all(lambda r, m: m.upload_finish())

    def download_file(self, which, path, offset=0):
        """Every rank writes its rows into the shared raw float32 file (sized first)."""
        with open(path, "wb") as f:
            f.truncate(4 * self.n * self.n + int(offset))
        self._all(lambda r, m: m.download_file(which, path, offset))

    def download(self, which=1, out=None):
        if out is None:
            out = np.empty((self.n, self.n), dtype=np.float32)
        if tuple(out.shape) != (self.n, self.n) or out.dtype != np.float32:
            raise ValueError("out must be float32 [n, n]")
        self._all(lambda r, m: m.download(which, out=out[self.parts[r][0]:self.parts[r][0] + self.parts[r][1]]))
        return out

    def clamp(self, vmin=None, vmax=None):
        self._all(lambda r, m: m.clamp(vmin, vmax))

    def scale(self, factor):
        self._all(lambda r, m: m.scale(factor))

    def checksum(self, which=0):
        return sum(self._all(lambda r, m: m.checksum(which))) & 0xffffffffffffffff

    def sync(self):
        self._all(lambda r, m: m.sync())

    # ---- collective passes: identical results on every rank, rank 0's returned -------------------------
    def row_stats(self, force=False):
        return self._all(lambda r, m: m.row_stats(force))[0]

    def set_mask(self, keep):
        self._all(lambda r, m: m.set_mask(keep))

    def ic(self, tol=1e-4, iteration=500):
        res = self._all(lambda r, m: m.ic(tol, iteration))
        out = dict(res[0])
        out["pass_bytes"] = sum(x["pass_bytes"] for x in res)
        out["ranks_identical"] = all(np.array_equal(x["bias"], res[0]["bias"]) for x in res[1:])
        return out

    def kr(self, tol=1e-12, delta=0.1, Delta=3.0):
        res = self._all(lambda r, m: m.kr(tol, delta, Delta))
        out = dict(res[0])
        out["pass_bytes"] = sum(x["pass_bytes"] for x in res)
        out["ranks_identical"] = all(np.array_equal(x["x"], res[0]["x"]) for x in res[1:])
        return out

    def apply_sym(self, s=None, masked_mode=0, in_place=False):
        return self._all(lambda r, m: m.apply_sym(s, masked_mode, in_place))[0]      # min / max are allreduced in the library

    def vc(self, sqroot=False, in_place=False):
        return self._all(lambda r, m: m.vc(sqroot, in_place))[0]

    def diag_stats(self, stats="median"):
        return self._all(lambda r, m: m.diag_stats(stats))[0]

    def diag_apply(self, avg, subtract=False, in_place=False):
        return self._all(lambda r, m: m.diag_apply(avg, subtract, in_place))[0]

    def info(self):
        return self._all(lambda r, m: m.info())


def gpus_for_map(n, requested=None):
    """How many GPUs the drop-in normalizers shard a map of n bins over: GCX_GPUS (or `requested`) when given, else 1 if the map
    and its result fit the first GPU's free HBM (2 x 4 n^2 bytes + vectors), else every visible GPU."""
    env = os.environ.get("GCX_GPUS") if requested is None else requested
    have = visible_gpus()
    if env:
        k = max(1, min(int(env), have))
        return k
    if have < 2:
        return 1
    with DeviceMap(8) as probe:
        free = probe.info()["hbm_free"]
    need = 8.0 * n * (n + 32) + (1 << 30)
    return 1 if need < 0.9 * free else have
