"""DeviceMap: one resident contact map (or one rank's row block of it) on a B200, driven through the
C ABI of libgcxnorm.so.  This is the only module that touches ctypes pointers; the reference-shaped
modules (normalizeCore, normalizeKnightRuiz, ccmapHelpers, cmstats, normalizer) are written on top."""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import GcxError, check, lib


def default_device():
    for key in ("GCX_DEVICE", "LOCAL_RANK"):
        if os.environ.get(key):
            return int(os.environ[key])
    return 0


def _as_f32_matrix(matrix):
    m = matrix
    if not isinstance(m, np.ndarray) or m.dtype != np.float32 or m.ndim != 2 or m.strides[1] != 4 or m.strides[0] % 4 != 0:
        m = np.ascontiguousarray(matrix, dtype=np.float32)
    return m


def device_pitch(n):
    """Row pitch (in floats) of the resident map: n rounded up to 32 floats = 128 bytes."""
    return (int(n) + 31) // 32 * 32


class PinnedArray:
    """float32 2-D array in CUDA pinned host memory (full-speed PCIe copies).  With pitched=True the rows use
    the device's own pitch, so uploads/downloads are one contiguous DMA transfer (and overlap with kernels);
    `.array` is the [rows, cols] view either way."""

    def __init__(self, shape, dtype=np.float32, pitched=False):
        self.shape = tuple(shape)
        rows, cols = self.shape
        ld = device_pitch(cols) if pitched else cols
        nbytes = int(rows) * int(ld) * np.dtype(dtype).itemsize
        self._ptr = C.c_void_p()
        check(lib().gcx_host_alloc(nbytes, C.byref(self._ptr)))
        buf = (C.c_char * nbytes).from_address(self._ptr.value)
        self.full = np.frombuffer(buf, dtype=dtype).reshape(rows, ld)
        self.full[:, cols:] = 0
        self.array = self.full[:, :cols]

    def free(self):
        if self._ptr is not None and self._ptr.value:
            self.array = None
            self.full = None
            lib().gcx_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceMap:
    def __init__(self, n, row0=0, nloc=None, device=None, dist=None, options=None):
        """dist: None or (rank, world, id128 bytes) for row-block sharding over NCCL.
        options: dict for gcx_set_option, applied before the map is created."""
        self._ctx = C.c_void_p()
        self.n = int(n)
        self.row0 = int(row0)
        self.nloc = int(self.n - self.row0 if nloc is None else nloc)
        self.device = default_device() if device is None else int(device)
        check(lib().gcx_ctx_create(self.device, C.byref(self._ctx)))
        try:
            if dist is not None:
                rank, world, id128 = dist
                buf = C.create_string_buffer(bytes(id128), 128)
                check(lib().gcx_dist_init(self._ctx, int(rank), int(world), buf))
            for key, val in (options or {}).items():
                self.set_option(key, val)
            check(lib().gcx_map_create(self._ctx, self.n, self.row0, self.nloc))
        except Exception:
            self.close()
            raise

    def set_option(self, key, value):
        check(lib().gcx_set_option(self._ctx, str(key).encode(), float(value)))

    # ---- lifetime -------------------------------------------------------------------------------
    def close(self):
        if self._ctx is not None and self._ctx.value:
            lib().gcx_ctx_destroy(self._ctx)
        self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- residency ------------------------------------------------------------------------------
    @classmethod
    def from_host(cls, matrix, device=None):
        m = _as_f32_matrix(matrix)
        if m.shape[0] != m.shape[1]:
            raise ValueError("contact map must be square")
        dm = cls(m.shape[0], device=device)
        dm.upload(m)
        return dm

    def upload(self, rows):
        """rows: float32 [nloc, n] (the local row block; the whole map on one GPU)."""
        m = _as_f32_matrix(rows)
        if m.shape != (self.nloc, self.n):
            raise ValueError(f"expected shape {(self.nloc, self.n)}, got {m.shape}")
        check(lib().gcx_map_upload(self._ctx, C.c_void_p(m.ctypes.data), m.strides[0] // 4))

    # ---- device-resident twin: torch tensors (or any CUDA pointer) on this device ---------------------------------
    def adopt_tensor(self, t):
        """Make a float32 CUDA tensor [nloc, n] of this device the resident map: one device-to-device copy, no host trip."""
        if not t.is_cuda or t.device.index != self.device or str(t.dtype) != "torch.float32" or tuple(t.shape) != (self.nloc, self.n) or t.stride(1) != 1:
            raise ValueError("expected a float32 CUDA tensor [%d, %d] on device %d with unit column stride" % (self.nloc, self.n, self.device))
        import torch
        torch.cuda.current_stream(t.device).synchronize()
        check(lib().gcx_map_adopt_device(self._ctx, C.c_void_p(t.data_ptr()), int(t.stride(0))))

    def export_tensor(self, which=1, out=None):
        """A result map as a float32 CUDA tensor of this device (device-to-device copy)."""
        import torch
        if out is None:
            out = torch.empty((self.nloc, self.n), dtype=torch.float32, device=torch.device("cuda", self.device))
        if not out.is_cuda or out.device.index != self.device or out.dtype != torch.float32 or tuple(out.shape) != (self.nloc, self.n) or out.stride(1) != 1:
            raise ValueError("out must be a float32 CUDA tensor [nloc, n] on this device")
        torch.cuda.current_stream(out.device).synchronize()
        check(lib().gcx_map_export_device(self._ctx, int(which), C.c_void_p(out.data_ptr()), int(out.stride(0))))
        return out

    def upload_rows(self, rows, row_begin, with_stats=True):
        """A chunk of the local row block (streaming sources: gzip / HDF5 readers); finish with upload_finish()."""
        m = _as_f32_matrix(rows)
        if m.ndim != 2 or m.shape[1] != self.n:
            raise ValueError(f"expected [rows, {self.n}] float32, got {m.shape}")
        check(lib().gcx_map_upload_rows(self._ctx, C.c_void_p(m.ctypes.data), m.strides[0] // 4, int(row_begin), m.shape[0], int(bool(with_stats))))

    def upload_finish(self):
        check(lib().gcx_map_upload_finish(self._ctx))

    def upload_file(self, path, offset=0, with_stats=True):
        """The local row block straight from a raw float32 file (.npbin): parallel pread -> pinned stages -> HBM."""
        check(lib().gcx_map_upload_file(self._ctx, os.fsencode(path), int(offset), int(bool(with_stats))))

    def download_file(self, which, path, offset=0):
        """A result map straight into a raw float32 file (created / sized as needed)."""
        check(lib().gcx_map_download_file(self._ctx, int(which), os.fsencode(path), int(offset)))

    def download(self, which=1, out=None):
        if out is None:
            out = np.empty((self.nloc, self.n), dtype=np.float32)
        if out.dtype != np.float32 or out.shape != (self.nloc, self.n) or out.strides[1] != 4:
            raise ValueError("out must be float32 [nloc, n] with unit column stride")
        check(lib().gcx_map_download(self._ctx, int(which), C.c_void_p(out.ctypes.data), out.strides[0] // 4))
        return out

    def download_async(self, which, out):
        """Start a device->host copy into a pinned array on the copy stream; pair with download_wait()."""
        if out.dtype != np.float32 or out.shape != (self.nloc, self.n) or out.strides[1] != 4:
            raise ValueError("out must be float32 [nloc, n] with unit column stride")
        check(lib().gcx_map_download_async(self._ctx, int(which), C.c_void_p(out.ctypes.data), out.strides[0] // 4))

    def download_wait(self):
        check(lib().gcx_download_wait(self._ctx))

    def synth(self, seed, scale=200.0, alpha=1.0, empty_frac=0.01):
        check(lib().gcx_map_synth(self._ctx, int(seed), float(scale), float(alpha), float(empty_frac)))

    def clamp(self, vmin=None, vmax=None):
        check(lib().gcx_map_clamp(self._ctx, vmin is not None, 0.0 if vmin is None else float(vmin),
                                  vmax is not None, 0.0 if vmax is None else float(vmax)))

    def scale(self, factor):
        check(lib().gcx_map_scale(self._ctx, float(factor)))

    def swap(self):
        check(lib().gcx_map_swap(self._ctx))

    def checksum(self, which=0):
        s = C.c_uint64(0)
        check(lib().gcx_map_checksum(self._ctx, int(which), C.byref(s)))
        return s.value

    def sync(self):
        check(lib().gcx_sync(self._ctx))

    def info(self):
        sm, tot, free = C.c_int(0), C.c_int64(0), C.c_int64(0)
        check(lib().gcx_device_info(self._ctx, C.byref(sm), C.byref(tot), C.byref(free)))
        return dict(sm_count=sm.value, hbm_total=tot.value, hbm_free=free.value)

    # ---- mask -----------------------------------------------------------------------------------
    def invalidate(self):
        """Drop the cached row statistics (benchmark steps that must re-read the map)."""
        check(lib().gcx_map_invalidate(self._ctx))

    def row_stats(self, force=False):
        """(rowsum float64[n], raw zero_count int32[n]).  Cached on the device until the map changes;
        force=True re-reads the map (benchmark steps)."""
        if force:
            check(lib().gcx_map_invalidate(self._ctx))
        rs = np.empty(self.n, dtype=np.float64)
        zc = np.empty(self.n, dtype=np.int32)
        check(lib().gcx_row_stats(self._ctx, rs.ctypes.data_as(_lib._c_f64p), zc.ctypes.data_as(_lib._c_i32p)))
        return rs, zc

    def row_negative_flags(self):
        """int32[n], 1 where the row holds negative entries (their float32 emptiness test is re-checked on the host)."""
        f = np.empty(self.n, dtype=np.int32)
        check(lib().gcx_row_negative_flags(self._ctx, f.ctypes.data_as(_lib._c_i32p)))
        return f

    def set_mask(self, keep):
        k = np.ascontiguousarray(np.asarray(keep, dtype=bool).astype(np.uint8))
        if k.shape != (self.n,):
            raise ValueError("mask must have n entries")
        check(lib().gcx_set_mask(self._ctx, k.ctypes.data_as(_lib._c_u8p)))

    # ---- solvers --------------------------------------------------------------------------------
    def ic(self, tol=1e-4, iteration=500):
        bias = np.empty(self.n, dtype=np.float64)
        passes, mv, l1 = C.c_int(0), C.c_int(0), C.c_double(0)
        check(lib().gcx_ic_run(self._ctx, float(tol), int(iteration), bias.ctypes.data_as(_lib._c_f64p),
                               C.byref(passes), C.byref(l1), C.byref(mv)))
        pb, sym = self.pass_bytes()
        launched, peer = self.last_run_products()
        return dict(bias=bias, passes=passes.value, l1=l1.value, matvecs=mv.value, products_launched=launched, peer_exchange=peer,
                    ms=self.last_run_ms(), pass_bytes=pb, symmetric=sym)

    def kr(self, tol=1e-12, delta=0.1, Delta=3.0):
        x = np.empty(self.n, dtype=np.float64)
        outer, mvp, rout = C.c_int(0), C.c_int(0), C.c_double(0)
        check(lib().gcx_kr_run(self._ctx, float(tol), float(delta), float(Delta), x.ctypes.data_as(_lib._c_f64p),
                               C.byref(outer), C.byref(mvp), C.byref(rout)))
        pb, sym = self.pass_bytes()
        return dict(x=x, outer=outer.value, MVP=mvp.value, rout=rout.value, ms=self.last_run_ms(), pass_bytes=pb, symmetric=sym)

    def transition(self, in_place=False):
        """Transition probability matrix of the resident map under the resident mask (lib/statDist.py:50-99);
        result in the output map (download(1)) or in place.  Returns (min non-zero, max)."""
        mn, mx = C.c_float(0), C.c_float(0)
        check(lib().gcx_transition_run(self._ctx, int(bool(in_place)), C.byref(mn), C.byref(mx)))
        return mn.value, mx.value

    def stationary(self, tol=1e-13, max_iter=100000):
        """Dominant left eigenvector of the kept block of the resident (transition) matrix by power iteration
        (replaces the dense eigendecomposition of lib/statDist.py:480-484).  x sums to 1 over the kept bins."""
        x = np.empty(self.n, dtype=np.float64)
        iters, delta = C.c_int(0), C.c_double(0)
        check(lib().gcx_stationary_run(self._ctx, float(tol), int(max_iter), x.ctypes.data_as(_lib._c_f64p),
                                       C.byref(iters), C.byref(delta)))
        return dict(x=x, iterations=iters.value, delta=delta.value, ms=self.last_run_ms(), pass_bytes=self.pass_bytes()[0])

    def corr(self, maskvalue=0.0, logspace=False, vmin=None, vmax=None):
        """Pairwise-masked Pearson correlation of the kept block's columns (lib/corrMatrix.py:92-135 +
        lib/corrMatrixCoreSRC.c) into the output map (download(1)); no-data rows/columns and the diagonal are 0."""
        check(lib().gcx_corr_run(self._ctx, int(maskvalue is not None), float(0.0 if maskvalue is None else maskvalue), int(bool(logspace)),
                                 int(vmin is not None), float(np.float32(0.0 if vmin is None else vmin)),
                                 int(vmax is not None), float(np.float32(0.0 if vmax is None else vmax))))

    def corr_last_route(self):
        """(slices, sxy_sliced) of the last correlation run: slices > 0 integer map, < 0 float32-valued map, 0 float64 products."""
        a, b = C.c_int(0), C.c_int(0)
        check(lib().gcx_corr_last_route(self._ctx, C.byref(a), C.byref(b)))
        return a.value, bool(b.value)

    def i8gemm(self, L=None, R=None, mb=None, symmetric=False, reps=1, want_result=True):
        """T = L R^T on the library's own int8 tensor-core kernel (test / measurement hook).  Returns (T int32 or None, ms per launch)."""
        if L is not None:
            L = np.ascontiguousarray(L, dtype=np.int8)
            mb = L.shape[0]
        if R is not None:
            R = np.ascontiguousarray(R, dtype=np.int8)
        T = np.empty((mb, mb), dtype=np.int32) if want_result else None
        ms = C.c_double(0)
        check(lib().gcx_debug_i8gemm(self._ctx, None if L is None else C.c_void_p(L.ctypes.data), None if R is None else C.c_void_p(R.ctypes.data), int(mb),
                                     None if T is None else C.c_void_p(T.ctypes.data), int(bool(symmetric)), int(reps), C.byref(ms)))
        return T, ms.value

    def corr_f64(self, in_array, maskvalue=None, covariance=False, out=None):
        """Correlation (or covariance) matrix of the rows of a square float64 array, float64 result
        (lib/_corrMatrixCore.pyx:49-146).  Needs no resident map."""
        a = np.ascontiguousarray(in_array, dtype=np.float64)
        if a.ndim != 2 or a.shape[0] != a.shape[1]:
            raise ValueError("in_array must be square")
        res = out if (out is not None and out.dtype == np.float64 and out.flags.c_contiguous and out.shape == a.shape) else np.empty_like(a)
        check(lib().gcx_corr_matrix_f64(self._ctx, a.ctypes.data_as(_lib._c_f64p), a.shape[0], int(maskvalue is not None),
                                        float(0.0 if maskvalue is None else maskvalue), int(bool(covariance)), res.ctypes.data_as(_lib._c_f64p)))
        if out is not None and res is not out:
            out[:] = res
        return res

    # ---- elementwise passes ---------------------------------------------------------------------
    def apply_sym(self, s=None, masked_mode=0, in_place=False):
        mn, mx = C.c_float(0), C.c_float(0)
        sp = None
        if s is not None:
            s = np.ascontiguousarray(s, dtype=np.float64)
            if s.shape != (self.n,):
                raise ValueError("scale vector must have n entries")
            sp = s.ctypes.data_as(_lib._c_f64p)
        check(lib().gcx_apply_sym_scale(self._ctx, sp, int(masked_mode), int(bool(in_place)), C.byref(mn), C.byref(mx)))
        return mn.value, mx.value

    def apply_norm_vectors(self, c1, c2, in_place=False):
        """count / (c1[row] * c2[col]) with +inf on a zero product (.hic import with KR / VC vectors, lib/hic2gcmap.py:82-98).
        Returns (min non-zero, max)."""
        c1 = np.ascontiguousarray(c1, dtype=np.float64)
        c2 = np.ascontiguousarray(c2, dtype=np.float64)
        if c1.shape != (self.n,) or c2.shape != (self.n,):
            raise ValueError("norm vectors must have n entries")
        mn, mx = C.c_float(0), C.c_float(0)
        check(lib().gcx_apply_norm_vectors(self._ctx, c1.ctypes.data_as(_lib._c_f64p), c2.ctypes.data_as(_lib._c_f64p), int(bool(in_place)),
                                           C.byref(mn), C.byref(mx)))
        return mn.value, mx.value

    def apply_sym_vc(self, keep_vc, masked_mode=0, sqroot=False):
        """Fused tail: scaled map (resident scale + mask) -> output 1, vanilla-coverage map under `keep_vc` -> output 2, one read
        of the input.  Returns ((min, max) of the scaled map, (min, max) of the VC map)."""
        k = np.ascontiguousarray(np.asarray(keep_vc, dtype=bool).astype(np.uint8))
        if k.shape != (self.n,):
            raise ValueError("mask must have n entries")
        a, b, c, d = C.c_float(0), C.c_float(0), C.c_float(0), C.c_float(0)
        check(lib().gcx_apply_sym_vc(self._ctx, k.ctypes.data_as(_lib._c_u8p), int(masked_mode), int(bool(sqroot)),
                                     C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return (a.value, b.value), (c.value, d.value)

    def vc(self, sqroot=False, in_place=False):
        mn, mx = C.c_float(0), C.c_float(0)
        csum = np.empty(self.n, dtype=np.float32)
        check(lib().gcx_vc_run(self._ctx, int(bool(sqroot)), int(bool(in_place)), C.byref(mn), C.byref(mx),
                               csum.ctypes.data_as(_lib._c_f32p)))
        return mn.value, mx.value, csum

    def diag_stats(self, stats="median"):
        avg = np.empty(self.n, dtype=np.float64)
        check(lib().gcx_diag_stats(self._ctx, 1 if stats == "median" else 0, avg.ctypes.data_as(_lib._c_f64p)))
        return avg

    def diag_avg(self):
        """The per-diagonal statistic left resident by the last diag_stats / MapBatch.mcfs."""
        avg = np.empty(self.n, dtype=np.float64)
        check(lib().gcx_diag_avg_read(self._ctx, avg.ctypes.data_as(_lib._c_f64p)))
        return avg

    def diag_apply(self, avg, subtract=False, in_place=False):
        mn, mx = C.c_float(0), C.c_float(0)
        ap = None
        if avg is not None:
            avg = np.ascontiguousarray(avg, dtype=np.float64)
            ap = avg.ctypes.data_as(_lib._c_f64p)
        check(lib().gcx_diag_apply(self._ctx, ap, int(bool(subtract)), int(bool(in_place)), C.byref(mn), C.byref(mx)))
        return mn.value, mx.value

    def compact_f64(self, keep, out=None):
        """float64 [nk, nk] copy of the kept rows/columns (remove_zeros, lib/ccmapHelpers.pyx:286-299)."""
        kept = np.ascontiguousarray(np.nonzero(np.asarray(keep, dtype=bool))[0].astype(np.int32))
        nk = int(kept.size)
        if out is None:
            out = np.empty((nk, nk), dtype=np.float64)
        if out.shape != (nk, nk) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 [nk, nk] array")
        check(lib().gcx_compact_f64(self._ctx, kept.ctypes.data_as(_lib._c_i32p), nk, out.ctypes.data_as(_lib._c_f64p)))
        return out

    def downsample(self, level=2, method="sum"):
        """(coarse float32 matrix, min non-zero, max): downSample2DMap of lib/ccmap.py:737-842, bit-identical."""
        code = {"sum": 0, "mean": 1, "max": 2}.get(method, 0)        # the reference treats unknown methods as 'sum'
        on = (self.n - 1 + int(level) - 1) // int(level) + 1
        out = np.empty((on, on), dtype=np.float32)
        mn, mx = C.c_float(0), C.c_float(0)
        check(lib().gcx_downsample(self._ctx, int(level), code, out.ctypes.data_as(_lib._c_f32p), on, C.byref(mn), C.byref(mx)))
        return out, mn.value, mx.value

    # ---- measurement ----------------------------------------------------------------------------
    def profile(self, enable=True):
        check(lib().gcx_profile(self._ctx, int(bool(enable))))

    def profile_read(self, reset=True):
        la = (C.c_int64 * _lib.NCLASS)()
        ms = (C.c_double * _lib.NCLASS)()
        check(lib().gcx_profile_read(self._ctx, int(bool(reset)), la, ms))
        return {name: dict(launches=int(la[i]), ms=float(ms[i])) for i, name in enumerate(_lib.KERNEL_CLASSES)}

    def timer_start(self):
        check(lib().gcx_timer_start(self._ctx))

    def timer_stop(self):
        ms = C.c_double(0)
        check(lib().gcx_timer_stop(self._ctx, C.byref(ms)))
        return ms.value

    def pass_bytes(self):
        """(HBM bytes per matrix pass of the last ic/kr run, used_upper_triangle)."""
        b, sym = C.c_int64(0), C.c_int(0)
        check(lib().gcx_pass_bytes(self._ctx, C.byref(b), C.byref(sym)))
        return b.value, bool(sym.value)

    def last_run_products(self):
        """(products the last ic/kr run launched, whether the sharded IC pass used the NVLink peer exchange)."""
        a, b = C.c_int(0), C.c_int(0)
        check(lib().gcx_last_run_products(self._ctx, C.byref(a), C.byref(b)))
        return a.value, bool(b.value)

    def last_run_ms(self):
        ms = C.c_double(0)
        check(lib().gcx_last_run_ms(self._ctx, C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_int64(0)
        check(lib().gcx_launch_count(self._ctx, C.byref(n)))
        return n.value

    def bench_matvec(self, iters=20, variant=-1):
        ms = C.c_double(0)
        check(lib().gcx_bench_matvec(self._ctx, int(variant), int(iters), C.byref(ms)))
        return ms.value / iters


class MapBatch:
    """Several resident maps on one device treated as one job (the chromosomes of a genome): each stage of the
    distance-frequency normalization is one grid over all of them (gcx_mcfs_batch)."""

    def __init__(self, maps):
        self.maps = list(maps)
        if not self.maps:
            raise ValueError("empty batch")
        self._arr = (C.c_void_p * len(self.maps))(*[m._ctx for m in self.maps])

    def diag_stats_pooled(self, stats="median"):
        """Per-diagonal statistic of the POOLED non-zero entries of all maps (lib/cmstats.py:572-578); length = largest map."""
        nmax = max(m.n for m in self.maps)
        avg = np.empty(nmax, dtype=np.float64)
        check(lib().gcx_diag_stats_pooled(self._arr, len(self.maps), 1 if stats == "median" else 0, avg.ctypes.data_as(_lib._c_f64p), nmax))
        return avg

    def mcfs(self, stats="median", subtract=False):
        """diag_stats + diag_apply (into output 1) for every map; returns [(min non-zero, max)] per map."""
        k = len(self.maps)
        mn = (C.c_float * k)()
        mx = (C.c_float * k)()
        check(lib().gcx_mcfs_batch(self._arr, k, 1 if stats == "median" else 0, int(bool(subtract)), mn, mx))
        return [(mn[i], mx[i]) for i in range(k)]


def dist_unique_id():
    buf = C.create_string_buffer(128)
    check(lib().gcx_dist_unique_id(buf))
    return buf.raw
