"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference (rjdkmr/gcMapExplorer v1.0.30).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this.  Nothing here is on the product path.

Two tiers:

* ``load_cores()``  -- works wherever ``oracle/_ref/`` exists (this container AND the GPU box): loads
  the three Cython hot-path modules compiled by ``make -C oracle ref`` from the reference's own
  ``lib/normalizeCore.pyx``, ``lib/normalizeKnightRuiz.pyx`` and ``lib/ccmapHelpers.pyx``
  (reference setup.py:40-42).  These are the reference's numeric cores, unmodified.
* ``load_full()``   -- only where ``/root/reference`` exists (this container): additionally imports the
  reference's pure-Python orchestration (``lib/normalizer.py``, ``lib/ccmap.py``, ``lib/cmstats.py``,
  ``lib/util.py``) straight from ``/root/reference`` so golden vectors can be generated from the
  reference's own public API (tests/golden/gen_golden.py).

No reference source is copied.  Four non-invasive shims make the 2017 code importable on
Python 3.12 / NumPy 2.3:
  1. ``appdirs`` stub               (gcMapExplorer/config.py:26-27 imports it; not installed)
  2. ``h5py`` stub                  (lib/ccmap.py:31, lib/gcmap.py:24, lib/cmstats.py:28; not installed)
  3. empty ``gcMapExplorer`` / ``gcMapExplorer.lib`` package objects with ``__path__`` set, bypassing
     ``gcMapExplorer/__init__.py`` and ``lib/__init__.py`` which import every sibling subsystem
  4. a NumPy proxy bound as the module-global ``np`` of the two compiled cores that restores
     ``np.float`` (lib/normalizeCore.pyx:37) and lets ``np.amin/np.amax`` take the ragged
     ``[1x1 array, scalar]`` lists of lib/normalizeKnightRuiz.pyx:185,187 (NumPy >= 1.24 rejects them)
"""
import importlib
import os
import sys
import tempfile
import types

import numpy as _np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO_DIR = os.path.join(_HERE, "_ref", "gcMapExplorer", "lib")
REF_SRC = os.environ.get("GCX_REFERENCE_SRC", "/root/reference")


class _NumpyProxy:
    """Shim 4: numpy with np.float restored and scalar-tolerant amin/amax."""

    float = float
    sqrt_calls = 0          # lib/normalizeCore.pyx:52 calls np.sqrt once per IC pass -> pass counter

    def __getattr__(self, name):
        return getattr(_np, name)

    def sqrt(self, *args, **kw):
        _NumpyProxy.sqrt_calls += 1
        return _np.sqrt(*args, **kw)

    @staticmethod
    def _flat(a):
        if isinstance(a, (list, tuple)):
            return [x.item() if isinstance(x, _np.ndarray) and x.size == 1 else x for x in a]
        return a

    def amin(self, a, *args, **kw):
        return _np.amin(self._flat(a), *args, **kw)

    def amax(self, a, *args, **kw):
        return _np.amax(self._flat(a), *args, **kw)


def _stub_module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _kth_diag_indices(k, a):
    """Own restatement of lib/util.py:75-98 for boxes without /root/reference (diagonal indices)."""
    rows, cols = _np.diag_indices_from(a)
    if k < 0:
        return rows[:k], cols[-k:]
    if k > 0:
        return rows[k:], cols[:-k]
    return rows, cols


def have_cores():
    return os.path.isdir(REF_SO_DIR) and any(f.startswith("normalizeCore") for f in os.listdir(REF_SO_DIR))


def have_full():
    return have_cores() and os.path.isfile(os.path.join(REF_SRC, "gcMapExplorer", "lib", "normalizer.py"))


_loaded = {}


def _prepare(full):
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)      # 2017-era regex / docstring escapes
    if "appdirs" not in sys.modules:
        class _AppDirs:
            def __init__(self, name):
                self.user_config_dir = os.path.join(tempfile.gettempdir(), "gcx_oracle_cfg")
        _stub_module("appdirs", system=sys.platform, AppDirs=_AppDirs)
    if "h5py" not in sys.modules:
        _stub_module("h5py", File=object, Group=object, Dataset=object)
    paths = [REF_SO_DIR]
    if full:
        paths.append(os.path.join(REF_SRC, "gcMapExplorer", "lib"))
    pkg = sys.modules.get("gcMapExplorer") or _stub_module("gcMapExplorer")
    pkg.__path__ = [os.path.join(REF_SRC, "gcMapExplorer")] if full else []
    lib = sys.modules.get("gcMapExplorer.lib") or _stub_module("gcMapExplorer.lib")
    lib.__path__ = paths
    lib.__package__ = "gcMapExplorer.lib"
    pkg.lib = lib
    if "gcMapExplorer.config" not in sys.modules or (full and not _loaded.get("full")):
        if full:
            sys.modules.pop("gcMapExplorer.config", None)
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):           # first import prints the default config
                importlib.import_module("gcMapExplorer.config")
        else:
            cfg = {"Dirs": {"WorkingDirectory": tempfile.gettempdir()}}
            _stub_module("gcMapExplorer.config", getConfig=lambda: cfg)
    if not full and "gcMapExplorer.lib.util" not in sys.modules:
        lib.util = _stub_module("gcMapExplorer.lib.util", kth_diag_indices=_kth_diag_indices)


def load_cores():
    """Return (normalizeCore, normalizeKnightRuiz, ccmapHelpers) -- the compiled reference cores."""
    if "cores" in _loaded:
        return _loaded["cores"]
    if not have_cores():
        raise ImportError("oracle/_ref is not built: run `make -C oracle ref` where /root/reference exists")
    _prepare(full=have_full())
    cmh = importlib.import_module("gcMapExplorer.lib.ccmapHelpers")
    core = importlib.import_module("gcMapExplorer.lib.normalizeCore")
    kr = importlib.import_module("gcMapExplorer.lib.normalizeKnightRuiz")
    proxy = _NumpyProxy()
    core.np = proxy
    kr.np = proxy
    _loaded["cores"] = (core, kr, cmh)
    _loaded["proxy"] = proxy
    return _loaded["cores"]


def ic_pass_counter(reset=False):
    """Number of IC passes the compiled reference core has run since the last reset."""
    n = _NumpyProxy.sqrt_calls
    if reset:
        _NumpyProxy.sqrt_calls = 0
    return n


def load_full():
    """Return the reference's ``lib.normalizer`` module (plus .ccmap/.cmstats reachable from it)."""
    if "full" in _loaded:
        return _loaded["full"]
    if not have_full():
        raise ImportError("full reference needs /root/reference (not present on the GPU box)")
    _prepare(full=True)
    load_cores()
    nz = importlib.import_module("gcMapExplorer.lib.normalizer")
    _loaded["full"] = nz
    return nz


# --------------------------------------------------------------------------------------------------
# the reference's C correlation core (lib/corrMatrixCoreSRC.c), compiled as it lies into
# oracle/_ref/gcMapExplorer/lib/libcorrMatrixCore.so.  Its Cython wrapper does not build against NumPy 2.x, so the
# stand-in below does what lib/_corrMatrixCore.pyx:49-98 does around the C call: zero-initialised float64 output,
# mask value rounded to float32 (:88) and passed by pointer, NULL for maskvalue=None.
# --------------------------------------------------------------------------------------------------
def have_corr():
    return os.path.isfile(os.path.join(REF_SO_DIR, "libcorrMatrixCore.so"))


class _CorrCore:
    def __init__(self):
        import ctypes as C
        self.C = C
        self.lib = C.CDLL(os.path.join(REF_SO_DIR, "libcorrMatrixCore.so"))
        dp = C.POINTER(C.c_double)
        for name in ("_correlationMatrix", "_covarianceMatrix"):
            f = getattr(self.lib, name)
            f.restype = C.c_int
            f.argtypes = [dp, dp, C.c_int, dp]

    def _run(self, fn, in_array, out, maskvalue):
        C = self.C
        a = _np.ascontiguousarray(in_array, dtype=_np.float64)
        size = a.shape[0]
        ret = out is None
        if out is None:
            out = _np.zeros((size, size), dtype=_np.float64)
        dp = C.POINTER(C.c_double)
        mask = None if maskvalue is None else C.byref(C.c_double(float(_np.float32(maskvalue))))
        # the C code reports progress on stdout (lib/corrMatrixCoreSRC.c:38-42); keep it off the test log
        fd = os.dup(1)
        devnull = os.open(os.devnull, os.O_WRONLY)
        try:
            sys.stdout.flush()
            os.dup2(devnull, 1)
            fn(a.ctypes.data_as(dp), out.ctypes.data_as(dp), size, mask)
        finally:
            os.dup2(fd, 1)
            os.close(fd)
            os.close(devnull)
        return out if ret else None

    def calculateCorrMatrix(self, in_array, out=None, maskvalue=None):
        return self._run(self.lib._correlationMatrix, in_array, out, maskvalue)

    def calculateCovMatrix(self, in_array, out=None, maskvalue=None):
        return self._run(self.lib._covarianceMatrix, in_array, out, maskvalue)


def load_corr():
    """The reference's compiled C correlation core behind the wrapper's calling convention."""
    if "corr" not in _loaded:
        if not have_corr():
            raise ImportError("oracle/_ref/libcorrMatrixCore.so is not built: run `make -C oracle ref` where /root/reference exists")
        _loaded["corr"] = _CorrCore()
    return _loaded["corr"]


def load_corr_module():
    """The reference's lib/corrMatrix.py (container only), with the stand-in above registered as its `_corrMatrixCore`."""
    if "corr_module" in _loaded:
        return _loaded["corr_module"]
    load_full()
    core = load_corr()
    _stub_module("gcMapExplorer.lib._corrMatrixCore", calculateCorrMatrix=core.calculateCorrMatrix, calculateCovMatrix=core.calculateCovMatrix,
                 calculateCorrelation=None, calculateCovariance=None, _calculateCorrMatrixOLDSLOW=None)
    _loaded["corr_module"] = importlib.import_module("gcMapExplorer.lib.corrMatrix")
    return _loaded["corr_module"]
