"""CCMAP -- the data carrier the normalizers consume and emit.

Behavioural mirror of ``gcMapExplorer/lib/ccmap.py`` for what the normalization path needs
(reference lines in brackets): same attribute set [ccmap.py:125-140], same memmap lifecycle
(make_readable / make_unreadable / make_writable / make_editable [:224-262]), ``copy(fill)`` creating a
new temporary backing file next to the original [:264-300], ``gen_from_matrix`` [:188-221], the
``.ccmap`` JSON + raw float32 ``.npbin[.gz]`` on-disk format [:302-567] (files written here load in
the reference and vice versa) and ``checkCCMapObjectOrFile`` [:604-653] including its bare-``None`` return
for a missing file.  Downsampling, smoothing, export and plotting are out of scope.

If the real ``gcMapExplorer.lib.ccmap.CCMAP`` is importable, instances of it are accepted everywhere a
CCMAP is expected (duck-typed on the attributes above).
"""
import gzip
import json
import logging
import os
import shutil
import tempfile

import numpy as np

dtype_npBINarray = "float32"

logger = logging.getLogger("ccmap")
logger.setLevel(logging.INFO)

_ATTRS = ("path2matrix", "xticks", "yticks", "binsize", "title", "xlabel", "ylabel", "shape", "minvalue",
          "maxvalue", "bNoData", "bLog", "dtype", "matrix", "state")


def _default_workdir():
    return os.environ.get("GCX_WORKDIR") or tempfile.gettempdir()


def gen_backup(path):
    """Rename an existing file out of the way (``name`` -> ``name.bakN``) instead of overwriting it."""
    if not os.path.exists(path):
        return
    k = 0
    while os.path.exists(f"{path}.bak{k}"):
        k += 1
    os.rename(path, f"{path}.bak{k}")


class CCMAP:
    """Metadata + a float32 ``np.memmap`` backing file.  ``state`` is 'temporary', 'saved' or
    'compressed'; temporary/compressed objects unlink their backing file when collected."""

    def __init__(self, dtype=dtype_npBINarray):
        self.path2matrix = ""
        self.xticks = None
        self.yticks = None
        self.binsize = 0
        self.title = None
        self.xlabel = None
        self.ylabel = None
        self.shape = None
        self.minvalue = 9e10
        self.maxvalue = -9e10
        self.bNoData = None
        self.bLog = False
        self.dtype = dtype
        self.matrix = None
        self.state = "temporary"

    def __del__(self):
        try:
            mm = getattr(self, "matrix", None)
            if mm is not None and getattr(mm, "_mmap", None) is not None:
                mm._mmap.close()
            self.matrix = None
            if self.state in ("temporary", "compressed") and self.path2matrix and os.path.exists(self.path2matrix):
                os.remove(self.path2matrix)
        except Exception:
            pass

    # ---- ticks ----------------------------------------------------------------------------------
    def get_ticks(self, binsize=None):
        step = self.binsize if binsize is None else binsize
        return (np.arange(self.xticks[0], self.xticks[1], step), np.arange(self.yticks[0], self.yticks[1], step))

    # ---- backing file ---------------------------------------------------------------------------
    def gen_matrix_file(self, workDir=None):
        """Reserve a fresh temp file *name* (prefix gcx_npBinary_, like the reference) without keeping the file."""
        fd, self.path2matrix = tempfile.mkstemp(prefix="gcx_npBinary_", suffix=".tmp", dir=workDir or _default_workdir())
        os.close(fd)
        if os.path.isfile(self.path2matrix):
            os.remove(self.path2matrix)

    def _map(self, mode):
        self.matrix = None
        if not self.path2matrix and getattr(self, "_gz_source", None):
            # lazily opened compressed map (load_ccmap(lazy_gz=True)) that somebody wants to look at after all: inflate now
            self.gen_matrix_file(workDir=getattr(self, "_gz_workdir", None))
            with gzip.open(self._gz_source, "rb") as f_in, open(self.path2matrix, "wb") as f_out:
                shutil.copyfileobj(f_in, f_out)
            self._gz_source = None
        self.matrix = np.memmap(self.path2matrix, dtype=self.dtype, mode=mode, shape=tuple(self.shape))

    def make_readable(self):
        self._map("r")

    def make_editable(self):
        self._map("r+")

    def make_writable(self):
        self.matrix = None
        if os.path.exists(self.path2matrix):
            gen_backup(self.path2matrix)
        self._map("w+")

    def make_unreadable(self):
        self.matrix = None

    # ---- construction ---------------------------------------------------------------------------
    def gen_from_matrix(self, matrix, binsize, xlabel=None, ylabel=None, workDir=None):
        if self.matrix is not None:
            raise AssertionError("CCMAP object already contains a map.")
        if self.path2matrix:
            raise AssertionError("CCMAP object has already contains a memory mapped file")
        self.gen_matrix_file(workDir=workDir)
        self.shape = matrix.shape
        self.dtype = matrix.dtype
        self.binsize = binsize
        self.make_writable()
        self.matrix[:] = matrix
        self.xlabel = xlabel if xlabel is not None else "unknown"
        self.ylabel = ylabel if ylabel is not None else "unknown"
        self.title = self.xlabel + "_vs_" + self.ylabel
        self.xticks = [0, self.shape[0] * binsize]
        self.yticks = [0, self.shape[1] * binsize]
        self.maxvalue = np.max(self.matrix)
        nz = np.ma.masked_equal(matrix, 0.0, copy=False)
        self.minvalue = nz.min()

    def _clone_meta(self):
        other = CCMAP(dtype=self.dtype)
        for key, val in self.__dict__.items():
            other.__dict__[key] = val
        other.matrix = None
        other.state = "temporary"
        other._gz_source = None
        other.gen_matrix_file(workDir=os.path.dirname(self.path2matrix) or getattr(self, "_gz_workdir", None) or None)
        return other

    def copy(self, fill=None):
        """New temporary CCMAP with its own backing file in the same directory; values copied, or filled."""
        other = self._clone_meta()
        self.make_readable()
        other.make_writable()
        if fill is None:
            other.matrix[:] = self.matrix[:]
        else:
            other.matrix.fill(fill)
        other.matrix.flush()
        self.make_unreadable()
        other.make_unreadable()
        return other

    def new_like(self):
        """New temporary CCMAP with an *uninitialised* writable backing file (the GPU result is downloaded
        straight into it -- avoids the reference's extra n^2 memmap-to-memmap copy, lib/normalizer.py:561)."""
        other = self._clone_meta()
        other.make_writable()
        return other


def new_like(ccMapObj):
    """New temporary CCMAP (this module's class) with an uninitialised writable backing file and the metadata of `ccMapObj`,
    which may be this module's CCMAP *or* the reference's ``gcMapExplorer.lib.ccmap.CCMAP`` (duck-typed on the attribute set
    of lib/ccmap.py:125-140) -- what ``ccMapObj.copy(fill=0.0)`` gives the reference's normalizers (lib/normalizer.py:308, 561)
    without the n^2 fill."""
    if isinstance(ccMapObj, CCMAP):
        return ccMapObj.new_like()
    other = CCMAP(dtype=getattr(ccMapObj, "dtype", dtype_npBINarray))
    for key in _ATTRS:
        if key in ("matrix", "state", "path2matrix"):
            continue
        if hasattr(ccMapObj, key):
            setattr(other, key, getattr(ccMapObj, key))
    other.shape = tuple(ccMapObj.shape)
    other.state = "temporary"
    other.gen_matrix_file(workDir=os.path.dirname(getattr(ccMapObj, "path2matrix", "") or "") or None)
    other.make_writable()
    return other


def result_like(dm, which, ccMapObj):
    """New temporary CCMAP with the metadata of `ccMapObj`, holding a device result (`which`: 0 input buffer, 1 / 2 output buffers
    of the DeviceMap), returned editable (memmap r+).  float32 maps (lib/ccmap.py:42) are written by the library straight into the
    new backing file (gcx_map_download_file: device -> pinned stages -> parallel pwrite; no np.memmap page faults on the way);
    a CCMAP of another dtype (gen_from_matrix keeps the dtype of the array it was given, lib/ccmap.py:203) gets the result cast
    into its own dtype, like the reference's in-place arithmetic would."""
    if isinstance(ccMapObj, CCMAP):
        out = ccMapObj._clone_meta()
    else:
        out = new_like(ccMapObj)
        out.matrix = None
    if np.dtype(out.dtype) == np.float32 and hasattr(dm, "download_file"):
        out.matrix = None
        if os.path.exists(out.path2matrix):
            os.remove(out.path2matrix)
        dm.download_file(which, out.path2matrix)
        out.make_editable()
    else:
        if out.matrix is None:
            out.make_writable()
        out.matrix[:] = dm.download(which)
        out.matrix.flush()
    return out


def upload_ccmap(ccMapObj, make_map):
    """The CCMAP's matrix into HBM by the cheapest route.  `make_map(n)` returns an empty DeviceMap / ShardedDeviceMap.
    A plain float32 backing file goes through gcx_map_upload_file (parallel pread into pinned stages, row statistics per stage as
    it lands); a lazily opened `.npbin.gz` (load_ccmap(..., lazy_gz=True)) is inflated chunk by chunk straight into the device --
    no temporary file (the reference gunzips to a temp file first, lib/ccmap.py:550-563); anything else is uploaded from the memmap."""
    n = int(ccMapObj.shape[0])
    dm = make_map(n)
    try:
        gz = getattr(ccMapObj, "_gz_source", None)
        path = getattr(ccMapObj, "path2matrix", "")
        if gz is not None:
            _stream_gz(dm, gz, n)
        elif (np.dtype(ccMapObj.dtype) == np.float32 and path and os.path.isfile(path) and os.path.getsize(path) == 4 * n * n
              and hasattr(dm, "upload_file")):
            dm.upload_file(path)
        else:
            ccMapObj.make_readable()
            try:
                dm.upload(ccMapObj.matrix)
            finally:
                ccMapObj.make_unreadable()
    except Exception:
        dm.close()
        raise
    return dm


def _stream_gz(dm, gz_path, n, chunk_bytes=64 << 20):
    """zlib stream -> row chunks -> device (every chunk's row statistics are taken as it lands)."""
    import zlib
    row_bytes = 4 * n
    rows_per = max(1, chunk_bytes // row_bytes)
    buf = bytearray()
    r0 = 0
    dec = zlib.decompressobj(zlib.MAX_WBITS | 16)          # gzip container
    with open(gz_path, "rb") as f:
        while True:
            raw = f.read(8 << 20)
            if not raw:
                break
            buf += dec.decompress(raw)
            while len(buf) >= rows_per * row_bytes:
                take = rows_per * row_bytes
                dm.upload_rows(np.frombuffer(bytes(buf[:take]), dtype=np.float32).reshape(rows_per, n), r0)
                del buf[:take]
                r0 += rows_per
        buf += dec.flush()
    if len(buf) % row_bytes or r0 + len(buf) // row_bytes != n:
        raise ValueError("{0}: {1} bytes do not make a {2} x {2} float32 matrix".format(gz_path, r0 * row_bytes + len(buf), n))
    if buf:
        rows = len(buf) // row_bytes
        dm.upload_rows(np.frombuffer(bytes(buf), dtype=np.float32).reshape(rows, n), r0)
    dm.upload_finish()


# ---- coarsening (lib/ccmap.py:655-842) ------------------------------------------------------------------
def getOutputShapeFor2DMapDownsampling(inputShape, level=2):
    """Shape of the map after coarsening by `level` (lib/ccmap.py:712-735)."""
    return (int(np.ceil(float(inputShape[0] - 1) / level)) + 1, int(np.ceil(float(inputShape[1] - 1) / level)) + 1)


def downSample2DMap(inMatrix, outMatrix=None, level=2, method='sum'):
    """Coarsen a square matrix by `level` on the GPU (lib/ccmap.py:737-842): 'sum', 'mean' or 'max' over level x level blocks,
    row/column 0 left at zero, NumPy's arithmetic reproduced bit for bit.  Fills `outMatrix` when given, else returns the result
    (the reference returns after the first band in that case -- an obvious slip at :839-840 that is not reproduced)."""
    from .device import DeviceMap
    with DeviceMap.from_host(inMatrix) as dm:
        out, _, _ = dm.downsample(int(level), method)
    if outMatrix is None:
        return out
    outMatrix[:] = out
    return None


def downSampleCCMap(cmap, level=2, method='sum', workDir=None):
    """Coarsened copy of a CCMAP (lib/ccmap.py:655-710): new temporary map with binsize * level, bNoData OR-ed over each
    group of fine bins (bin 0 never flagged), minvalue / maxvalue of the result."""
    from .device import DeviceMap
    level = int(level)
    cmap.make_readable()
    with DeviceMap.from_host(cmap.matrix) as dm:
        out, mn, mx = dm.downsample(level, method)
    res = CCMAP(dtype=cmap.dtype)
    res.shape = out.shape
    res.gen_matrix_file(workDir=workDir)
    res.make_writable()
    res.matrix[:] = out
    res.matrix.flush()
    if cmap.bNoData is not None:
        b = np.zeros(out.shape[0], dtype=bool)
        pad = int((out.shape[0] - 1) * level - (cmap.shape[0] - 1))
        x = np.asarray(cmap.bNoData[1:], dtype=np.float64)
        if pad != 0:
            x = np.hstack((x, np.zeros(pad)))
        b[1:] = x.reshape(-1, level).sum(axis=1)
        res.bNoData = b
    res.minvalue = mn
    res.maxvalue = mx
    res.binsize = cmap.binsize * level
    res.xlabel = cmap.xlabel
    res.ylabel = cmap.ylabel
    return res


def downSample1D(array, level=2, masked_equal=0.0, func='max'):
    """Coarsen a 1-D track the way the maps are coarsened (lib/ccmap.py:845-870): element 0 is kept apart, the
    rest is zero-padded to a multiple of `level` and reduced in groups ('sum', 'mean' over the entries different
    from `masked_equal`, anything else: max).  O(n) host work."""
    level = int(level)
    groups = int(np.ceil(float(array.shape[0] - 1) / level))
    pad = int(groups * level - (array.shape[0] - 1))
    out = np.zeros(groups + 1, dtype=array.dtype)
    body = np.hstack((array[1:], np.zeros(pad))) if pad != 0 else array[1:]
    body = body.reshape(-1, level)
    if func == 'sum':
        out[1:] = body.sum(axis=1)
    elif func == 'mean':
        out[1:] = np.ma.filled(np.ma.mean(np.ma.masked_equal(body, masked_equal), axis=1))[:]
    else:
        out[1:] = body.max(axis=1)
    return out


def is_ccmap(obj):
    return isinstance(obj, CCMAP) or all(hasattr(obj, a) for a in ("path2matrix", "shape", "make_readable", "bNoData", "state"))


# ---- (de)serialisation ----------------------------------------------------------------------------
def jsonify(ccMapObj):
    """In-place conversion of attributes to JSON-safe strings (same encoding as the reference, ccmap.py:302-357)."""
    if ccMapObj.bNoData is not None:
        ccMapObj.bNoData = "".join("1" if b else "0" for b in np.asarray(ccMapObj.bNoData))
    if ccMapObj.xticks is not None:
        ccMapObj.xticks = [str(v) for v in ccMapObj.xticks]
    if ccMapObj.yticks is not None:
        ccMapObj.yticks = [str(v) for v in ccMapObj.yticks]
    ccMapObj.minvalue = str(float(ccMapObj.minvalue))
    ccMapObj.maxvalue = str(float(ccMapObj.maxvalue))
    ccMapObj.shape = [str(v) for v in ccMapObj.shape]
    if hasattr(ccMapObj, "binsize"):
        ccMapObj.binsize = str(ccMapObj.binsize)
    if isinstance(ccMapObj.dtype, np.dtype):
        ccMapObj.dtype = ccMapObj.dtype.name


def dejsonify(ccMapObj, json_dict=None):
    if json_dict is not None:
        ccMapObj.__dict__.update(json_dict)
    if ccMapObj.bNoData is not None:
        ccMapObj.bNoData = np.array([ch == "1" for ch in ccMapObj.bNoData], dtype=bool)
    if ccMapObj.xticks is not None:
        ccMapObj.xticks = [int(v) for v in ccMapObj.xticks]
    if ccMapObj.yticks is not None:
        ccMapObj.yticks = [int(v) for v in ccMapObj.yticks]
    ccMapObj.minvalue = float(ccMapObj.minvalue)
    ccMapObj.maxvalue = float(ccMapObj.maxvalue)
    ccMapObj.shape = tuple(int(v) for v in ccMapObj.shape)
    if hasattr(ccMapObj, "binsize"):
        ccMapObj.binsize = int(ccMapObj.binsize)


def save_ccmap(ccMapObj, outfile, compress=False, logHandler=None):
    """Write ``<name>.ccmap`` (JSON) + ``<name>.npbin`` (raw float32, gzip'd when compress=True)."""
    log = logging.getLogger("save_ccmap")
    if logHandler is not None:
        log.propagate = False
        log.addHandler(logHandler)
    log.setLevel(logging.INFO)
    if getattr(ccMapObj, "_gz_source", None) and not ccMapObj.path2matrix:
        ccMapObj.make_readable()          # lazily opened compressed map: materialise its matrix first
    ccMapObj.make_unreadable()
    if os.path.splitext(outfile)[1] != ".ccmap":
        outfile = outfile + ".ccmap"
    outpath = os.path.abspath(os.path.expanduser(outfile))
    outdir = os.path.dirname(outpath)
    npbin_name = os.path.splitext(os.path.basename(outfile))[0] + ".npbin"
    npbin = os.path.join(outdir, npbin_name)
    log.info(" Saving ccmap to file [{0}] and [{1}] ...".format(outfile, npbin))
    if ccMapObj.path2matrix == npbin:
        logging.warning("File {0} already exists !!!".format(ccMapObj.path2matrix))
    else:
        if os.path.exists(npbin):
            gen_backup(npbin)
        shutil.copy(ccMapObj.path2matrix, npbin)
    if os.path.exists(outpath):
        gen_backup(outpath)
    keep_path, keep_state = ccMapObj.path2matrix, ccMapObj.state
    ccMapObj.path2matrix = npbin_name
    jsonify(ccMapObj)
    ccMapObj.state = "compressed" if compress else "saved"
    try:
        with open(outpath, "w") as fout:
            json.dump({k: v for k, v in ccMapObj.__dict__.items() if not k.startswith("_")}, fout, indent=4, separators=(",", ":"))   # matrix is None here
        if compress:
            if os.path.exists(npbin + ".gz"):
                gen_backup(npbin + ".gz")
            with open(npbin, "rb") as f_in, gzip.open(npbin + ".gz", "wb") as f_out:
                shutil.copyfileobj(f_in, f_out)
            if os.path.exists(npbin):
                os.remove(npbin)
    finally:
        ccMapObj.path2matrix = keep_path
        ccMapObj.state = "temporary" if keep_state == "temporary" else keep_state
        dejsonify(ccMapObj)
    log.info("       Finished!!!\n")


def load_ccmap(infile, workDir=None, lazy_gz=False):
    """lazy_gz=True (used by the normalizers for file inputs): a compressed matrix is NOT inflated into a temporary file; the
    object carries `_gz_source` and upload_ccmap streams it into the device.  Such an object has no backing file: make_readable
    inflates it on first use, so code that does look at `.matrix` still works."""
    with open(infile, "r") as fin:
        jd = json.load(fin)
    obj = CCMAP(dtype=np.dtype(jd["dtype"]))
    dejsonify(obj, jd)
    obj.matrix = None
    indir = os.path.dirname(os.path.abspath(os.path.expanduser(infile)))
    npbin = os.path.join(indir, obj.path2matrix)
    if os.path.exists(npbin + ".gz") and lazy_gz and np.dtype(obj.dtype) == np.float32:
        obj._gz_source = npbin + ".gz"
        obj._gz_workdir = workDir
        obj.path2matrix = ""
        obj.state = "temporary"
        return obj
    if os.path.exists(npbin + ".gz"):
        obj.gen_matrix_file(workDir=workDir)
        try:
            with gzip.open(npbin + ".gz", "rb") as f_in, open(obj.path2matrix, "wb") as f_out:
                shutil.copyfileobj(f_in, f_out)
        except (KeyboardInterrupt, SystemExit):
            os.remove(obj.path2matrix)
            raise
        obj.state = "temporary"
    else:
        obj.path2matrix = npbin
    return obj


def checkCCMapObjectOrFile(ccMap, workDir=None, lazy_gz=False):
    """(CCMAP, 'Object' | 'File'); a missing file returns a bare None like the reference (ccmap.py:644-647),
    so callers that tuple-unpack raise TypeError."""
    if is_ccmap(ccMap):
        return ccMap, "Object"
    if not os.path.isfile(ccMap):
        logger.info(" {0} file not found.".format(ccMap))
        logger.info(" Not able to normalize.")
        return None
    return load_ccmap(ccMap, workDir=workDir, lazy_gz=lazy_gz), "File"
