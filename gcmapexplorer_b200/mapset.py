"""Map sets: the containers the batch drivers ``normalizeGCMapBy*`` read and write.

The reference keeps all chromosomes of a genome, each at a pyramid of resolutions, in one HDF5 file (``.gcmap``,
lib/gcmap.py).  h5py is not part of every installation (it is absent from the image this package was built in), so a map set
has two interchangeable carriers here:

* a DIRECTORY in the same hierarchy -- ``<root>/<mapName>/<resolution>.ccmap`` + ``.npbin[.gz]`` (the reference's own
  ``.ccmap`` file format, lib/ccmap.py:302-567; group = chromosome, dataset = resolution, ``bNoData`` inside the JSON) -- or a
  flat directory / list of ``.ccmap`` files, one map each;
* the HDF5 ``.gcmap`` file itself, through h5py when it can be imported (same groups, datasets and attributes as
  lib/gcmap.py:710-998: group attrs xlabel / ylabel / compression; dataset attrs minvalue / maxvalue / binsize / xshape / yshape;
  ``<resolution>-bNoData``).  Without h5py such a path raises ImportError at call time, nothing earlier.

The driver semantics are the reference's: maps in ascending size (lib/gcmap.py:219-262, 345-357), every normalized map added
with its coarser levels -- a factor of two per level until the map is smaller than 500 bins (lib/gcmap.py:582-604), coarsened
on the GPU (ccmap.downSampleCCMap, bit-identical to the reference).
"""
import glob
import os

import numpy as np

from . import ccmap as cmp
from . import util


def _h5py():
    try:
        import h5py
        if not hasattr(h5py, "File") or not hasattr(h5py, "Dataset") or not getattr(h5py, "__version__", None):
            raise ImportError("the module named h5py that is loaded is a stub, not the HDF5 binding")
        return h5py
    except ImportError as e:
        raise ImportError("reading or writing the HDF5 .gcmap container needs h5py, which is not installed here; pass a map-set "
                          "directory (<root>/<map>/<resolution>.ccmap) or a list of .ccmap files instead") from e


def is_hdf5_path(path):
    return isinstance(path, str) and not os.path.isdir(path) and os.path.splitext(path)[1].lower() in (".gcmap", ".h5", ".hdf5")


class MapSet:
    """Read side: names of the maps in ascending size, the resolutions stored for each (finest first) and their CCMAPs."""

    def __init__(self, source):
        self.source = source
        self.entries = {}            # mapName -> {resolution: loader(workDir) -> CCMAP}
        self.sizes = {}              # mapName -> bins of the finest resolution
        if isinstance(source, (list, tuple)):
            for p in source:
                self._add_ccmap_file(p)
        elif is_hdf5_path(source):
            self._open_hdf5(source)
        elif os.path.isdir(source):
            for p in sorted(glob.glob(os.path.join(source, "*.ccmap"))):
                self._add_ccmap_file(p)
            for d in sorted(glob.glob(os.path.join(source, "*", ""))):
                name = os.path.basename(os.path.dirname(d))
                for p in sorted(glob.glob(os.path.join(d, "*.ccmap"))):
                    self._add_ccmap_file(p, name=name, resolution=os.path.splitext(os.path.basename(p))[0])
        else:
            raise IOError("{0}: not a map-set directory, a list of .ccmap files or a .gcmap file".format(source))
        if not self.entries:
            raise IOError("{0}: no maps found".format(source))

    def _add_ccmap_file(self, path, name=None, resolution=None):
        import json
        with open(path) as f:
            jd = json.load(f)
        shape = tuple(int(v) for v in jd["shape"])
        if name is None:
            name = jd.get("xlabel") if jd.get("xlabel") == jd.get("ylabel") else "{0}-{1}".format(jd.get("xlabel"), jd.get("ylabel"))
            if name in self.entries or not name or name == "None":
                name = os.path.splitext(os.path.basename(path))[0]
        if resolution is None:
            resolution = util.binsizeToResolution(int(jd["binsize"]))
        self.entries.setdefault(name, {})[resolution] = lambda workDir=None, p=path: cmp.load_ccmap(p, workDir=workDir, lazy_gz=True)
        finest = min(self.entries[name], key=util.resolutionToBinsize)
        if finest == resolution:
            self.sizes[name] = max(shape)

    def _open_hdf5(self, path):
        h5py = _h5py()
        with h5py.File(path, "r") as h:
            for name in h.keys():
                for res in [k for k in h[name].keys() if "bNoData" not in k]:
                    self.entries.setdefault(name, {})[res] = lambda workDir=None, n=name, r=res: _load_hdf5_map(path, n, r, workDir)
                finest = min(self.entries[name], key=util.resolutionToBinsize)
                d = h[name][finest]
                self.sizes[name] = max(int(d.attrs["xshape"]), int(d.attrs["yshape"]))

    def names_smallest_first(self):
        """lib/gcmap.py:219-262 genMapNameList(sortBy='size'): ascending number of bins of the finest map."""
        names = list(self.entries)
        order = np.argsort([self.sizes[k] for k in names], kind="stable")
        return [names[i] for i in order]

    def resolutions(self, name):
        """Stored resolutions of a map, finest first (the walk of lib/normalizer.py:1003-1030 goes finer -> coarser)."""
        return sorted(self.entries[name], key=util.resolutionToBinsize)

    def load(self, name, resolution=None, workDir=None):
        res = self.resolutions(name)[0] if resolution is None else resolution
        return self.entries[name][res](workDir)


def _load_hdf5_map(path, name, resolution, workDir):
    """lib/gcmap.py:loadGCMapAsCCMap for one (group, resolution), into a temporary CCMAP."""
    h5py = _h5py()
    with h5py.File(path, "r") as h:
        g = h[name]
        d = g[resolution]
        c = cmp.CCMAP(dtype=d.dtype)
        c.xlabel, c.ylabel = str(g.attrs["xlabel"]), str(g.attrs["ylabel"])
        c.minvalue, c.maxvalue, c.binsize = float(d.attrs["minvalue"]), float(d.attrs["maxvalue"]), int(d.attrs["binsize"])
        c.shape = (int(d.attrs["xshape"]), int(d.attrs["yshape"]))
        c.xticks, c.yticks = [0, c.shape[0] * c.binsize], [0, c.shape[1] * c.binsize]
        c.title = c.xlabel + "_vs_" + c.ylabel
        if resolution + "-bNoData" in g:
            c.bNoData = np.asarray(g[resolution + "-bNoData"][:], dtype=bool)
        c.gen_matrix_file(workDir)
        c.make_writable()
        d.read_direct(c.matrix)
        c.matrix.flush()
        c.make_unreadable()
    return c


def _group_name(c):
    return c.xlabel if c.xlabel == c.ylabel else "{0}-{1}".format(c.xlabel, c.ylabel)


def add_map(cmap, dest, compression="lzf", generateCoarse=True, coarseningMethod="sum", replaceCMap=True, compress_files=True):
    """addCCMap2GCMap (lib/gcmap.py:831-998) for both carriers: the map at its own resolution and -- generateCoarse -- every
    coarser level (x2 until both sides are below 500 bins, lib/gcmap.py:599-604), each coarsened on the GPU from the level
    before it.  Returns the list of resolutions written."""
    if coarseningMethod not in ("sum", "mean", "max"):
        raise ValueError(' "{0}" is not a valid keyword to downsample. Use: "sum", "mean" or "max". '.format(coarseningMethod))
    written = []
    level_map = cmap
    owned = False
    while True:
        res = util.binsizeToResolution(level_map.binsize)
        if is_hdf5_path(dest):
            _write_hdf5_level(level_map, dest, res, compression, replace=(replaceCMap and level_map is cmap))
        else:
            d = os.path.join(dest, _group_name(cmap))
            os.makedirs(d, exist_ok=True)
            cmp.save_ccmap(level_map, os.path.join(d, res + ".ccmap"), compress=compress_files)
        written.append(res)
        if not generateCoarse or not (level_map.shape[0] > 500 or level_map.shape[1] > 500):
            break
        coarse = cmp.downSampleCCMap(level_map, level=2, method=coarseningMethod)
        coarse.title = cmap.title
        coarse.xticks, coarse.yticks = [0, coarse.shape[0] * coarse.binsize], [0, coarse.shape[1] * coarse.binsize]
        if owned:
            del level_map
        level_map, owned = coarse, True
    return written


def _write_hdf5_level(c, path, resolution, compression, replace):
    h5py = _h5py()
    name = _group_name(c)
    with h5py.File(path, "a") as h:
        if replace and name in h:
            del h[name]
        g = h.require_group(name)
        g.attrs["xlabel"], g.attrs["ylabel"], g.attrs["compression"] = c.xlabel, c.ylabel, compression
        for key in (resolution, resolution + "-bNoData"):
            if key in g:
                del g[key]
        c.make_readable()
        d = g.create_dataset(resolution, tuple(c.shape), dtype=c.dtype, data=c.matrix, chunks=True, compression=compression, shuffle=True)
        if c.bNoData is not None:
            b = np.asarray(c.bNoData, dtype=bool)
            g.create_dataset(resolution + "-bNoData", b.shape, dtype=b.dtype, data=b, chunks=True, compression=compression, shuffle=True)
        if c.minvalue == 0:
            c.minvalue = float(np.ma.masked_equal(c.matrix, 0.0, copy=False).min())
        d.attrs["minvalue"], d.attrs["maxvalue"], d.attrs["binsize"] = c.minvalue, c.maxvalue, c.binsize
        d.attrs["xshape"], d.attrs["yshape"] = c.shape[0], c.shape[1]
        c.make_unreadable()
