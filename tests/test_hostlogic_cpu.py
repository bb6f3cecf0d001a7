"""CPU: host-side logic added in round 2 -- map-set containers of the batch drivers, the single-process sharding wrapper's
row bookkeeping, lazily opened compressed maps.  No GPU: the device layer is replaced by recording stubs where needed."""
import os

import numpy as np
import pytest

from oracle import oracle_np as onp


def _ccmap(A, name, tmp_path, binsize=10000):
    from gcmapexplorer_b200 import ccmap as cmp
    c = cmp.CCMAP()
    c.gen_from_matrix(A, binsize, xlabel=name, ylabel=name, workDir=str(tmp_path))
    c.bNoData = ~onp.get_nonzeros_index(A)
    return c


def test_mapset_directory_lists_maps_smallest_first_and_resolutions(tmp_path):
    """lib/gcmap.py:219-262 (sortBy='size') and the finer -> coarser walk of lib/normalizer.py:1003-1030 on the directory carrier."""
    from gcmapexplorer_b200 import mapset, ccmap as cmp
    root = tmp_path / "set"
    sizes = {"chr2": 60, "chr10": 25, "chrX": 41}
    for name, n in sizes.items():
        mapset.add_map(_ccmap(onp.synth_map(n, seed=n), name, tmp_path), str(root), generateCoarse=False)
    # a second, coarser resolution of chr2 stored by hand
    mapset.add_map(_ccmap(onp.synth_map(31, seed=3), "chr2", tmp_path, binsize=20000), str(root), generateCoarse=False, replaceCMap=False)
    ms = mapset.MapSet(str(root))
    assert ms.names_smallest_first() == ["chr10", "chrX", "chr2"]
    assert ms.resolutions("chr2") == ["10kb", "20kb"] and ms.resolutions("chr10") == ["10kb"]
    c = ms.load("chr2", workDir=str(tmp_path))
    assert tuple(c.shape) == (60, 60) and c.binsize == 10000 and getattr(c, "_gz_source", None)       # compressed: opened lazily
    c.make_readable()
    assert np.array_equal(np.array(c.matrix), onp.synth_map(60, seed=60))
    c20 = ms.load("chr2", resolution="20kb", workDir=str(tmp_path))
    assert tuple(c20.shape) == (31, 31)
    # a flat list of .ccmap files is a map set too
    files = [str(root / "chr10" / "10kb.ccmap"), str(root / "chrX" / "10kb.ccmap")]
    assert mapset.MapSet(files).names_smallest_first() == ["chr10", "chrX"]
    with pytest.raises(IOError):
        mapset.MapSet(str(tmp_path / "empty_dir_that_does_not_exist"))
    with pytest.raises((ImportError, OSError)):
        mapset.MapSet(str(tmp_path / "x.gcmap"))              # HDF5 carrier: needs h5py (absent here)


def test_lazy_gz_ccmap_materialises_on_demand_and_saves(tmp_path):
    from gcmapexplorer_b200 import ccmap as cmp
    A = onp.synth_map(33, seed=2)
    c = _ccmap(A, "chrL", tmp_path)
    cmp.save_ccmap(c, str(tmp_path / "l.ccmap"), compress=True)
    lazy = cmp.load_ccmap(str(tmp_path / "l.ccmap"), workDir=str(tmp_path), lazy_gz=True)
    assert lazy.path2matrix == "" and lazy._gz_source.endswith("l.npbin.gz")
    clone = cmp.new_like(lazy)                                  # output carrier next to a file-less input: lands in workDir
    assert os.path.dirname(clone.path2matrix) == str(tmp_path) and clone._gz_source is None
    cmp.save_ccmap(lazy, str(tmp_path / "again.ccmap"), compress=False)     # materialises first
    back = cmp.load_ccmap(str(tmp_path / "again.ccmap"))
    back.make_readable()
    assert np.array_equal(np.array(back.matrix), A)
    import json
    assert not any(k.startswith("_") for k in json.load(open(tmp_path / "again.ccmap")))


def test_sharded_wrapper_row_bookkeeping(monkeypatch):
    """ShardedDeviceMap hands every rank exactly its rows: whole-map uploads, global row chunks of a stream, downloads."""
    from gcmapexplorer_b200 import sharded

    class FakeMap:
        def __init__(self, n, row0, nloc, device=None, dist=None, options=None):
            self.n, self.row0, self.nloc, self.rank = n, row0, nloc, dist[0]
            self.rows = np.full((nloc, n), np.nan, dtype=np.float32)
            self.options = dict(options or {})
            self.finished = False

        def upload(self, rows):
            assert rows.shape == (self.nloc, self.n)
            self.rows[:] = rows

        def upload_rows(self, rows, row_begin, with_stats=True):
            self.rows[row_begin:row_begin + rows.shape[0]] = rows

        def upload_finish(self):
            self.finished = True

        def download(self, which=1, out=None):
            out[:] = self.rows * 2
            return out

        def close(self):
            pass
    monkeypatch.setattr(sharded, "DeviceMap", FakeMap)
    monkeypatch.setattr(sharded, "dist_unique_id", lambda: b"\0" * 128)
    n = 5000
    A = np.arange(n * n, dtype=np.float32).reshape(n, n)
    sm = sharded.ShardedDeviceMap(n, devices=[0, 1, 2])
    assert [p[0] for p in sm.parts] == [0, 2048, 3072] and sum(p[1] for p in sm.parts) == n      # cuts on multiples of 1024 rows
    assert all(m.options.get("exchange_allreduce") == 1 for m in sm.maps)
    sm.upload(A)
    for m in sm.maps:
        assert np.array_equal(m.rows, A[m.row0:m.row0 + m.nloc])
        m.rows[:] = np.nan
    for r0 in range(0, n, 700):                                 # a stream of global row chunks that straddle the cuts
        sm.upload_rows(A[r0:r0 + 700], r0)
    sm.upload_finish()
    for m in sm.maps:
        assert np.array_equal(m.rows, A[m.row0:m.row0 + m.nloc]) and m.finished
    assert np.array_equal(sm.download(1), 2 * A)
    sm.close()
    with pytest.raises(ValueError):
        sharded.ShardedDeviceMap(n, devices=[0])


def test_gpus_for_map_env(monkeypatch):
    from gcmapexplorer_b200 import sharded
    monkeypatch.setattr(sharded, "visible_gpus", lambda: 4)
    monkeypatch.setenv("GCX_GPUS", "2")
    assert sharded.gpus_for_map(1000) == 2
    monkeypatch.setenv("GCX_GPUS", "16")
    assert sharded.gpus_for_map(1000) == 4                      # capped at what is visible
    monkeypatch.setenv("GCX_GPUS", "1")
    assert sharded.gpus_for_map(10 ** 6) == 1
