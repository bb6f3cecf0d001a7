"""Wall time of the drop-in API on a file-backed CCMAP at chr1 @ 10 kb (what a gcMapExplorer user would see)."""
import os, sys, time, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gcmapexplorer_b200 import device as gx, ccmap as cmp, normalizer as nz
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24926
work = tempfile.mkdtemp(prefix="gcx_probe_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
with gx.DeviceMap(n) as dm:
    dm.synth(seed=110)
    c = cmp.CCMAP()
    c.shape = (n, n); c.binsize = 10000; c.xlabel = c.ylabel = "chr1"; c.title = "chr1_vs_chr1"
    c.xticks = c.yticks = [0, n * 10000]
    c.gen_matrix_file(workDir=work)
    c.make_writable()
    dm.download(0, out=c.matrix)
    c.matrix.flush(); c.minvalue = 1.0; c.maxvalue = 1000.0
    c.make_unreadable()
res = {"n": n, "workdir": work}
for name, fn in (("IC", lambda: nz.normalizeCCMapByIC(c)), ("KR", lambda: nz.normalizeCCMapByKR(c)), ("VC", lambda: nz.normalizeCCMapByVCNorm(c)),
                 ("MCFS_median", lambda: nz.normalizeCCMapByMCFS(c))):
    fn()                      # warm (page cache, first-touch of the library)
    t0 = time.perf_counter()
    out = fn()
    res[name + "_s"] = round(time.perf_counter() - t0, 3)
    res[name + "_info"] = {k: v for k, v in nz.last_run_info.items() if k in ("passes", "outer", "MVP", "device_ms")}
    del out
# plain-copy floor on the same file system: read the 4 n^2-byte input once, write a 4 n^2-byte output once (what any implementation must do)
raw = np.empty(n * n, dtype=np.float32)
t0 = time.perf_counter()
with open(c.path2matrix, "rb") as f:
    f.readinto(raw)
t_read = time.perf_counter() - t0
t0 = time.perf_counter()
with open(os.path.join(work, "floor.out"), "wb") as f:
    f.write(raw)
t_write = time.perf_counter() - t0
os.remove(os.path.join(work, "floor.out"))
res["plain_copy_floor_s"] = {"read_input_single_thread": round(t_read, 3), "write_output_single_thread": round(t_write, 3),
                             "note": "single-threaded read() / write() of one 4 n^2-byte file each on the same file system; the library uses 16 pread / pwrite threads"}
print(json.dumps(res))
