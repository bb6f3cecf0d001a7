"""Where does the drop-in call's wall time go?  Stage-by-stage timing of the file <-> HBM path at chr1 @ 10 kb on tmpfs."""
import os, sys, time, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from gcmapexplorer_b200 import device as gx, ccmap as cmp, normalizer as nz
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24926
work = tempfile.mkdtemp(prefix="gcx_io_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
src = os.path.join(work, "in.npbin")
res = {"n": n, "bytes": 4 * n * n, "cores": os.cpu_count()}

def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return round(best, 4)

with gx.DeviceMap(n) as dm:
    dm.synth(seed=110)
    dm.download_file(0, src)
    res["upload_file_s"] = timed(lambda: dm.upload_file(src))
    res["upload_file_nostats_s"] = timed(lambda: dm.upload_file(src, with_stats=False))
    mm = np.memmap(src, dtype=np.float32, mode="r", shape=(n, n))
    res["upload_memmap_s"] = timed(lambda: dm.upload(mm))
    dst = os.path.join(work, "out.npbin")
    def dl_new():
        if os.path.exists(dst): os.remove(dst)
        dm.download_file(0, dst)
    res["download_file_new_s"] = timed(dl_new)
    res["download_file_overwrite_s"] = timed(lambda: dm.download_file(0, dst))
    def dl_mm():
        out = np.memmap(os.path.join(work, "mm.npbin"), dtype=np.float32, mode="w+", shape=(n, n))
        dm.download(0, out=out); out.flush(); del out
        os.remove(os.path.join(work, "mm.npbin"))
    res["download_memmap_new_s"] = timed(dl_mm, 2)
    pin = gx.PinnedArray((n, n), pitched=True)
    res["download_pinned_s"] = timed(lambda: dm.download(0, out=pin.array))
    res["upload_pinned_s"] = timed(lambda: dm.upload(pin.array))
    pin.free()
    t0 = time.perf_counter(); os.remove(dst); res["unlink_s"] = round(time.perf_counter() - t0, 4)
# plain parallel pread / pwrite of the same file with Python threads (no GPU): the file system's own floor
buf = np.empty(n * n, dtype=np.float32)
bv = memoryview(buf).cast("B")
def par_read(T):
    fd = os.open(src, os.O_RDONLY); tot = 4 * n * n; step = (tot + T - 1) // T
    def work_(t):
        a, b = t * step, min(tot, (t + 1) * step)
        while a < b:
            a += os.preadv(fd, [bv[a:b]], a)
    with ThreadPoolExecutor(T) as ex: list(ex.map(work_, range(T)))
    os.close(fd)
def par_write(T):
    p = os.path.join(work, "w.npbin")
    if os.path.exists(p): os.remove(p)
    fd = os.open(p, os.O_WRONLY | os.O_CREAT); tot = 4 * n * n; step = (tot + T - 1) // T
    os.ftruncate(fd, tot)
    def work_(t):
        a, b = t * step, min(tot, (t + 1) * step)
        while a < b:
            a += os.pwritev(fd, [bv[a:b]], a)
    with ThreadPoolExecutor(T) as ex: list(ex.map(work_, range(T)))
    os.close(fd)
for T in (1, 4, 16):
    res["plain_pread_%dthr_s" % T] = timed(lambda: par_read(T), 2)
    res["plain_pwrite_new_%dthr_s" % T] = timed(lambda: par_write(T), 2)
os.remove(os.path.join(work, "w.npbin"))
c = cmp.CCMAP()
c.shape = (n, n); c.binsize = 10000; c.xlabel = c.ylabel = "chr1"; c.title = "chr1_vs_chr1"; c.xticks = c.yticks = [0, n * 10000]
c.path2matrix = src; c.state = "saved"; c.minvalue = 1.0; c.maxvalue = 1000.0
for name, fn in (("IC", lambda: nz.normalizeCCMapByIC(c)), ("KR", lambda: nz.normalizeCCMapByKR(c))):
    o = fn(); del o
    t0 = time.perf_counter(); o = fn(); res["normalizeCCMapBy%s_s" % name] = round(time.perf_counter() - t0, 3)
    t0 = time.perf_counter(); del o; res["del_result_%s_s" % name] = round(time.perf_counter() - t0, 3)
print(json.dumps(res))
import shutil; shutil.rmtree(work, ignore_errors=True)
