"""Where does a launch of the deferred-scalar IC pass spend its time?  Per-CTA globaltimer stamps of one launch (GCX_SYM_DEBUG=1):
entry -> first tiles issued -> vector phase -> grid barrier -> product (static share, dynamic chunks) -> end."""
import os, sys, ctypes as C
os.environ["GCX_SYM_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gcmapexplorer_b200 import device as gx, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24926
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:          # under torchrun: the row-sharded pass with the NVLink peer exchange, rank 0 reports
    import torch, torch.distributed as tdist
    from gcmapexplorer_b200 import dist as gd
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    tdist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    ctxmgr = gd.sharded_device_map(n, device=int(os.environ["LOCAL_RANK"]), partition="balanced", symmetric=True)
    if int(os.environ["RANK"]) != 0:
        sys.stdout = open(os.devnull, "w")
else:
    ctxmgr = gx.DeviceMap(n)
with ctxmgr as dm:
    dm.synth(seed=110)
    dm.set_mask(np.ones(n, dtype=bool))
    for rep in range(2):
        r = dm.ic(1e-4, 500)
    print("peer exchange:", r.get("peer_exchange"), "world", world)
    buf = (C.c_int64 * (16 * 1024))(); cnt = C.c_int(0)
    _lib.check(_lib.lib().gcx_debug_sym_stamps(dm._ctx, buf, 16 * 1024, C.byref(cnt)))
    G = cnt.value
    st = np.array(buf[:16 * G], dtype=np.int64).reshape(G, 16)
    t0 = st[:, 0].min()
    names = ["entry", "tiles issued", "vector phase done", "past grid barrier", "product starts", "static share done", "all done"]
    for k, nm in ((14, "past the pass barrier"), (15, "fold + vector update done (thread 0)")):
        v = (st[:, k] - t0) / 1e3
        print("%-38s min %8.2f  median %8.2f  max %8.2f us" % (nm, v.min(), np.median(v), v.max()))
    if world > 1:
        for k, nm in ((12, "partials stored to peers"), (13, "peers' partials arrived")):
            v = (st[:, k] - t0) / 1e3
            print("%-26s min %8.2f  median %8.2f  max %8.2f us" % (nm, v.min(), np.median(v), v.max()))
    print("n", n, "passes", r["passes"], "device ms", r["ms"], "-> us per launch", r["ms"] * 1e3 / r["products_launched"], "G", G)
    for k, nm in enumerate(names):
        v = (st[:, k] - t0) / 1e3
        print("%-20s min %8.2f  p10 %8.2f  median %8.2f  p90 %8.2f  max %8.2f us" % (nm, v.min(), np.percentile(v, 10), np.median(v), np.percentile(v, 90), v.max()))
    ch = st[:, 7]
    print("dynamic chunks per CTA: min %d median %d max %d total %d" % (ch.min(), np.median(ch), ch.max(), ch.sum()))
    end = (st[:, 6] - t0) / 1e3
    stat = (st[:, 5] - t0) / 1e3
    print("static done by SM half (CTA < 148 / >= 148): %.1f / %.1f ; end: %.1f / %.1f" % (stat[:148].mean(), stat[148:].mean(), end[:148].mean(), end[148:].mean()))
    order = np.argsort(-end)
    print("last CTAs to finish:", [(int(c), round(float(end[c]), 1), int(ch[c])) for c in order[:10]])
    print("first CTAs to finish:", [(int(c), round(float(end[c]), 1), int(ch[c])) for c in order[-6:]])
    # who is slow?  static-share time per CTA against its SM, its position in the unit list and the panels it walks
    dur = (st[:, 5] - st[:, 4]) / 1e3
    smid, kf, kl = st[:, 9], st[:, 10], st[:, 11]
    print("static share duration: min %.1f median %.1f max %.1f us" % (dur.min(), np.median(dur), dur.max()))
    print("by CTA-index decile (panel-major position):", [round(float(dur[i].mean()), 1) for i in np.array_split(np.arange(G), 10)])
    order = np.argsort(smid)
    print("by SM-id decile:", [round(float(dur[i].mean()), 1) for i in np.array_split(order, 10)])
    print("corr(duration, CTA index) %.3f  corr(duration, smid) %.3f  corr(duration, first panel) %.3f  corr(duration, panels spanned) %.3f" % (
        np.corrcoef(dur, np.arange(G))[0, 1], np.corrcoef(dur, smid)[0, 1], np.corrcoef(dur, kf)[0, 1], np.corrcoef(dur, kl - kf)[0, 1]))
    same_sm = {}
    for c in range(G):
        same_sm.setdefault(int(smid[c]), []).append(c)
    pairs = [v for v in same_sm.values() if len(v) == 2]
    if pairs:
        d = np.array([[dur[a], dur[b]] for a, b in pairs])
        print("SMs with two CTAs: %d; corr between the two CTAs of an SM %.3f; mean |difference| %.1f us" % (len(pairs), np.corrcoef(d[:, 0], d[:, 1])[0, 1], np.abs(d[:, 0] - d[:, 1]).mean()))
    print("vector phase by CTA (us): p50 %.1f p90 %.1f max %.1f" % tuple(np.percentile((st[:, 2] - st[:, 1]) / 1e3, [50, 90, 100])))
