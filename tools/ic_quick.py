"""IC at chr1 @ 10 kb a few times: microseconds per pass for the current environment knobs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gcmapexplorer_b200 import device as gx
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24926
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
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
    best = 1e9
    for rep in range(5):
        r = dm.ic(1e-4, 500)
        best = min(best, r["ms"] * 1e3 / r["products_launched"])
    tag = " ".join("%s=%s" % (k, os.environ[k]) for k in sorted(os.environ) if k.startswith("GCX_"))
    print("world %d peer %s " % (world, r.get("peer_exchange")), end="")
    print("%-60s n %d passes %d: %.2f us per pass, %.3f ms to converge, frac %.4f" % (tag, n, r["passes"], best, r["ms"], r["pass_bytes"] / best / 1e3 / 6534.8), flush=True)
