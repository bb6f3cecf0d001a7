"""Sweep the compiled k_matvec (rows-per-band, occupancy) variants at a given n; prints GB/s per variant."""
import sys, json
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from gcmapexplorer_b200 import device as gx
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24926
names = {0: "R8/occ2", 1: "R8/occ3", 2: "R4/occ4", 3: "R4/occ6", 4: "R16/occ1", 5: "R2/occ8", 6: "R8/occ4",
         7: "TMA R4/st6/occ2", 8: "TMA R4/st12/occ1", 9: "TMA R8/st6/occ1", 10: "TMA R4/st4/occ3", 11: "TMA R2/st8/occ3",
         12: "TMA R4/st7/occ2", 13: "TMA R8/st3/occ2", 14: "TMA R2/st12/occ2", 15: "TMA R4/st5/occ2", 16: "TMA R2/st14/occ2", 17: "TMA R1/st24/occ2", 100: "SYM upper-triangle TMA R4/st6/occ2 (+fold)"}
if len(sys.argv) > 2:
    names = {int(v): names[int(v)] for v in sys.argv[2].split(",")}
with gx.DeviceMap(n) as dm:
    dm.synth(seed=110)
    for v in names:
        ms = min(dm.bench_matvec(20, v) for _ in range(3))
        print(json.dumps({"n": n, "variant": v, "name": names[v], "us": round(ms * 1e3, 1), "dense_equiv_GBps": round(4.0 * n * n / ms / 1e6, 1)}))
