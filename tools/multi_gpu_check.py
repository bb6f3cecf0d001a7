"""Run under torchrun (one rank per GPU): row-sharded IC / KR / VC / MCFS vs a single-GPU run of the same
device-generated map on rank 0.  Prints one JSON line; exit code 1 on any mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as tdist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcmapexplorer_b200 import device as gx      # noqa: E402
from gcmapexplorer_b200 import dist as gd        # noqa: E402


def mark(msg):
    if os.environ.get("GCX_TRACE"):
        print(f"[rank {os.environ.get('RANK')}] {msg}", file=sys.stderr, flush=True)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6001
    partition = sys.argv[2] if len(sys.argv) > 2 else "rows"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dm = gd.sharded_device_map(n, device=local, partition=partition, symmetric=True)
    dm.synth(seed=5, scale=200.0, alpha=1.0, empty_frac=0.01)
    mark('created+synth'); rs, zc = dm.row_stats()
    keep = rs != 0
    res = {}
    mark('row_stats')
    dm.set_mask(np.ones(n, dtype=bool))
    ic = dm.ic(1e-4, 500)
    mark(f'ic done sym={ic["symmetric"]} passes={ic["passes"]}')
    mn_ic, mx_ic = dm.apply_sym(None, masked_mode=1, in_place=False)
    cs_ic = dm.checksum(1)
    dm.set_mask(keep)
    mark('apply ic'); kr = dm.kr(1e-12, 0.1, 3.0)
    mark('kr done')
    mn_kr, mx_kr = dm.apply_sym(None, masked_mode=0, in_place=False)
    cs_kr = dm.checksum(1)
    mn_vc, mx_vc, csum = dm.vc(False, in_place=False)
    cs_vc = dm.checksum(1)
    mark('vc done'); avg_mean = dm.diag_stats("mean")
    avg_med = dm.diag_stats("median")
    mark('diag done')
    mn_tp, mx_tp = dm.transition(in_place=False)                 # lib/statDist.py: transition matrix, then its stationary vector
    cs_tp = dm.checksum(1)
    dm.swap()
    sd = dm.stationary()
    dm.swap()
    mark('statdist done')
    # all ranks hold identical vectors / counters
    def same_everywhere(arr):
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).copy()).cuda()
        ref = t.clone()
        tdist.broadcast(ref, 0)
        ok = torch.tensor([int(torch.equal(t, ref))], device="cuda")
        tdist.all_reduce(ok, op=tdist.ReduceOp.MIN)
        return bool(ok.item())
    res["ranks_identical"] = all(same_everywhere(a) for a in (ic["bias"], kr["x"], rs, zc, csum, avg_mean, avg_med, sd["x"]))
    # checksums of the sharded outputs add up
    cs = torch.tensor([cs_ic & 0x7fffffffffffffff, cs_kr & 0x7fffffffffffffff, cs_vc & 0x7fffffffffffffff, cs_tp & 0x7fffffffffffffff],
                      dtype=torch.int64, device="cuda")
    hi = torch.tensor([cs_ic >> 63, cs_kr >> 63, cs_vc >> 63, cs_tp >> 63], dtype=torch.int64, device="cuda")
    tdist.all_reduce(cs)
    tdist.all_reduce(hi)
    tot = [((int(c) + (int(h) << 63)) & 0xffffffffffffffff) for c, h in zip(cs.tolist(), hi.tolist())]
    ok = True
    if rank == 0:
        with gx.DeviceMap(n, device=local) as one:
            one.synth(seed=5, scale=200.0, alpha=1.0, empty_frac=0.01)
            rs1, zc1 = one.row_stats()
            one.set_mask(np.ones(n, dtype=bool))
            ic1 = one.ic(1e-4, 500)
            m1 = one.apply_sym(None, masked_mode=1, in_place=False)
            c_ic = one.checksum(1)
            one.set_mask(rs1 != 0)
            kr1 = one.kr(1e-12, 0.1, 3.0)
            m2 = one.apply_sym(None, masked_mode=0, in_place=False)
            c_kr = one.checksum(1)
            m3 = one.vc(False, in_place=False)
            c_vc = one.checksum(1)
            a_mean = one.diag_stats("mean")
            a_med = one.diag_stats("median")
            m4 = one.transition(in_place=False)
            c_tp = one.checksum(1)
            one.swap()
            sd1 = one.stationary()
        res.update(partition=partition, sharded_symmetric_pass=bool(ic["symmetric"]), single_symmetric_pass=bool(ic1["symmetric"]),
            peer_exchange=bool(ic.get("peer_exchange")), products_launched=(ic.get("products_launched"), ic1.get("products_launched")),
            rowstats_equal=bool(np.array_equal(rs, rs1) and np.array_equal(zc, zc1)),
            ic_passes=(ic["passes"], ic1["passes"]), ic_bias_rel=float(np.max(np.abs(ic["bias"] / ic1["bias"] - 1))),
            kr_counters=((kr["outer"], kr["MVP"]), (kr1["outer"], kr1["MVP"])),
            kr_x_rel=float(np.max(np.abs(kr["x"][keep] / kr1["x"][keep] - 1))),
            minmax_ic=((mn_ic, mx_ic), m1), minmax_kr=((mn_kr, mx_kr), m2), minmax_vc=((mn_vc, mx_vc), m3[:2]),
            checksum_vc_equal=bool(tot[2] == c_vc), checksum_ic_equal=bool(tot[0] == c_ic), checksum_kr_equal=bool(tot[1] == c_kr),
            median_equal=bool(np.array_equal(avg_med, a_med)), mean_rel=float(np.max(np.abs(avg_mean - a_mean) / np.maximum(np.abs(a_mean), 1e-300))),
            checksum_tp_equal=bool(tot[3] == c_tp), minmax_tp=((mn_tp, mx_tp), m4),
            sd_iterations=(sd["iterations"], sd1["iterations"]), sd_rel=float(np.max(np.abs(sd["x"][keep] / sd1["x"][keep] - 1))),
            ic_ms=(ic["ms"], ic1["ms"]), kr_ms=(kr["ms"], kr1["ms"]), sd_ms=(sd["ms"], sd1["ms"]))
        ok = (res["ranks_identical"] and res["rowstats_equal"] and res["ic_passes"][0] == res["ic_passes"][1]
              and res["ic_bias_rel"] < 1e-12 and res["kr_counters"][0] == res["kr_counters"][1] and res["kr_x_rel"] < 1e-11
              and res["checksum_vc_equal"] and res["median_equal"] and res["mean_rel"] < 1e-13
              and np.float32(mn_vc) == np.float32(m3[0]) and np.float32(mx_vc) == np.float32(m3[1])
              and abs(mn_ic / m1[0] - 1) < 1e-6 and abs(mx_kr / m2[1] - 1) < 1e-6
              and res["checksum_tp_equal"] and (mn_tp, mx_tp) == tuple(m4) and abs(sd["iterations"] - sd1["iterations"]) <= 1
              and res["sd_rel"] < 1e-11)
        res["ok"] = bool(ok)
        print(json.dumps(res, default=str))
    flag = torch.tensor([int(ok)], device="cuda")
    tdist.broadcast(flag, 0)
    dm.close()
    tdist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
