#!/usr/bin/env python
"""bench.py -- the driver-facing benchmark of the normalization hot path.

  python bench.py --gpus N --steps K --warmup W                  # this repo's sm_100a path
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path, same metric

Workload (BASELINE.json): N=1 -> configs[1], synthetic chr1 @ 10 kb (n = 24,926, 2.485 GB fp32 dense), IC + VC.
N>1 -> the same pipeline on a row-sharded map with the per-GPU block held at 2.485 GB (n = 24,926*sqrt(N):
weak scaling), one NCCL allgather of the n-vector per IC pass.  `--workload chr1_1kb` runs BASELINE
configs[4] (n = 249,251, needs >= 2 GPUs), `--workload chr1_5kb_kr` runs configs[2] (KR on one GPU).

A "step" = one complete pass of the hot path over the resident map:
    row statistics/mask (4n^2 B) + IC to convergence (matvecs x 4n^2 B) + final apply (8n^2 B) + VC apply (8n^2 B).
metric  = algorithmic bytes of the step / step time, in GB/s (whole job, all ranks) -- BASELINE's "per-iteration HBM
          GB/s"; the time-to-converge BASELINE also names is carried in `time_to_converge_s` / `ms_per_step`.
e2e     = the same metric for the C-ABI call sequence on HOST buffers: pinned-host upload + mask + IC + apply +
          download + VC + download, all copies inside the timed region.
roofline= the dominant kernel (k_matvec): 4*nloc*n bytes per launch / mean CUDA-event duration of the launches inside
          the timed region, against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline = the reference's own compiled Cython cores (oracle/_ref, kind "reference") -- or the NumPy port when
          they are absent -- timed on this box's host cores on a bounded sample of the same map.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_CHR1_10KB = 24926     # hg19 chr1 length 249,250,621 // 10,000 + 1  (lib/importer.py:1695-1707 shape rule)
N_CHR1_5KB = 49851
N_CHR1_1KB = 249251
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md fallback
IC_TOL, IC_ITER = 1e-4, 500    # API defaults, lib/normalizer.py:505
REF_IC_CAP_PASSES = IC_ITER + 2


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def recorded_traffic(kernel, n):
    """DRAM bytes per launch of `kernel` at size n from this round's ncu --set full capture (profiles/r01_traffic.json), else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            return json.load(f).get("%s@%d" % (kernel, n), {}).get("bytes")
    except Exception:
        return None


def step_bytes(n, nloc_total, matvecs):
    """Algorithmic HBM bytes of one step: mask 4, IC 4/matvec, apply 8, VC apply 8 -- times rows x n."""
    return (4.0 * (1 + matvecs) + 16.0) * nloc_total * n


# ======================================================================================================
# this repo's arm
# ======================================================================================================
def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    dist = None
    tdist = None
    if world > 1:
        import torch
        import torch.distributed as tdist
        torch.cuda.set_device(local)
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gcmapexplorer_b200 import device as gx

    if args.workload == "chr1_1kb":
        n = args.n or N_CHR1_1KB
        wl = "synthetic chr1 @ 1 kb (n=%d), IC + VC, row-sharded" % n
    elif args.workload == "chr1_5kb_kr":
        n = args.n or N_CHR1_5KB
        wl = "synthetic chr1 @ 5 kb (n=%d), KR" % n
    else:
        n = args.n or int(round(N_CHR1_10KB * math.sqrt(world)))
        wl = ("synthetic chr1 @ 10 kb (n=%d), IC + VC" % n) if world == 1 else \
             ("synthetic chr1-like map, n=%d (2.485 GB fp32 per GPU), IC + VC, row-sharded" % n)
    from gcmapexplorer_b200 import dist as gd
    rpr = (n + world - 1) // world
    options = {}
    if world > 1 and args.partition == "balanced":
        # near-equal row blocks (cuts on 1024 rows); each unordered pair of bins is read once, the off-diagonal blocks dealt out
        # to the ranks in a circulant pattern.  The synthetic map is symmetric by construction, which a rank cannot verify
        # locally -> assume_symmetric.
        row0, nloc = gd.row_partition_aligned(n, world, rank)
        options = {"exchange_allreduce": 1, "assume_symmetric": 1}
    elif world > 1 and args.partition == "equal_area":
        # equal shares of the upper triangle per rank (panel-aligned cuts): the IC/KR passes read the upper triangle only.
        # The synthetic map is symmetric by construction, which a rank cannot verify locally -> assume_symmetric.
        row0, nloc = gd.row_partition_equal_area(n, world, rank)
        options = {"exchange_allreduce": 1, "assume_symmetric": 1, "sym_blocks": 0}
    else:
        row0, nloc = gd.row_partition(n, world, rank)
    if world > 1:
        import torch
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(gx.dist_unique_id()), dtype=torch.uint8, device="cuda")
        tdist.broadcast(idt, 0)
        dist = (rank, world, bytes(idt.cpu().tolist()))
    dm = gx.DeviceMap(n, row0, nloc, device=local, dist=dist, options=options)
    dm.synth(seed=args.seed, scale=200.0, alpha=1.0, empty_frac=0.01)
    peak, peak_kind = measured_peak()
    kr_mode = args.workload == "chr1_5kb_kr" or args.kr

    def barrier():
        dm.sync()
        if tdist is not None:
            tdist.barrier()

    def max_over_ranks(x):
        if tdist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if tdist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        tdist.all_reduce(t, op=tdist.ReduceOp.SUM)
        return float(t.item())

    ones = np.ones(n, dtype=bool)
    info = {}
    # the out-of-place result buffer doubles the footprint; with the equal-area partition the last rank holds ~35 % of the
    # rows, so at chr1 @ 1 kb (88 GB there) it does not fit -> such runs time mask + solver only and say so
    free_b = dm.info()["hbm_free"]
    fits = min_over_ranks_flag(tdist, free_b > 4.0 * nloc * (n + 32) * 1.02 + (2 << 30))
    solver_only = not fits

    def step():
        rs, zc = dm.row_stats(force=True)                 # K1: mask pass
        keep = rs != 0.0
        if kr_mode:
            dm.set_mask(keep)
            r = dm.kr(1e-12, 0.1, 3.0)
            if not solver_only:
                dm.apply_sym(None, masked_mode=0, in_place=False)
            info.update(matvecs=r["MVP"] + 1, outer=r["outer"], converge_ms=r["ms"], passes=r["MVP"], pass_bytes=r["pass_bytes"], symmetric=r["symmetric"])
        else:
            dm.set_mask(ones)                             # IC without filter runs on the whole map (normalizer.py:592)
            r = dm.ic(IC_TOL, IC_ITER)                    # K2 x passes + vector epilogues
            if not solver_only:
                dm.apply_sym(None, masked_mode=1, in_place=False)   # K4
                dm.set_mask(keep)
                dm.vc(False, in_place=False)              # K5 (row sums reused from the mask pass)
            info.update(matvecs=r["matvecs"], passes=r["passes"], converge_ms=r["ms"], l1=r["l1"], pass_bytes=r["pass_bytes"], symmetric=r["symmetric"])

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()          # sampled through warm-up + timed region (both keep the GPU under the same load)
    for _ in range(args.warmup):
        step()
    # ---- timed region: device-resident ---------------------------------------------------------------
    dm.profile(True)
    dm.profile_read(reset=True)
    l0 = dm.launch_count()
    barrier()
    dm.timer_start()
    conv_ms = []
    for _ in range(args.steps):
        step()
        conv_ms.append(info["converge_ms"])
    ms_total = dm.timer_stop()
    barrier()
    ms_total = max_over_ranks(ms_total)
    prof = dm.profile_read(reset=True)
    dm.profile(False)
    launches = dm.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    mv = info["matvecs"]
    # bytes the step actually has to move: mask 4n^2 + matvecs x (bytes of one pass over the resident layout:
    # 4n^2 dense, ~2n(n+1024) when the upper-triangle pass is used) + apply 8n^2 (+ VC apply 8n^2)
    pass_bytes_all = sum_over_ranks(float(info["pass_bytes"]))
    ew = 0.0 if solver_only else (8.0 if kr_mode else 16.0)
    bytes_step = 4.0 * n * n + mv * pass_bytes_all + ew * n * n
    dense_equiv = (4.0 * (1 + mv) + ew) * n * n
    value = bytes_step / (ms_per_step * 1e-3) / 1e9

    # roofline of the dominant kernel: mean duration over the *effective* launches (launches after the
    # device-side convergence flag exit immediately; their few microseconds stay in the sum)
    mv_ms = max_over_ranks(prof["matvec"]["ms"]) / max(1, args.steps * mv)
    mv_bytes = max_over_ranks(float(info["pass_bytes"]))      # the most loaded rank's share of one pass
    achieved = mv_bytes / (mv_ms * 1e-3) / 1e9
    kname = ("k_matvec_sym_tma+k_sym_fold" if kr_mode else "k_ic_pass_sym_tma") if info["symmetric"] else ("k_matvec_tma" if kr_mode else "k_ic_pass_tma")
    roofline = {"kernel": kname, "layout": "dense fp32, upper triangle read (symmetric pass)" if info["symmetric"] else "dense fp32, all rows read", "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": recorded_traffic(kname, n) if world == 1 else None, "bytes_per_launch": mv_bytes, "us_per_launch": round(mv_ms * 1e3, 2),
                "launches_timed": args.steps * mv,
                "other_kernels_ms_per_step": {k: round(v["ms"] / args.steps, 4) for k, v in prof.items() if k != "matvec" and v["launches"]}}

    # ---- e2e: the same pipeline through the C ABI with HOST (pinned) buffers ------------------------------
    e2e = None
    if not args.no_e2e and not solver_only and nloc * n * 8 < 0.4 * psutil_avail():
        hin = gx.PinnedArray((nloc, n))
        hout = gx.PinnedArray((nloc, n))
        dm.download(0, out=hin.array)

        hout2 = None if kr_mode else gx.PinnedArray((nloc, n))

        def e2e_step():
            dm.upload(hin.array)                           # H2D 4*nloc*n
            rs, zc = dm.row_stats(force=True)
            keep = rs != 0.0
            if kr_mode:
                dm.set_mask(keep)
                dm.kr(1e-12, 0.1, 3.0)
                dm.apply_sym(None, masked_mode=0, in_place=False)
                dm.download(1, out=hout.array)             # D2H 4*nloc*n
                return 1
            dm.set_mask(keep)
            dm.vc(False, in_place=False)                   # VC first: its result streams out ...
            dm.download_async(1, hout2.array)              # ... D2H on the copy stream while IC iterates
            dm.set_mask(ones)
            dm.ic(IC_TOL, IC_ITER)
            dm.apply_sym(None, masked_mode=1, in_place=True)    # the uploaded copy is not needed any more
            dm.download_wait()
            dm.download(0, out=hout.array)                 # D2H: IC result
            return 2
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        ke = max(1, min(args.steps, 5))
        for _ in range(ke):
            nd = e2e_step()
        dm.sync()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / ke)
        barrier()
        e2e = {"value": round(dense_equiv / e2e_s / 1e9, 2), "unit": "GB/s", "value_basis": "dense-equivalent algorithmic bytes (same basis as the reference arm)", "seconds_per_step": round(e2e_s, 4),
               "h2d_bytes_per_step": int(4 * nloc * n), "d2h_bytes_per_step": int(4 * nloc * n * nd),
               "path": "gcx_map_upload(pinned) -> row_stats -> vc_run -> gcx_map_download_async || ic_run -> apply(in place) -> gcx_map_download" if not kr_mode else "gcx_map_upload(pinned) -> row_stats -> kr_run -> apply -> gcx_map_download"}
        host_map = hin.array if (rank == 0 and world == 1) else None
    else:
        host_map = None

    # ---- CPU baseline beside it (rank 0, N=1 only) -------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and host_map is not None:
        if kr_mode:
            cpu = cpu_reference_sample_kr(host_map, mv, bytes_step)
        else:
            cpu = cpu_reference_sample(host_map, mv - 1, budget_s=args.cpu_budget)

    if rank == 0:
        line = {
            "metric": "IC+VC normalization throughput, chr1 10 kb (algorithmic HBM GB/s over the whole step; time-to-converge alongside)"
                      if not kr_mode else "KR normalization throughput (algorithmic HBM GB/s over the whole step)",
            "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 accumulate / f32 storage", "data": "synthetic (device-generated Poisson power-law map, 1% empty bins)",
            "config": {"workload": wl, "n": n, "rows_this_rank": nloc, "map_bytes_per_gpu_mean": int(4 * rpr * n), "partition": (args.partition if world > 1 else "single"),
                       "step": "mask + solver only (out-of-place result buffer does not fit on the largest rank)" if solver_only else "mask + solver + apply (+ VC)", "l2": "inputs larger than L2 (no flush needed)",
                       "tol": IC_TOL if not kr_mode else 1e-12, "parallelism": ("row-block x%d (%s), one NCCL %s of the n-vector per pass" % (world, args.partition, "allreduce" if info["symmetric"] else "allgather")) if world > 1 else "single GPU"},
            "time_to_converge_s": round(float(np.median(conv_ms)) * 1e-3, 5), "passes": info["passes"], "matvecs": mv,
            "hbm_bytes_per_step": bytes_step, "dense_equivalent_bytes_per_step": dense_equiv,
            "value_dense_equivalent": round(dense_equiv / (ms_per_step * 1e-3) / 1e9, 1),
            "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": int(launches), "clocks": clk,
        }
        print(json.dumps(line))
    dm.close()
    if tdist is not None:
        tdist.destroy_process_group()


WG_25KB_BINS = [9971, 9728, 7921, 7647, 7237, 6845, 6366, 5855, 5649, 5422, 5401, 5355, 4607, 4294, 4102, 3615, 3248,
                3124, 2366, 2522, 1926, 2053, 6211, 2375]     # hg19 chr1..22, X, Y at 25 kb


def run_mcfs_wg(args):
    """BASELINE configs[3]: distance-frequency (MCFS o/e) over the 24 whole-genome maps at 25 kb on one GPU.
    Maps are device-generated and resident; a step = per-diagonal stats + apply for all 24 maps."""
    from gcmapexplorer_b200 import device as gx
    stats = args.mcfs_stats
    maps = []
    for k, n in enumerate(WG_25KB_BINS):
        dm = gx.DeviceMap(n, device=int(os.environ.get("LOCAL_RANK", "0")))
        dm.synth(seed=25000 + k, scale=200.0, alpha=1.0, empty_frac=0.01)
        maps.append(dm)
    peak, peak_kind = measured_peak()

    def one(dm):
        dm.diag_stats(stats)
        dm.diag_apply(None, subtract=False, in_place=False)

    pool = None
    if args.mcfs_threads > 1:
        # the 24 maps are independent and every DeviceMap has its own context + stream; ctypes releases the GIL, so a few host
        # threads keep several maps' (launch-latency-bound) kernels in flight at once
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=args.mcfs_threads)

    def step():
        if pool is None:
            for dm in maps:
                one(dm)
        else:
            list(pool.map(one, maps))
    for _ in range(args.warmup):
        step()
    for dm in maps:
        dm.profile(True); dm.profile_read(True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    for dm in maps:
        dm.sync()
    sec = (time.perf_counter() - t0) / args.steps
    red_ms = app_ms = 0.0
    launches = 0
    for dm in maps:
        p = dm.profile_read(True)
        red_ms += p["diag_reduce"]["ms"] / args.steps
        app_ms += p["diag_apply"]["ms"] / args.steps
    nn = float(sum(n * n for n in WG_25KB_BINS))
    traversals = 5 if stats == "median" else 1
    bytes_step = (2.0 * traversals + 8.0) * nn
    line = {"metric": "MCFS (distance-frequency o/e, stats=%s) throughput, whole genome @ 25 kb, 24 maps (algorithmic GB/s)" % stats,
            "value": round(bytes_step / sec / 1e9, 1), "unit": "GB/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 stats / f32 storage",
            "data": "synthetic (device-generated)", "config": {"workload": "24 maps, sum n^2*4 = %.3f GB" % (4 * nn / 1e9), "stats": stats, "host_threads": args.mcfs_threads},
            "kernel_ms_per_step": {"diag_reduce": round(red_ms, 3), "diag_apply": round(app_ms, 3)},
            "roofline": {"kernel": "k_diag_apply", "bound": "hbm", "achieved": round(8.0 * nn / (app_ms * 1e-3) / 1e9, 1), "peak": peak, "peak_kind": peak_kind,
                         "unit": "GB/s", "frac": round(8.0 * nn / (app_ms * 1e-3) / 1e9 / peak, 4), "traffic": None}}
    if not args.no_cpu_baseline:
        cpu = cpu_reference_sample_mcfs(stats, nn)
        cpu["value"] = round(bytes_step / cpu["seconds_per_step"] / 1e9, 4)
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    for dm in maps:
        dm.close()


def min_over_ranks_flag(tdist, flag):
    if tdist is None:
        return bool(flag)
    import torch
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device="cuda")
    tdist.all_reduce(t, op=tdist.ReduceOp.MIN)
    return bool(t.item())


def run_chr21(args):
    """BASELINE configs[0]: chr21 @ 100 kb (n = 482) -- the one size where the reference runs as-is to convergence on the CPU.
    Three-way report: the reference's compiled IC core (fp32 in place), the fp64 restatement, and this library, same map."""
    from gcmapexplorer_b200 import device as gx
    from oracle import oracle_np as onp
    n = args.n or 482
    A = onp.synth_map(n, seed=21100, scale=2000.0)
    kind, ic, vc, mask = _load_reference()
    t0 = time.perf_counter()
    M = A.copy()
    ic(M, IC_TOL, IC_ITER)
    t_ref = time.perf_counter() - t0
    passes_ref = None
    try:
        from oracle import ref_loader as rl
        rl.ic_pass_counter(reset=True)
        M2 = A.copy()
        ic(M2, IC_TOL, IC_ITER)
        passes_ref = rl.ic_pass_counter()
    except Exception:
        pass
    M64, B64, passes64, _ = onp.ic_fp64(A, IC_TOL, IC_ITER)
    with gx.DeviceMap.from_host(A) as dm:
        dm.set_mask(np.ones(n, dtype=bool))
        for _ in range(args.warmup):
            r = dm.ic(IC_TOL, IC_ITER)
        ts = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            dm.upload(A)
            dm.set_mask(np.ones(n, dtype=bool))
            r = dm.ic(IC_TOL, IC_ITER)
            dm.apply_sym(None, masked_mode=1, in_place=True)
            out = dm.download(0)
            ts.append(time.perf_counter() - t0)
    nz = M64 != 0
    line = {"metric": "IC time-to-converge, chr21 @ 100 kb (s, whole call incl. upload/download)", "value": round(float(np.median(ts)), 6), "unit": "s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(float(np.median(ts)) * 1e3, 3), "higher_is_better": False,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 accumulate / f32 storage", "data": "synthetic (host-generated, n=%d)" % n,
            "config": {"workload": "synthetic chr21 @ 100 kb (n=%d), IC" % n},
            "passes": {"gpu": r["passes"], "fp64_restatement": passes64, "reference_as_is": passes_ref}, "device_ms": r["ms"],
            "cpu_baseline": {"value": round(t_ref, 4), "unit": "s", "cores": 1, "kind": kind, "sample": "full run of the reference's IC core, as-is"},
            "agreement": {"gpu_vs_fp64_max_rel": float(np.max(np.abs(out[nz] - M64[nz].astype(np.float32)) / M64[nz])),
                          "gpu_vs_reference_max_rel": float(np.max(np.abs(out[nz] - M[nz]) / M64[nz])),
                          "reference_vs_fp64_max_rel": float(np.max(np.abs(M[nz] - M64[nz]) / M64[nz]))}}
    print(json.dumps(line))


def run_statdist(args):
    """SURVEY.md 8f row N3 at chr1 @ 10 kb: transition probability matrix of the synthetic map (lib/statDist.py:50-99) and the
    stationary distribution of that matrix (lib/statDist.py:436-498) -- the reference's dense eigendecomposition is replaced by a
    power iteration with one read of the resident matrix per step (k_matvec_t)."""
    from gcmapexplorer_b200 import device as gx
    n = args.n or 24926
    peak, peak_src = measured_peak()
    dm = gx.DeviceMap(n)
    dm.synth(seed=args.seed, scale=200.0, alpha=1.0, empty_frac=0.01)
    rs, zc = dm.row_stats()
    keep = rs != 0
    dm.set_mask(keep)

    def step():
        dm.invalidate()                              # the row-sum pass is part of the transition-matrix call
        dm.timer_start()
        dm.transition(in_place=False)
        t_tp = dm.timer_stop()
        dm.swap()                                    # the transition matrix becomes the input of the power iteration
        dm.timer_start()
        r = dm.stationary()
        t_sd = dm.timer_stop()
        dm.swap()
        return t_tp, t_sd, r
    for _ in range(args.warmup):
        step()
    clocks = ClockSampler(0)
    clocks.start()
    dm.profile(True)
    dm.profile_read(reset=True)
    t0 = time.perf_counter()
    parts = [step() for _ in range(args.steps)]
    wall = (time.perf_counter() - t0) / args.steps
    prof = dm.profile_read(reset=True)
    dm.profile(False)
    ck = clocks.stop()
    r = parts[-1][2]
    t_tp = float(np.median([p[0] for p in parts]))
    t_sd = float(np.median([p[1] for p in parts]))
    mv = prof["matvec"]
    # launches enqueued ahead of the convergence check exit at once (a few us): divide by the products that did run
    products = sum(p[2]["iterations"] for p in parts)
    us = mv["ms"] / max(1, products) * 1e3
    achieved = 4.0 * n * n / (us * 1e-6) / 1e9
    line = {"metric": "stationary distribution time-to-converge (s) and per-product HBM GB/s, chr1 @ 10 kb transition matrix",
            "value": round((t_tp + t_sd) * 1e-3, 6), "unit": "s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(wall * 1e3, 3), "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 accumulate / f32 storage", "data": "synthetic (device-generated)",
            "config": {"workload": "synthetic chr1 @ 10 kb (n=%d): transition matrix + stationary distribution" % n, "n": n,
                       "l2": "inputs larger than L2 (2.49 GB per pass)"},
            "transition_ms": round(t_tp, 3), "transition_gbs": round(16.0 * n * n / (t_tp * 1e-3) / 1e9, 1),   # row sums 4 + minimum pass 4 + output pass 8 bytes per entry
            "stationary_ms": round(t_sd, 3), "iterations": r["iterations"], "last_delta": r["delta"],
            "roofline": {"bound": "hbm", "kernel": "k_matvec_t", "achieved": round(achieved, 1), "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": recorded_traffic("k_matvec_t", n), "us_per_launch": round(us, 2),
                         "algorithmic_bytes_per_launch": int(4 * n * n)},
            "gpu_launches": int(sum(v["launches"] for v in prof.values())), "clocks": ck}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_statdist_sample(n, args.cpu_budget)
    print(json.dumps(line))


def cpu_statdist_sample(n, budget_s):
    """The reference's own arithmetic (NumPy row loop + scipy.linalg.eig, restated in oracle/oracle_np.py and checked against
    the reference's outputs in tests/golden/sd_*) on a bounded sample: the transition matrix on a block of rows scaled by the
    row count, the dense eigendecomposition at a small size scaled by (n/m)^3."""
    from oracle import oracle_np as onp
    m = 1200
    A = onp.synth_map(m, seed=5, scale=200.0)
    b = ~onp.get_nonzeros_index(A)
    t0 = time.perf_counter()
    P, _ = onp.transition_matrix(A, b)
    t_tp = time.perf_counter() - t0
    t0 = time.perf_counter()
    onp.stationary_eig(P, b)
    t_eig = time.perf_counter() - t0
    est = t_tp * (n / m) ** 2 + t_eig * (n / m) ** 3
    return {"value": round(est, 1), "unit": "s", "cores": os.cpu_count(), "kind": "port",
            "sample": "n=%d: transition matrix %.3f s (scaled by (n/m)^2), scipy.linalg.eig %.2f s (scaled by (n/m)^3; LAPACK threads as configured); "
                      "extrapolated -- a 24,926^2 float64 eigendecomposition does not finish in a bench run" % (m, t_tp, t_eig)}


def run_corr(args):
    """SURVEY.md 8f row N4 at chr1 @ 10 kb: pairwise-masked Pearson correlation matrix of the synthetic map (lib/corrMatrix.py:92-135,
    lib/corrMatrixCoreSRC.c).  The pair sums are GEMM-shaped library products (cuBLAS): on the integer-valued synthetic counts all of them
    run exactly on the int8 / bf16 tensor cores; with --corr-float the map is scaled by 0.37 first (a normalised, non-integer map) and the
    two value products run on the float64 pipe (DGEMM + DSYRK).  The bound is the tensor rate, not HBM."""
    from gcmapexplorer_b200 import device as gx
    n = args.n or 24926
    dm = gx.DeviceMap(n)
    dm.synth(seed=args.seed, scale=200.0, alpha=1.0, empty_frac=0.01)
    if args.corr_float:
        dm.scale(0.37)
    rs, zc = dm.row_stats()
    keep = rs != 0
    m = int(keep.sum())
    dm.set_mask(keep)
    for _ in range(max(1, args.warmup)):
        dm.corr(0.0)
    clocks = ClockSampler(0)
    clocks.start()
    ts = []
    l0 = dm.launch_count()
    for _ in range(args.steps):
        dm.timer_start()
        dm.corr(0.0)
        ts.append(dm.timer_stop())
    launches = dm.launch_count() - l0
    ck = clocks.stop()
    t = float(np.median(ts)) * 1e-3
    variant = int(os.environ.get("GCX_CORR_VARIANT", "1"))
    mb = (m + 63) // 64 * 64
    slices, sxy_sliced = dm.corr_last_route()
    K = abs(slices)
    if variant != 1:
        kind, ops, peak, unit = "cublasDgemm x3 (N = I I^T, Sx = X0 I^T, Sxy = X0 X0^T)", 6.0 * m ** 3, 40.0, "TFLOP/s"
    elif slices == 0:
        kind, ops, peak, unit = "cublasGemmEx bf16 (N = I I^T) + cublasDgemm (Sx = X0 I^T) + cublasDsyrk (Sxy = X0 X0^T)", 3.0 * m ** 3, 40.0, "TFLOP/s"
    else:
        # exact 7-bit slices of X0 (integer map: plain digits; float32-valued map: one exponent per row): K products for Sx and, when Sxy is
        # sliced too, K (K + 1) / 2 for Sxy (pairs a <= b) -- all int8 x int8 -> int32 -- plus the bf16 counts product
        nprod = K + (K * (K + 1) // 2 if sxy_sliced else 0)
        kind = "cublasGemmEx int8 x%d (%d seven-bit slices of X0: Sx%s) + bf16 x1 (N = I I^T)%s" % (
            nprod, K, ", Sxy" if sxy_sliced else "", "" if sxy_sliced else " + cublasDsyrk (Sxy)")
        ops, peak, unit = (nprod + 1) * 2.0 * mb ** 3, 4500.0, "TOP/s"
    line = {"metric": "correlation matrix time (s), chr1 @ 10 kb", "value": round(t, 4), "unit": "s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(1, args.warmup), "ms_per_step": round(t * 1e3, 2), "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if (variant != 1 or slices == 0) else "int8/int32 exact slice products, f64 accumulation and epilogue",
            "data": "synthetic (device-generated)%s" % (", scaled by 0.37" if args.corr_float else ""),
            "config": {"workload": "synthetic chr1 @ 10 kb (n=%d, %d bins with data): correlation matrix" % (n, m), "n": n, "kept": m,
                       "map_values": "non-integer" if args.corr_float else "integer counts"},
            "roofline": {"bound": "tensor", "kernel": kind, "achieved": round(ops / t / 1e12, 2), "peak": peak,
                         "peak_source": "nominal data-sheet rate (no measured fp64 / int8 peak on file)", "unit": unit,
                         "frac": round(ops / t / 1e12 / peak, 4), "traffic": None,
                         "note": "achieved = operations of the products over the WHOLE call (prep, slicing, products, accumulation, variances, scatter): a lower bound for the GEMMs"},
            "gpu_launches": int(launches), "clocks": ck}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_corr_sample(n, m)
    print(json.dumps(line))


def cpu_corr_sample(n, m):
    """The reference's C core (OpenMP, all host threads), compiled as it lies into oracle/_ref, on an 1200-bin map; scaled by (m/1200)^3."""
    from oracle import oracle_np as onp
    from oracle import ref_loader as rl
    ms = 1200
    A = onp.synth_map(ms, seed=5, scale=200.0)
    keep = onp.get_nonzeros_index(A)
    X = A[keep][:, keep].T.astype(np.float64, order="C")
    if rl.have_corr():
        kind, fn = "reference", (lambda: rl.load_corr().calculateCorrMatrix(X, maskvalue=0.0))
    else:
        kind, fn = "port", (lambda: onp.corr_matrix(X[:300, :300], maskvalue=0.0))
    t0 = time.perf_counter()
    fn()
    dt = time.perf_counter() - t0
    mk = X.shape[0] if kind == "reference" else 300
    return {"value": round(dt * (m / mk) ** 3, 1), "unit": "s", "cores": os.cpu_count(), "kind": kind,
            "sample": "%d bins: %.2f s, scaled by (m/%d)^3 -- extrapolated" % (mk, dt, mk)}


def psutil_avail():
    try:
        import psutil
        return psutil.virtual_memory().available
    except Exception:
        return 64 << 30


# ======================================================================================================
# CPU reference: bounded sample of the same workload
# ======================================================================================================
def _load_reference():
    from oracle import ref_loader as rl
    if rl.have_cores():
        core, _, cmh = rl.load_cores()
        return "reference", core.performIterativeCorrection, core.performVCNormalization, cmh.get_nonzeros_index
    from oracle import oracle_np as onp          # NumPy port of the same lines (kind "port")

    def ic(M, tol, it):
        onp._ic_recurrence(M, tol, it, np.float32)

    def vc(A, Out=None, sqroot=False, bNoData=None):
        onp.vc(A, bNoData=bNoData, sqroot=sqroot, out=Out)
    return "port", ic, vc, onp.get_nonzeros_index


def cpu_reference_sample(host_map, passes_same_work, budget_s=25.0):
    """Times the reference's IC core for exactly 2 passes (iteration=0 -> lib/normalizeCore.pyx:55-57 breaks at
    the second pass) on the full map, and its VC + mask on a leading principal block, then scales:
    IC by passes, VC/mask by (n/n_s)^2."""
    kind, ic, vc, mask = _load_reference()
    n = host_map.shape[0]
    threads = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    M = np.array(host_map, dtype=np.float32, copy=True)
    t0 = time.perf_counter()
    ic(M, IC_TOL, 0)
    t_pass = (time.perf_counter() - t0) / 2.0
    del M
    ns = min(n, 3000)
    sub = np.ascontiguousarray(host_map[:ns, :ns])
    t0 = time.perf_counter()
    keep = mask(sub)
    t_mask = (time.perf_counter() - t0) * (n / ns) ** 2
    out = sub.copy()
    t0 = time.perf_counter()
    vc(sub, Out=out, sqroot=False, bNoData=~np.asarray(keep, dtype=bool))
    t_vc = (time.perf_counter() - t0) * (n / ns) ** 2
    t_cap = t_mask + REF_IC_CAP_PASSES * t_pass + t_vc
    t_same = t_mask + passes_same_work * t_pass + t_vc
    b_cap = step_bytes(n, n, REF_IC_CAP_PASSES + 1)
    b_same = step_bytes(n, n, passes_same_work + 1)
    return {"value": round(b_same / t_same / 1e9, 4), "unit": "GB/s", "cores": 1, "host_cores_available": os.cpu_count(), "blas_threads": threads,
            "kind": kind, "seconds_per_step_same_passes": round(t_same, 1), "seconds_per_step_at_reference_cap": round(t_cap, 1),
            "value_at_reference_cap": round(b_cap / t_cap / 1e9, 4), "s_per_ic_pass": round(t_pass, 3), "s_mask_scaled": round(t_mask, 2), "s_vc_scaled": round(t_vc, 2),
            "sample": f"IC: 2 passes of the reference core on the full n={n} map (single-threaded NumPy), scaled to {passes_same_work} passes "
                      f"(the fp64-converged count; the as-is fp32 reference stalls above tol and runs {REF_IC_CAP_PASSES} = its cap, SURVEY.md fact 3); "
                      f"mask + VC: reference cores on the leading {ns}x{ns} block, scaled by (n/{ns})^2"}


def cpu_reference_sample_kr(host_map, matvecs, bytes_step):
    """KR on the CPU is np.dot(float32 A, float64 x) per matrix-vector product (lib/normalizeKnightRuiz.pyx:120,139,169),
    which NumPy evaluates by materialising a float64 copy of A and calling OpenBLAS dgemv.  Times 2 products on the full
    map after one warm-up and scales by the product count (lower bound: compaction, vector algebra and
    GenNormMatrixRAM's temp files are left out)."""
    n = host_map.shape[0]
    x = np.ones((n, 1))
    np.dot(host_map[:64], x)
    t0 = time.perf_counter()
    for _ in range(2):
        y = np.dot(host_map, x)
    t_mvp = (time.perf_counter() - t0) / 2
    total = t_mvp * matvecs
    return {"value": round(bytes_step / total / 1e9, 4), "unit": "GB/s", "cores": int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1)),
            "kind": "reference", "s_per_matvec": round(t_mvp, 3), "seconds_per_step_same_matvecs": round(total, 1),
            "sample": f"2 x np.dot(float32 A[{n}x{n}], float64 x) (the reference's matrix-vector product, OpenBLAS threads as available), scaled to {matvecs} "
                      "products; compaction, O(n) algebra and the final matrix generation not included (lower bound)"}


def cpu_reference_sample_mcfs(stats, nn_total):
    """getAvgContactByDistance (lib/cmstats.py:569-605, pure Python: NumPy port in oracle_np) + the reference's compiled
    normalizeByAvgContactByDivision on one 2,366-bin map (chrY @ 25 kb shape), scaled by sum(n^2)/2366^2."""
    from oracle import oracle_np as onp
    from oracle import ref_loader as rl
    n = 2366
    A = onp.synth_map(n, seed=25023)
    t0 = time.perf_counter()
    avg = onp.avg_contact_by_distance(A, stats)
    t_stats = time.perf_counter() - t0
    out = np.zeros_like(A)
    t0 = time.perf_counter()
    if rl.have_cores():
        rl.load_cores()[0].normalizeByAvgContactByDivision(A, avg, Out=out)
        kind = "reference (apply) + port (stats)"
    else:
        out = onp.mcfs_divide(A, avg)
        kind = "port"
    t_apply = time.perf_counter() - t0
    scale = nn_total / float(n * n)
    total = (t_stats + t_apply) * scale
    return {"value": None, "unit": "GB/s", "cores": 1, "kind": kind, "seconds_per_step": round(total, 1),
            "s_stats_one_map": round(t_stats, 2), "s_apply_one_map": round(t_apply, 2),
            "sample": f"one {n}-bin map (stats={stats}) through the reference's per-diagonal Python loops, scaled by sum(n^2)/{n}^2 = {scale:.1f}"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation on this box's host cores, same metric/unit/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_np as onp
    n = args.n or N_CHR1_10KB
    # the same synthetic law as the device generator, drawn with the host RNG on a band-limited budget:
    # the reference's cost per pass does not depend on the values, only on n.
    rng = np.random.default_rng(args.seed)
    t_gen = time.perf_counter()
    A = np.zeros((n, n), dtype=np.float32)
    blk = 2048
    idx = np.arange(n)
    for r0 in range(0, n, blk):
        r1 = min(n, r0 + blk)
        d = np.abs(idx[r0:r1, None] - idx[None, :]).astype(np.float32)
        A[r0:r1] = rng.poisson(200.0 / (d + 1.0)).astype(np.float32)
    A = np.triu(A)
    A = A + np.triu(A, 1).T
    empty = rng.choice(n, size=max(2, round(0.01 * n)), replace=False)
    A[empty, :] = 0
    A[:, empty] = 0
    t_gen = time.perf_counter() - t_gen
    samples = []
    cpu = None
    for i in range(args.warmup + args.steps):
        cpu = cpu_reference_sample(A, args.ref_passes, budget_s=args.cpu_budget)
        if i >= args.warmup:
            samples.append(cpu)
        if sum(s["s_per_ic_pass"] * 2 for s in samples) > 240:     # keep the whole run within a few minutes
            break
    val = float(np.median([s["value"] for s in samples]))
    sec = float(np.median([s["seconds_per_step_same_passes"] for s in samples]))
    cpu = dict(samples[-1], value=round(val, 4))
    line = {"impl": "reference", "metric": "IC+VC normalization throughput, chr1 10 kb (algorithmic HBM GB/s over the whole step; time-to-converge alongside)",
            "value": round(val, 4), "unit": "GB/s", "n_gpus": args.gpus, "steps": len(samples), "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (reference as-is)", "data": "synthetic (host-generated Poisson power-law map)",
            "config": {"workload": "synthetic chr1 @ 10 kb (n=%d), IC + VC" % n, "n": n, "passes_assumed": args.ref_passes},
            "cpu_baseline": cpu,
            "e2e": {"value": round(val, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "chr1_10kb", "chr1_5kb_kr", "chr1_1kb", "wg_25kb_mcfs", "chr21_100kb", "chr1_10kb_statdist", "chr1_10kb_corr"])
    ap.add_argument("--mcfs-stats", default="median", choices=["median", "mean"])
    ap.add_argument("--mcfs-threads", type=int, default=6, help="host threads driving the 24 independent maps of wg_25kb_mcfs")
    ap.add_argument("--n", type=int, default=0, help="override the number of bins (quick checks)")
    ap.add_argument("--seed", type=int, default=110)
    ap.add_argument("--ref-passes", type=int, default=60, help="IC passes the reference sample is scaled to (fp64-converged count)")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--partition", default="balanced", choices=["balanced", "equal_area", "rows"],
                    help="N>1: balanced = near-equal row blocks + symmetric passes with circulant block assignment; equal_area = equal "
                         "upper-triangle shares per rank; rows = equal row blocks + dense passes")
    ap.add_argument("--kr", action="store_true", help="run KR instead of IC + VC on the selected workload")
    ap.add_argument("--corr-float", action="store_true", help="chr1_10kb_corr: scale the map by 0.37 first (non-integer values -> float64 products)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "wg_25kb_mcfs":
        run_mcfs_wg(args)
    elif args.workload == "chr21_100kb":
        run_chr21(args)
    elif args.workload == "chr1_10kb_statdist":
        run_statdist(args)
    elif args.workload == "chr1_10kb_corr":
        run_corr(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
