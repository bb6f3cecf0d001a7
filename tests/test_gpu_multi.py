"""GPU, >= 2 devices: row-sharded run under torchrun vs a single-GPU run (skipped on a 1-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    sys.path.insert(0, ROOT)
    from gcmapexplorer_b200 import _lib
    return _lib.device_count()


@pytest.mark.parametrize("n,partition", [(6001, "rows"), (9000, "equal_area"), (9000, "balanced"), (12345, "balanced")])
def test_sharded_matches_single_gpu(n, partition):
    g = _ngpu()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if g < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + n % 97), os.path.join(ROOT, "tools", "multi_gpu_check.py"), str(n), partition]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]


def _ccmap(A, tmp_path):
    from gcmapexplorer_b200 import ccmap as cmp
    c = cmp.CCMAP()
    c.gen_from_matrix(A, 10000, xlabel="chrS", ylabel="chrS", workDir=str(tmp_path))
    c.make_unreadable()
    return c


def test_dropin_normalizers_shard_over_gpus_in_one_process(tmp_path, monkeypatch):
    """GCX_GPUS=2: normalizeCCMapBy* row-shard the map over two GPUs from this one process (one host thread per GPU, NVLink peer
    exchange inside the IC pass, NCCL for the rest) and give the single-GPU result: same mask, same counters, matrices equal to
    an fp32 ulp, VC / MCFS bit-identical."""
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    import numpy as np
    from oracle import oracle_np as onp
    from gcmapexplorer_b200 import normalizer as nz
    A = onp.synth_map(5200, seed=41)
    res = {}
    for gpus in ("1", "2"):
        monkeypatch.setenv("GCX_GPUS", gpus)
        c = _ccmap(A, tmp_path)
        out = {}
        for name, fn, kw in (("ic", nz.normalizeCCMapByIC, {}), ("icf", nz.normalizeCCMapByIC, {"percentile_threshold_no_data": 95}),
                             ("kr", nz.normalizeCCMapByKR, {}), ("vc", nz.normalizeCCMapByVCNorm, {}), ("mcfs", nz.normalizeCCMapByMCFS, {})):
            o = fn(c, **kw)
            assert o is not None
            o.make_readable()
            out[name] = (np.array(o.matrix), np.array(o.bNoData), float(o.minvalue), float(o.maxvalue), dict(nz.last_run_info))
        res[gpus] = out
    for name in res["1"]:
        m1, b1, mn1, mx1, i1 = res["1"][name]
        m2, b2, mn2, mx2, i2 = res["2"][name]
        assert np.array_equal(b1, b2), name
        for key in ("passes", "outer", "MVP"):
            if key in i1:
                assert i1[key] == i2[key], (name, key)
        if name in ("vc", "mcfs"):
            assert np.array_equal(m1, m2), name
            assert (mn1, mx1) == (mn2, mx2)
        else:
            nzm = m1 != 0
            assert np.array_equal(nzm, m2 != 0)
            assert np.max(np.abs(m1[nzm] - m2[nzm]) / np.abs(m1[nzm])) < 2.5e-7, name
            assert abs(mn1 / mn2 - 1) < 2.5e-7 and abs(mx1 / mx2 - 1) < 2.5e-7


def test_sharded_map_tests_symmetry_instead_of_trusting_the_caller():
    """A row-sharded map decides by a cross-rank hash (k_sym_hash) whether the upper-triangle passes apply: a symmetric map takes
    them, a map with ONE asymmetric entry takes the dense passes and gives exactly what a single GPU's dense passes give."""
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    import numpy as np
    from oracle import oracle_np as onp
    from gcmapexplorer_b200 import device as gx
    from gcmapexplorer_b200.sharded import ShardedDeviceMap
    n = 4500
    A = onp.synth_map(n, seed=43)
    ones = np.ones(n, dtype=bool)
    with ShardedDeviceMap.from_host(A, devices=[0, 1]) as sm:
        sm.set_mask(ones)
        r = sm.ic(1e-4, 500)
        assert r["symmetric"] and r["peer_exchange"] and r["ranks_identical"]
    with gx.DeviceMap.from_host(A) as dm:
        dm.set_mask(ones)
        r1 = dm.ic(1e-4, 500)
    assert r["passes"] == r1["passes"]
    np.testing.assert_allclose(r["bias"], r1["bias"], rtol=1e-12)
    B = A.copy()
    B[17, 3900] += 2.0                    # rank 0's rows, rank 1's columns: neither rank can see the mismatch alone
    with ShardedDeviceMap.from_host(B, devices=[0, 1]) as sm:
        sm.set_mask(ones)
        r = sm.ic(1e-4, 500)
        assert not r["symmetric"] and r["ranks_identical"]
        rs, _ = sm.row_stats()
        sm.set_mask(rs != 0)
        k = sm.kr(1e-12, 0.1, 3.0)
        assert not k["symmetric"]
    with gx.DeviceMap(n, options={"sym": 0}) as dm:
        dm.upload(B)
        dm.set_mask(ones)
        r1 = dm.ic(1e-4, 500)
        dm.set_mask(rs != 0)
        k1 = dm.kr(1e-12, 0.1, 3.0)
    assert r["passes"] == r1["passes"] and (k["outer"], k["MVP"]) == (k1["outer"], k1["MVP"])
    np.testing.assert_allclose(r["bias"], r1["bias"], rtol=1e-12)
    np.testing.assert_allclose(k["x"][rs != 0], k1["x"][rs != 0], rtol=1e-11)
