"""CPU, world_size 2 over gloo: the host-side sharding logic (partition, padded allgather) and the sharded
formulation of IC -- local row-block matvec, allgather of the n-vector, identical vector step on every rank
 -- give the oracle's result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gcmapexplorer_b200 import dist as gd      # noqa: E402
from oracle import oracle_np as onp            # noqa: E402


@pytest.mark.parametrize("n,world", [(10, 1), (10, 2), (11, 2), (24926, 8), (249251, 8), (5, 8)])
def test_row_partition_covers_everything(n, world):
    seen = 0
    for r in range(world):
        row0, nloc = gd.row_partition(n, world, r)
        assert row0 == seen or nloc == 0
        assert nloc <= gd.rows_per_rank(n, world)
        seen += nloc
    assert seen == n and gd.padded_length(n, world) >= n


@pytest.mark.parametrize("n,world", [(24926, 2), (49852, 4), (70501, 8), (249251, 8), (3000, 2)])
def test_equal_area_partition(n, world):
    """Cuts are panel aligned, cover [0, n) and give every rank (nearly) the same share of the upper triangle."""
    cuts = [gd.row_partition_equal_area(n, world, r) for r in range(world)]
    assert cuts[0][0] == 0 and sum(c[1] for c in cuts) == n
    area = []
    for r, (row0, nloc) in enumerate(cuts):
        assert row0 % 1024 == 0 and (r == world - 1 or (row0 + nloc) % 1024 == 0)
        assert r == 0 or row0 == cuts[r - 1][0] + cuts[r - 1][1]
        i = np.arange(row0, row0 + nloc, dtype=np.float64)
        area.append(float((n - i).sum()))
    if n >= 20000:
        assert max(area) / (sum(area) / world) < 1.0 + 2.5 * 1024.0 * world / n     # panel rounding only


def _sharded_ic(rank, world, port, n, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A = onp.synth_map(n, seed=17).astype(np.float64)
        row0, nloc = gd.row_partition(n, world, rank)
        Aloc = A[row0:row0 + nloc]

        def ag(outs, slot):
            touts = [torch.zeros(slot.shape[0], dtype=torch.float64) for _ in outs]
            tdist.all_gather(touts, torch.from_numpy(slot))
            for o, t in zip(outs, touts):
                o[:] = t.numpy()

        # the one-matvec-per-pass recurrence of k_ic_epilogue, with the matvec sharded by rows
        v = np.ones(n)
        B = np.ones(n)
        cc = None
        m = 0
        while True:
            t = gd.allgather_padded(Aloc @ v, n, world, rank, ag)
            p = v * t
            S, C = p.sum(), np.count_nonzero(p)
            if m == 0:
                cc, f, g = S, 1.0, 1.0
            else:
                f, g = np.sqrt(S / cc), cc / S
            mean = g * S / C
            d = g * p
            bn = B if m == 0 else f / v
            l1 = np.abs(B - bn).sum()
            B = bn
            v = np.where(d != 0, mean / (bn * np.where(d != 0, d, 1.0)), 1.0 / bn)
            if m >= 2 and (l1 < 1e-4 or (m - 1) > 500):
                break
            m += 1
        ret[rank] = (m, B)
    finally:
        tdist.destroy_process_group()


def test_sharded_ic_world2_gloo_matches_oracle():
    n, world = 203, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_sharded_ic, args=(world, port, n, ret), nprocs=world, join=True)
    A = onp.synth_map(n, seed=17)
    _, Bo, passes, _ = onp.ic_fp64(A, 1e-4, 500)
    (m0, B0), (m1, B1) = ret[0], ret[1]
    assert m0 == m1 == passes
    assert np.array_equal(B0, B1)                       # ranks hold bitwise identical state
    np.testing.assert_allclose(B0, Bo, rtol=1e-10)


# --------------------------------------------------------------------------------------------------
# circulant block assignment of the sharded upper-triangle pass (mirror of sym_ready() in csrc/gcx_capi.cu)
# --------------------------------------------------------------------------------------------------
def _aligned_cut(n, W, r):
    if r <= 0:
        return 0
    if r >= W:
        return n
    return min(n, (2 * r * n // W + 1024) // 2048 * 1024)


def _panel_table(n, W, me):
    """(row0, nloc, [(bands, full)] per 1024-column panel) for rank `me` -- same rules as the C++ table builder."""
    ld = (n + 31) // 32 * 32
    nch = (ld // 4 + 255) // 256
    r0, r1 = _aligned_cut(n, W, me), _aligned_cut(n, W, me + 1)
    nloc = r1 - r0

    def half_cut(r):
        a, b = _aligned_cut(n, W, r), _aligned_cut(n, W, r + 1)
        return a + ((b - a) // 2 + 512) // 1024 * 1024
    all_bands = (nloc + 3) // 4
    out = []
    for k in range(nch):
        col0, nb, full, s = k * 1024, 0, 0, 0
        while W > 1 and s + 1 < W and _aligned_cut(n, W, s + 1) <= col0:
            s += 1
        if W == 1 or s == me:
            rows_end = min(r1, (k + 1) * 1024 if k < nch - 1 else n)
            nb = (rows_end - r0 + 3) // 4 if rows_end > r0 else 0
            if W > 1 and col0 >= r1:
                nb = 0
        else:
            d = (s - me) % W
            if 2 * d < W:
                nb, full = all_bands, 1
            elif 2 * d == W:
                full = 1
                nb = (half_cut(me) - r0) // 4 if me < s else (all_bands if col0 >= half_cut(s) else 0)
        out.append((nb, full))
    return r0, nloc, out


@pytest.mark.parametrize("n,W", [(9000, 1), (9000, 2), (12345, 2), (13000, 3), (17000, 4), (21000, 5), (30000, 8)])
def test_circulant_blocks_cover_every_pair_once(n, W):
    """Every ordered pair (i, j) -- i.e. every term A_ij z_j of t_i -- is produced by exactly one rank, and the aligned
    partition matches the library's gcx_partition_rows."""
    C = np.zeros((n, n), dtype=np.int8)
    work = []
    for me in range(W):
        r0, nloc, tab = _panel_table(n, W, me)
        assert (r0, nloc) == gd.row_partition_aligned(n, W, me)
        w = 0
        for k, (nb, full) in enumerate(tab):
            if nb == 0:
                continue
            rows = np.arange(r0, min(r0 + nb * 4, r0 + nloc))
            cols = np.arange(k * 1024, min((k + 1) * 1024, n))
            w += len(rows) * len(cols)
            I, J = np.meshgrid(rows, cols, indexing="ij")
            if full:
                C[I, J] += 1
                C[J, I] += 1
            else:
                m = J >= I
                C[I[m], J[m]] += 1
                m = J > I
                C[J[m], I[m]] += 1
        work.append(w)
    assert (C == 1).all()
    assert max(work) / (sum(work) / W) < 1.15
