"""Host-side plumbing for row-block sharding over several B200s (one process per GPU).

The data path stays in libgcxnorm.so (NCCL allgather of the n-vector after every matrix pass, on the
context's stream); this module only (a) defines the partition, (b) moves the 128-byte ncclUniqueId between
ranks with torch.distributed, (c) builds the per-rank DeviceMap.
"""
import numpy as np


def rows_per_rank(n, world):
    """ceil(n / world): every rank's allgather slot has this many entries (the last rank's tail is padding)."""
    return (int(n) + int(world) - 1) // int(world)


def row_partition(n, world, rank):
    """(row0, nloc) of the contiguous row block owned by `rank`."""
    rpr = rows_per_rank(n, world)
    row0 = min(int(n), rank * rpr)
    return row0, max(0, min(rpr, int(n) - row0))


def row_partition_aligned(n, world, rank):
    """(row0, nloc) of the library's aligned partition: near-equal row blocks with cuts on multiples of 1024 rows
    (gcx_partition_rows).  With symmetric=True a sharded map then runs the upper-triangle pass with the off-diagonal
    blocks dealt out in a circulant pattern -- rows, PCIe traffic and symmetric work all balanced."""
    import ctypes as C
    from ._lib import check, lib
    r0, nl = C.c_int64(0), C.c_int64(0)
    check(lib().gcx_partition_rows(int(n), int(world), int(rank), C.byref(r0), C.byref(nl)))
    return r0.value, nl.value


def row_partition_equal_area(n, world, rank, align=1024):
    """(row0, nloc) so that every rank owns the same share of the UPPER TRIANGLE: row i has n-i entries right of the
    diagonal, so the cut after a fraction f of the area is at n (1 - sqrt(1 - f)).  Cuts are rounded to `align` rows
    (the 1024-column panels of the symmetric pass).  Used with the options exchange_allreduce=1, assume_symmetric=1."""
    def cut(r):
        if r >= world:
            return int(n)
        x = n * (1.0 - (1.0 - r / world) ** 0.5)
        return min(int(n), int(round(x / align)) * align)
    row0, row1 = cut(rank), cut(rank + 1)
    return row0, max(0, row1 - row0)


def padded_length(n, world):
    return rows_per_rank(n, world) * int(world)


def allgather_padded(local, n, world, rank, all_gather_fn):
    """Reference semantics of the in-library exchange: `local` holds this rank's nloc results; returns the full
    n-vector assembled from equal-size padded slots.  all_gather_fn(list_of_out_arrays, in_array) is the
    collective (torch.distributed.all_gather on CPU tensors in the gloo tests)."""
    rpr = rows_per_rank(n, world)
    slot = np.zeros(rpr, dtype=local.dtype)
    slot[: local.shape[0]] = local
    outs = [np.zeros(rpr, dtype=local.dtype) for _ in range(world)]
    all_gather_fn(outs, slot)
    return np.concatenate(outs)[:n]


def broadcast_unique_id(rank, device=None):
    """ncclUniqueId created on rank 0 and broadcast with torch.distributed (any initialised backend)."""
    import torch
    import torch.distributed as tdist
    from .device import dist_unique_id
    backend = tdist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t = torch.tensor(list(dist_unique_id()), dtype=torch.uint8, device=dev)
    tdist.broadcast(t, 0)
    return bytes(t.cpu().tolist())


def sharded_device_map(n, device=None, partition="rows", symmetric=False, assume_symmetric=False):
    """DeviceMap for this rank's row block, wired to the other ranks (torch.distributed must be initialised).

    partition="rows": equal ceil(n/world) row blocks, n-vectors exchanged by allgather, dense passes.
    partition="balanced" (+ symmetric=True): near-equal row blocks cut on multiples of 1024 rows; IC/KR passes read each
    unordered pair of bins once, the off-diagonal blocks dealt out to the ranks in a circulant pattern; allreduce exchange.
    partition="equal_area": row blocks holding equal shares of the upper triangle, n-vectors exchanged by allreduce, IC/KR
    passes read the upper triangle only.
    Whether the sharded map is symmetric is tested across the ranks by the library (k_sym_hash); `symmetric` is kept for
    callers of round 1 and ignored, assume_symmetric=True skips the test (the caller vouches)."""
    import torch.distributed as tdist
    from .device import DeviceMap
    rank, world = tdist.get_rank(), tdist.get_world_size()
    options = {}
    if partition == "balanced" and world > 1:
        row0, nloc = row_partition_aligned(n, world, rank)
        options = {"exchange_allreduce": 1, "assume_symmetric": 1 if assume_symmetric else 0}
    elif partition == "equal_area" and world > 1:
        row0, nloc = row_partition_equal_area(n, world, rank)
        options = {"exchange_allreduce": 1, "assume_symmetric": 1 if assume_symmetric else 0, "sym_blocks": 0}
    else:
        row0, nloc = row_partition(n, world, rank)
    if nloc == 0:
        raise ValueError("more ranks than row blocks")
    dist = (rank, world, broadcast_unique_id(rank)) if world > 1 else None
    return DeviceMap(n, row0, nloc, device=device, dist=dist, options=options)
