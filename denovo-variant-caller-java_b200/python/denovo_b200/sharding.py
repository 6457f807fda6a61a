"""Interval sharding of sites across ranks and the final gather of call records.

The path shards by independent units: every site's inference
depends only on its own 12 counts, the reference already runs one task per site
(M/ReadCaller.java:91-95) and one VariantsBuffer per position stride
(M/VariantCaller.java:92-102).  Rank r owns the r-th contiguous range of the
concatenated genomic axis, so rank order = genomic order, and there is NO
data-path collective: the only exchange is the final gather of call records and
counters to rank 0, which replaces the shared PrintWriter + synchronized
writeCalls (M/DenovoCaller.java:56-59).  The product gather is below the C ABI (dnv_gather: the
kernels store their records straight into rank 0's HBM; denovo_b200.Gather); `gather_records` here is the
collective fallback for launchers without peer access, on any torch.distributed backend (NCCL with CUDA
tensors, gloo with CPU tensors in the tests).
"""
import numpy as np


def shard_range(n_total, world, rank):
    """Contiguous [lo, hi) of rank `rank`; the first n_total % world ranks get one extra unit."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def stride_shards(start, end, num_threads):
    """The reference's own per-thread position shards of one contig (closed ranges),
    M/VariantCaller.java:89-98: stride = (end - start) / numThreads, last thread to the contig end."""
    stride = (end - start) // num_threads
    out = []
    for i in range(num_threads):
        s = start + i * stride
        e = end if i == num_threads - 1 else s + stride - 1
        out.append((s, e))
    return out


def gather_records(local, dist=None, dst=0):
    """Gather variable-length record arrays (torch uint8 [k, itemsize], CPU or CUDA) to `dst`
    in rank order.  Two collectives and ONE host sync: an all-gather of the per-rank counts,
    then a gather padded to the largest count.  Returns (records on dst | None, per-rank counts)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local, [int(local.shape[0])]
    world, rank = dist.get_world_size(), dist.get_rank()
    cnt = torch.full((1,), local.shape[0], dtype=torch.int64, device=local.device)
    allc = torch.empty((world,), dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(allc, cnt)
    counts = allc.tolist()
    mx = max(max(counts), 1)
    if local.shape[0] == mx:
        send = local.contiguous()
    else:
        send = torch.empty((mx, local.shape[1]), dtype=local.dtype, device=local.device)
        send[:local.shape[0]] = local
    recv = list(torch.empty((world, mx, local.shape[1]), dtype=local.dtype, device=local.device).unbind(0)) \
        if rank == dst else None
    dist.gather(send, recv, dst=dst)
    if rank != dst:
        return None, counts
    return torch.cat([recv[r][:counts[r]] for r in range(world)], dim=0), counts


def reduce_counters(values, dist=None, device="cpu"):
    """Sum a list of integer counters over ranks (sites, calls, hard-path hits)."""
    import torch
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return [int(x) for x in t.tolist()]


def globalize_sites(records_np, site_offset):
    """Shift shard-relative site indices of CALL_DT records to global indices."""
    out = records_np.copy()
    out["site"] += np.int64(site_offset)
    return out
