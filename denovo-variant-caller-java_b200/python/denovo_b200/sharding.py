"""Interval sharding of sites across ranks and the final gather of call records.

The path shards by independent units: every site's inference
depends only on its own 12 counts, the reference already runs one task per site
(M/ReadCaller.java:91-95) and one VariantsBuffer per position stride
(M/VariantCaller.java:92-102).  Rank r owns the r-th contiguous range of the
concatenated genomic axis, so rank order = genomic order, and there is NO
data-path collective: the only exchange is the final gather of call records and
counters to rank 0, which replaces the shared PrintWriter + synchronized
writeCalls (M/DenovoCaller.java:56-59).  Works on any torch.distributed backend
(NCCL with CUDA tensors on the GPU box, gloo with CPU tensors in the tests).
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


class PeerGather:
    """Gather of call records to `dst` WITHOUT a data collective: every rank's compaction kernel
    stores its dnv_call records straight into rank dst's HBM over NVLink (peer stores through a
    torch symmetric-memory mapping), then publishes its count and joins a signal-pad barrier.
    Layout of the destination buffer: world slots of [64-byte header (int64 count) | cap records].
    Rank order = genomic order, so the slots read in order are the gathered result.
    Falls back to `gather_records` (NCCL) when symmetric memory is unavailable."""

    HEADER = 64

    def __init__(self, dist, cap_records, itemsize=80, dst=0):
        import torch
        import torch.distributed._symmetric_memory as symm_mem
        self.torch, self.dist, self.dst = torch, dist, dst
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.cap, self.itemsize = int(cap_records), int(itemsize)
        self.slot = (self.HEADER + self.cap * self.itemsize + 255) // 256 * 256
        self.buf = symm_mem.empty((self.world * self.slot,), dtype=torch.uint8, device="cuda")
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD)
        self.dst_view = self.hdl.get_buffer(dst, (self.world * self.slot,), torch.uint8)  # rank dst's buffer
        lo = self.rank * self.slot
        self.my_header = self.dst_view[lo:lo + 8].view(torch.int64)
        self.my_records_ptr = int(self.hdl.buffer_ptrs[dst]) + lo + self.HEADER

    def records_ptr(self):
        """Device pointer (in rank dst's memory) where this rank's kernels write their dnv_call records."""
        return self.my_records_ptr

    def count_ptr(self):
        """Device pointer (in rank dst's memory) of this rank's int64 record count."""
        return int(self.hdl.buffer_ptrs[self.dst]) + self.rank * self.slot

    def barrier(self):
        """Signal-pad barrier on the current stream: after it, every rank's stores have landed in dst."""
        self.hdl.barrier()

    def publish(self, n_calls):
        """Call after the producing kernels were enqueued on the current stream."""
        self.my_header.fill_(int(n_calls))
        self.hdl.barrier()

    def counts(self):
        """On dst (after publish): per-rank record counts."""
        heads = self.buf.view(self.world, self.slot)[:, :8].contiguous().view(self.torch.int64)
        return heads.view(-1).tolist()

    def records(self, counts=None):
        """On dst: the gathered records as one contiguous uint8 [total, itemsize] tensor."""
        counts = self.counts() if counts is None else counts
        rows = self.buf.view(self.world, self.slot)[:, self.HEADER:self.HEADER + self.cap * self.itemsize]
        rows = rows.reshape(self.world, self.cap, self.itemsize)
        return self.torch.cat([rows[r, :counts[r]] for r in range(self.world)], dim=0)

