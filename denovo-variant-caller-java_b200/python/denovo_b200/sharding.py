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
    """Gather of call records to `dst` WITHOUT a data collective: every rank's kernels store their
    dnv_call records and the record count straight into rank dst's HBM over NVLink (peer stores
    through a torch symmetric-memory mapping).  The destination holds `slots` generations of
    world x [64-byte header (int64 count) | cap records]: batch k of a pipelined run writes
    generation k % slots, so ranks never wait for each other per batch — ONE signal-pad barrier per
    group of <= slots batches (`barrier()`) tells dst that everything has landed.
    Rank order = genomic order, so the regions of one generation read in rank order are the
    gathered result.  `gather_records` (NCCL) is the fallback when symmetric memory is unavailable."""

    HEADER = 64

    def __init__(self, dist, cap_records, itemsize=80, dst=0, slots=1):
        import torch
        import torch.distributed._symmetric_memory as symm_mem
        self.torch, self.dist, self.dst = torch, dist, dst
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.cap, self.itemsize, self.slots = int(cap_records), int(itemsize), int(slots)
        self.slot = (self.HEADER + self.cap * self.itemsize + 255) // 256 * 256
        self.total = self.slots * self.world * self.slot
        self.buf = symm_mem.empty((self.total,), dtype=torch.uint8, device="cuda")
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD)
        self.dst_view = self.hdl.get_buffer(dst, (self.total,), torch.uint8)  # rank dst's buffer
        self.dst_base = int(self.hdl.buffer_ptrs[dst])

    def _off(self, rank, gen):
        return ((gen % self.slots) * self.world + rank) * self.slot

    def records_ptr(self, gen=0):
        """Device pointer (in rank dst's memory) where this rank's kernels write their dnv_call records."""
        return self.dst_base + self._off(self.rank, gen) + self.HEADER

    def count_ptr(self, gen=0):
        """Device pointer (in rank dst's memory) of this rank's int64 record count."""
        return self.dst_base + self._off(self.rank, gen)

    def barrier(self):
        """Signal-pad barrier on the current stream: after it, every rank's stores have landed in dst."""
        self.hdl.barrier()

    def publish(self, n_calls, gen=0):
        """Host-known count: write it next to the records, then barrier (synchronous callers)."""
        lo = self._off(self.rank, gen)
        self.dst_view[lo:lo + 8].view(self.torch.int64).fill_(int(n_calls))
        self.hdl.barrier()

    def counts(self, gen=0):
        """On dst (after barrier): per-rank record counts of one generation."""
        lo = self._off(0, gen)
        heads = self.buf[lo:lo + self.world * self.slot].view(self.world, self.slot)[:, :8].contiguous()
        return heads.view(self.torch.int64).view(-1).tolist()

    def records(self, counts=None, gen=0):
        """On dst: the gathered records of one generation as one contiguous uint8 [total, itemsize] tensor."""
        counts = self.counts(gen) if counts is None else counts
        lo = self._off(0, gen)
        rows = self.buf[lo:lo + self.world * self.slot].view(self.world, self.slot)
        rows = rows[:, self.HEADER:self.HEADER + self.cap * self.itemsize].reshape(self.world, self.cap, self.itemsize)
        return self.torch.cat([rows[r, :counts[r]] for r in range(self.world)], dim=0)
