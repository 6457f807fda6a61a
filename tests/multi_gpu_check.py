"""Worker of the multi-rank gather check (launched with torch.distributed.run, world >= 2): interval-sharded
inference with the gather of include/denovo_b200.h (dnv_gather: the kernels store records and counts straight
into rank 0's HBM; the 64-byte IPC handle of that buffer is the only thing the launcher broadcasts) must equal
single-context inference over the whole range, record for record, in genomic order — for several generations
in flight and for the NCCL / gloo collective fallback.

  real GPUs :  python -m torch.distributed.run --nproc-per-node 2 tests/multi_gpu_check.py
  one GPU   :  DNV_TEST_SAME_DEVICE=1 ... (every rank uses cuda:0, gloo carries the handle): the inter-process path
               (cudaIpcOpenMemHandle, release flags, the destination's wait) is the same."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denovo-variant-caller-java_b200", "python"))
import denovo_b200 as D  # noqa: E402
from denovo_b200 import sharding  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    same = os.environ.get("DNV_TEST_SAME_DEVICE") == "1"
    dev = 0 if same else local
    torch.cuda.set_device(dev)
    if same:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    total, seed = 3_000_001, 0x5EED0003
    lo, hi = sharding.shard_range(total, world, rank)
    n = hi - lo
    cap = 20000
    eng = D.TrioEngine(device=dev)
    counts = torch.empty((n, 12), dtype=torch.int16, device="cuda")
    eng.synth_counts_device(seed, lo, n, counts)
    torch.cuda.synchronize()
    want = None
    if rank == 0:
        whole = torch.empty((total, 12), dtype=torch.int16, device="cuda")
        eng.synth_counts_device(seed, 0, total, whole)
        ref_calls = torch.zeros((4 * cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
        nref = eng.infer_device(whole, total, ref_calls, 4 * cap)
        want = ref_calls[:nref].cpu().numpy().reshape(-1).view(D.CALL_DT).copy()
        del whole
    # ---- the C-ABI gather: three generations in flight, two rounds (the slots are reused)
    g = D.Gather(eng, world, rank, cap, gens=3)
    g.exchange(dist)
    for rnd in range(2):
        for gen in range(3):
            g.submit(counts, n, lo, gen)
        g.arrive()
        if rank == 0:
            g.wait()
            for gen in range(3):
                got, per = g.fetch(gen)
                assert int(per.sum()) == len(got) and len(per) == world
                assert len(got) == len(want) and got.tobytes() == want.tobytes(), ("gather mismatch", rnd, gen, len(got), len(want))
        eng.sync()
        dist.barrier()  # the slots may be rewritten once rank 0 has fetched them
        if rank == 0:
            print("multi_gpu_check OK: world %d, peer-store gather round %d, %d calls" % (world, rnd, len(want)))
    g.close()
    # ---- collective fallback (NCCL on GPUs, gloo on CPU tensors)
    local_calls = torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
    nc = eng.infer_device(counts, n, local_calls, cap)
    send = local_calls[:nc] if not same else local_calls[:nc].cpu()
    rec, cnts = sharding.gather_records(send, dist)
    if rank == 0:
        got = rec.cpu().numpy().reshape(-1).view(D.CALL_DT).copy()
        offs = np.cumsum([0] + cnts)
        for r in range(world):  # shard-relative -> global site index
            got["site"][offs[r]:offs[r + 1]] += sharding.shard_range(total, world, r)[0]
        assert got.tobytes() == want.tobytes(), "collective gather mismatch"
        print("multi_gpu_check OK: world %d, collective gather, %d calls" % (world, len(got)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
