"""torchrun worker (N >= 2 GPUs): interval-sharded inference with the peer-store gather must equal
single-GPU inference over the whole range, record for record, in genomic order.
Run:  python -m torch.distributed.run --nproc-per-node 2 tests/multi_gpu_check.py"""
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
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    total, seed = 3_000_001, 0x5EED0003
    lo, hi = sharding.shard_range(total, world, rank)
    n = hi - lo
    cap = 20000
    eng = D.TrioEngine(device=local)
    counts = torch.empty((n, 12), dtype=torch.int16, device="cuda")
    eng.synth_counts_device(seed, lo, n, counts)
    for use_peer in ("async", True, False):
        if use_peer == "async":
            # pipelined batches: three generations in flight, the kernels store records AND counts into
            # rank 0's HBM, one barrier for the group; every generation must hold the same gathered result
            pg = sharding.PeerGather(dist, cap, D.CALL_DT.itemsize, slots=3)
            torch.cuda.set_stream(torch.cuda.Stream())
            eng.set_stream(torch.cuda.current_stream().cuda_stream)
            for g in range(3):
                eng.infer_device_async(counts, n, pg.records_ptr(g), cap, pg.count_ptr(g))
            pg.barrier()
            eng.sync()
            torch.cuda.synchronize()
            if rank == 0:
                gens = [pg.records(gen=g).cpu().numpy().tobytes() for g in range(3)]
                assert gens[0] == gens[1] == gens[2], "generations differ"
            got = pg.records(gen=2).cpu().numpy().reshape(-1).view(D.CALL_DT) if rank == 0 else None
            offs = np.cumsum([0] + pg.counts(gen=2))[:-1] if rank == 0 else None
            eng.set_stream(None)
            torch.cuda.set_stream(torch.cuda.default_stream())
        elif use_peer:
            pg = sharding.PeerGather(dist, cap, D.CALL_DT.itemsize)
            nc = eng.infer_device(counts, n, pg.records_ptr(), cap)
            pg.publish(nc)
            torch.cuda.synchronize()
            got = pg.records().cpu().numpy().reshape(-1).view(D.CALL_DT) if rank == 0 else None
            offs = np.cumsum([0] + pg.counts())[:-1] if rank == 0 else None
        else:
            local_calls = torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
            nc = eng.infer_device(counts, n, local_calls, cap)
            rec, cnts = sharding.gather_records(local_calls[:nc], dist)
            got = rec.cpu().numpy().reshape(-1).view(D.CALL_DT) if rank == 0 else None
            offs = np.cumsum([0] + cnts)[:-1] if rank == 0 else None
        if rank == 0:
            got = got.copy()
            bounds = [sharding.shard_range(total, world, r)[0] for r in range(world)]
            ends = list(offs[1:]) + [len(got)]
            for r in range(world):  # shard-relative -> global site index
                got["site"][offs[r]:ends[r]] += bounds[r]
            whole = torch.empty((total, 12), dtype=torch.int16, device="cuda")
            eng.synth_counts_device(seed, 0, total, whole)
            ref_calls = torch.zeros((4 * cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
            nref = eng.infer_device(whole, total, ref_calls, 4 * cap)
            want = ref_calls[:nref].cpu().numpy().reshape(-1).view(D.CALL_DT)
            assert len(got) == len(want) and got.tobytes() == want.tobytes(), ("mismatch", use_peer, len(got), len(want))
            print("multi_gpu_check OK: world %d, %s gather, %d calls" % (world, {"async": "pipelined peer-store", True: "peer-store", False: "NCCL"}[use_peer], len(got)))
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
