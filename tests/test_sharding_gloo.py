"""world_size-2 (and 3) gloo tests of the N>1 path on CPU: interval sharding + the
gather of call records to rank 0 reproduces the single-rank result, in genomic order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import denovo_b200 as D
from denovo_b200 import sharding
import oracle_lib as O

N_TOTAL = 6001
SEED = 0x5EED0001


def fake_engine_calls(lo, hi):
    """Stand-in for TrioEngine.infer on a CPU-only box: the ORACLE decides, records are laid
    out exactly like dnv_call (the engine itself cannot run without a GPU)."""
    counts = O.synth_counts(SEED, lo, hi - lo)
    ref = O.infer(O.params(), counts)
    sel = np.nonzero(ref["denovo_bayes"])[0]
    rec = np.zeros(len(sel), D.CALL_DT)
    rec["site"] = sel
    rec["counts"] = counts[sel].reshape(-1, 12)
    for f in ("dad", "mom", "child", "p", "lr", "max_ll", "log_men", "log_den"):
        rec[f] = ref[f][sel]
    rec["is_denovo"] = 1
    return rec


def worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(N_TOTAL, world, rank)
    rec = sharding.globalize_sites(fake_engine_calls(lo, hi), lo)
    t = torch.from_numpy(rec.view(np.uint8).reshape(-1, D.CALL_DT.itemsize).copy())
    got, counts = sharding.gather_records(t, dist)
    tot = sharding.reduce_counters([hi - lo, len(rec)], dist)
    if rank == 0:
        np.save(out_path, got.numpy())
        assert tot[0] == N_TOTAL and tot[1] == sum(counts) == got.shape[0]
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gather_equals_single_rank(world, tmp_path):
    out = str(tmp_path / "gathered.npy")
    mp.spawn(worker, args=(world, free_port(), out), nprocs=world, join=True)
    got = np.load(out).reshape(-1).view(D.CALL_DT)
    want = fake_engine_calls(0, N_TOTAL)
    assert got.tobytes() == want.tobytes()
    assert np.all(np.diff(got["site"]) > 0)


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 10_000_000, 3_095_693_981):
        for world in (1, 2, 3, 4, 8):
            r = [sharding.shard_range(n, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_reference_stride_shards():
    """M/VariantCaller.java:89-98 with start 1, end 1000, 4 threads."""
    assert sharding.stride_shards(1, 1000, 4) == [(1, 249), (250, 498), (499, 747), (748, 1000)]
    assert sharding.stride_shards(1, 10, 1) == [(1, 10)]
