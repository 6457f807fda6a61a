#!/usr/bin/env python
"""Per-chunk times of the whole-genome walk's inner loop (bench.py genome_walk, one GPU): generator, then one
asynchronous batch + flush per 2^26-site chunk, synchronised per chunk."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denovo-variant-caller-java_b200", "python"))
import torch  # noqa: E402
import denovo_b200 as D  # noqa: E402

chunk = 1 << 26
stream = torch.cuda.Stream()
h = torch.zeros(2, dtype=torch.int64).pin_memory()
with D.TrioEngine(1e-2, 1e-8, 1.0, "BAYES") as eng:
    eng.set_stream(stream.cuda_stream)
    bufs = [torch.empty((chunk, 12), dtype=torch.int16, device="cuda") for _ in range(2)]
    cap = chunk // 50
    calls = [torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda") for _ in range(2)]
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    out = []
    with torch.cuda.stream(stream):
        for k, first in enumerate([0, chunk, 2 * chunk, 3 * chunk, 40 * chunk, 41 * chunk, 42 * chunk, 46 * chunk - 12345]):
            m = chunk if k < 7 else 8_000_000
            b = k % 2
            evs[0].record(stream)
            eng.synth_counts_device(0x5EED0004, first, m, bufs[b], depth=30)
            evs[1].record(stream)
            evs[3].record(stream)
            eng.infer_device_async(bufs[b], m, calls[b], cap, h.data_ptr() + 8 * b, site_base=first)
            eng.flush()
            evs[2].record(stream)
            eng.sync()
            out.append((first, m, evs[0].elapsed_time(evs[1]), evs[3].elapsed_time(evs[2]), int(h[b])))
for o in out:
    print("first %12d n %9d: generator %.2f ms, inference %.3f ms, calls %d" % o)
