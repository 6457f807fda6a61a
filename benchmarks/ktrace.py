#!/usr/bin/env python
"""Device-side time line of the READ pipeline (DNV_KTRACE=1): for each of K asynchronous batches the first-CTA
start and last-CTA end (%globaltimer) of the screen, list and finish kernels, relative to the first screen's
start.  Shows what actually overlaps on the GPU.   usage: ktrace.py [batches] [sites]"""
import ctypes as C
import os
import sys

os.environ["DNV_KTRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denovo-variant-caller-java_b200", "python"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import denovo_b200 as D  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 12
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
eng = D.TrioEngine()
counts = torch.empty((n, 12), dtype=torch.int16, device="cuda")
eng.synth_counts_device(0x5EED0003, 0, n, counts)
cap = n // 100
calls = [torch.zeros((cap, 80), dtype=torch.uint8, device="cuda") for _ in range(4)]
h = torch.zeros(4, dtype=torch.int64).pin_memory()
ROUNDS = 4  # the first rounds warm up (allocations, graph capture)
MODE = sys.argv[3] if len(sys.argv) > 3 else "async"
import time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = torch.cuda.Stream()
eng.set_stream(st.cuda_stream)
for rep in range(ROUNDS):
    torch.cuda.synchronize()
    e0.record(st)
    t0 = time.perf_counter()
    for i in range(K):
        if MODE == "sync":   # one synchronous call per batch: chunks over the lanes, graph replay
            eng.infer_device(counts, n, calls[0], cap)
        else:
            eng.infer_device_async(counts, n, calls[i % 4], cap, h[i % 4:i % 4 + 1])
    t1 = time.perf_counter()
    eng.flush()
    e1.record(st)
    eng.sync()
    print("round %d: %.1f us per batch (host clock), host enqueue %.1f us per batch, CUDA events %.1f us per batch"
          % (rep, (time.perf_counter() - t0) / K * 1e6, (t1 - t0) / K * 1e6, e0.elapsed_time(e1) * 1e3 / K))
if MODE == "sync":
    K *= 3
SLOTS = 1024
buf = np.zeros((4, SLOTS, 6), np.uint64)
seq = np.zeros(4, np.uint32)
eng.lib.dnv_debug_ktrace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
lanes = eng.lib.dnv_debug_ktrace(eng._h, buf.ctypes.data_as(C.c_void_p), seq.ctypes.data_as(C.c_void_p))
rows = []
for l in range(lanes):
    for s in range(int(seq[l])):
        rows.append((l,) + tuple(int(x) for x in buf[l, s % SLOTS]))
rows.sort(key=lambda r: r[1])
rows = rows[-K:]
t0 = rows[0][1]
print("lanes=%s screen_ctas=%s graphs=%s (us, relative to the first screen's start of the last round)"
      % (os.environ.get("DNV_LANES", "default"), os.environ.get("DNV_SCREEN_CTAS", "default"), os.environ.get("DNV_GRAPHS", "default")))
print("batch lane  screen[start end]   list[start end]   finish[start end]")
for i, r in enumerate(rows):
    v = [(x - t0) / 1e3 for x in r[1:]]
    print("%3d  %3d  %7.1f %7.1f   %7.1f %7.1f   %7.1f %7.1f" % (i, r[0], *v))
print("steady step: %.1f us; first screen start -> last finish end: %.1f us = %.1f us per batch" % (
    (rows[-1][2] - rows[2][2]) / 1e3 / (K - 3), (max(r[6] for r in rows) - t0) / 1e3, (max(r[6] for r in rows) - t0) / 1e3 / K))
