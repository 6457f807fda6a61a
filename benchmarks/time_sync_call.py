#!/usr/bin/env python
"""Host-clock time of one synchronous dnv_infer_counts_device call (10 M sites resident in HBM, raw pointers), the
figure bench.py reports as config.ms_per_step_synchronous_api.   usage: time_sync_call.py [sites] [calls]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denovo-variant-caller-java_b200", "python"))
import torch  # noqa: E402
import denovo_b200 as D  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
with D.TrioEngine() as eng:
    counts = torch.empty((n, 12), dtype=torch.int16, device="cuda")
    eng.synth_counts_device(0x5EED0003, 0, n, counts)
    cap = max(n // 100, 1024)
    calls = torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    cp, cl = counts.data_ptr(), calls.data_ptr()
    for _ in range(5):
        k = eng.infer_device(cp, n, cl, cap)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(reps):
            k = eng.infer_device(cp, n, cl, cap)
        best = min(best, (time.perf_counter() - t0) / reps)
    st = eng.stats()
    print("sync call: %.1f us per call of %d sites (%d calls, %d undecided) lanes=%s host_collect=%s min_chunk=%s" % (
        best * 1e6, n, k, st["n_sites"] - st["n_screened_out"], os.environ.get("DNV_LANES", "default"),
        "off" if os.environ.get("DNV_NO_HOST_COLLECT") else "on", os.environ.get("DNV_SYNC_MIN_CHUNK", "default")))
