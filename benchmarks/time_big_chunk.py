#!/usr/bin/env python
"""Times one 2^26-site chunk (the unit of the whole-genome walk) through the asynchronous entry point and prints the
kernel split the library reports (DNV_F_TIME_KERNELS).  Environment knobs of the library apply."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denovo-variant-caller-java_b200", "python"))
import torch  # noqa: E402
import denovo_b200 as D  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
stream = torch.cuda.Stream()
h = torch.zeros(2, dtype=torch.int64).pin_memory()
with D.TrioEngine(1e-2, 1e-8, 1.0, "BAYES", time_kernels=True) as eng:
    eng.set_stream(stream.cuda_stream)
    buf = torch.empty((n, 12), dtype=torch.int16, device="cuda")
    cap = n // 50
    calls = torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(stream):
        eng.synth_counts_device(0x5EED0004, 0, n, buf, depth=30)
        for _ in range(2):
            eng.infer_device_async(buf, n, calls, cap, h.data_ptr(), site_base=0)
            eng.flush()
        eng.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            eng.infer_device_async(buf, n, calls, cap, h.data_ptr(), site_base=0)
            eng.flush()
        e1.record(stream)
        eng.sync()
        ms = e0.elapsed_time(e1) / 5
        k = eng.infer_device(buf, n, calls, cap)
        st = eng.stats()
print("chunk %d sites: %.3f ms async (%.1f G sites/s); sync call: screen %.3f ms, kernels %.3f ms, undecided %d, calls %d  [%s]" % (
    n, ms, n / ms / 1e6, st["screen_kernel_ms"], st["fused_kernel_ms"], st["n_sites"] - st["n_screened_out"], k,
    " ".join("%s=%s" % (a, os.environ[a]) for a in sorted(os.environ) if a.startswith("DNV_"))))
