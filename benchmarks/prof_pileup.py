#!/usr/bin/env python
"""Smallest driver of the pileup / fused reads->calls kernels for ncu: 1M consecutive sites at 30x."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denovo-variant-caller-java_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import denovo_b200 as D  # noqa: E402
import synth_inputs as S  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rs, bo, ba, pe = S.make_reads(41, 1, n, depth=30)
pos = np.arange(1, n + 1, dtype=np.int64)
eng = D.TrioEngine()
d = [torch.from_numpy(x).cuda() for x in (rs, bo, ba, pe, pos)]
counts = torch.zeros((n, 12), dtype=torch.int16, device="cuda")
calls = torch.zeros((n // 20, 80), dtype=torch.uint8, device="cuda")
for _ in range(3):
    eng.pileup_device(d[0], d[1], d[2], d[3], len(rs), d[4], n, counts)
    k = eng.infer_reads_device(d[0], d[1], d[2], d[3], len(rs), d[4], n, calls, n // 20)
torch.cuda.synchronize()
print("prof_pileup: %d sites, %d reads, %d calls" % (n, len(rs), k))
