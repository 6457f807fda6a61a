#!/usr/bin/env python
"""Smallest driver of the READ pipeline for ncu: a few 10 M-site batches on ONE serial context (DNV_LANES=1, every
kernel alone on the GPU), BAYES, 30x.   usage: prof_read.py [batches] [sites] [method]"""
import os
import sys

os.environ.setdefault("DNV_LANES", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denovo-variant-caller-java_b200", "python"))
import torch  # noqa: E402
import denovo_b200 as D  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
method = sys.argv[3] if len(sys.argv) > 3 else "BAYES"
eng = D.TrioEngine(method=method)
counts = torch.empty((n, 12), dtype=torch.int16, device="cuda")
eng.synth_counts_device(0x5EED0003, 0, n, counts)
cap = n // 100
calls = torch.zeros((cap, 80), dtype=torch.uint8, device="cuda")
for i in range(K):
    k = eng.infer_device(counts, n, calls, cap)
print("prof_read: %d batches of %d sites, %d calls each" % (K, n, k))
