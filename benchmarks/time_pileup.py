#!/usr/bin/env python
"""Times the pileup kernel alone (2M consecutive sites, 30x, read length 100) and checks a window against the oracle.
Environment knobs of the library (DNV_PILEUP_S, DNV_PILEUP_KERNEL) apply."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denovo-variant-caller-java_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import denovo_b200 as D  # noqa: E402
import oracle_lib as O  # noqa: E402
import synth_inputs as S  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
rs, bo, ba, pe = S.make_reads(41, 1, n, depth=30)
pos = np.arange(1, n + 1, dtype=np.int64)
eng = D.TrioEngine()
d = [torch.from_numpy(x).cuda() for x in (rs, bo, ba, pe, pos)]
counts = torch.zeros((n, 12), dtype=torch.int16, device="cuda")
for _ in range(3):
    eng.pileup_device(d[0], d[1], d[2], d[3], len(rs), d[4], n, counts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eng.pileup_device(d[0], d[1], d[2], d[3], len(rs), d[4], n, counts)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
got = counts.cpu().numpy().view(np.uint16).reshape(n, 3, 4)
m = 50000
sel = slice(n // 2, n // 2 + m)
keep = (rs >= pos[sel][0] - 200) & (rs <= pos[sel][-1] + 1)
idx = np.nonzero(keep)[0]
bo2 = np.zeros(len(idx) + 1, np.int64)
np.cumsum(np.diff(bo)[idx], out=bo2[1:])
ba2 = np.concatenate([ba[bo[i]:bo[i + 1]] for i in idx])
rc, want = O.pileup(rs[idx], bo2, ba2, pe[idx], pos[sel])
alg = len(ba) + len(rs) * 17 + n * 32
print("pileup: %.3f ms  %.2f G sites/s  %.0f GB/s algorithmic  exact=%s  S=%s kernel=%s" % (
    ms, n / ms / 1e6, alg / ms / 1e6, rc == 0 and np.array_equal(got[sel], want),
    os.environ.get("DNV_PILEUP_S", "default"), os.environ.get("DNV_PILEUP_KERNEL", "default")))
