"""Small invocation of every kernel, meant to be run under compute-sanitizer:
    compute-sanitizer --tool memcheck python tests/sanitize_smoke.py
Sizes are deliberately ragged (not multiples of 32 / 128 / 512) to exercise every tail path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denovo-variant-caller-java_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import denovo_b200 as D  # noqa: E402
import oracle_lib as O  # noqa: E402
import synth_inputs as S  # noqa: E402


def main():
    counts = O.synth_counts(0x5EED0001, 0, 20011)
    deep = O.synth_counts(0x5EED0005, 0, 3001, depth=200, seq_err=1e-3)
    for method in ("BAYES", "MAP", "LRT"):
        with D.TrioEngine(method=method) as eng:
            a, da = eng.infer(counts, dense=True)
            b, _ = eng.infer(counts.astype(np.uint8))
            c, _ = eng.infer(counts[:1])
            d, _ = eng.infer(counts[:33], dense=False)
            assert a.tobytes() == b.tobytes()
        with D.TrioEngine(1e-3, 1e-8, 1.0, method) as eng:
            eng.infer(deep, dense=True)
    with D.TrioEngine(force_direct=True) as eng:
        eng.infer(counts[:2003], dense=True)
    rs, bo, ba, pe = S.make_reads(5, 1000, 2500, depth=30, var_len=True)
    with D.TrioEngine() as eng:
        eng.pileup(rs, bo, ba, pe, np.arange(950, 3600, dtype=np.int64))           # dense tiles + ragged tail
        eng.pileup(rs, bo, ba, pe, np.arange(1000, 3400, 37, dtype=np.int64))      # sparse tiles
        calls = S.make_callset(3, 7001)
        eng.mendel_filter(calls)
        eng.mendel_filter(calls[:5])
        import torch
        buf = torch.empty((5003, 12), dtype=torch.int16, device="cuda")
        eng.synth_counts_device(7, 11, 5003, buf)
        eng.synth_counts_at(7, np.arange(0, 999, 3))
        # asynchronous batches over the context's lanes, and a two-rank device group (peer-store gather)
        out = torch.zeros((5003, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
        h_nc = torch.zeros(4, dtype=torch.int64).pin_memory()
        for i in range(3):
            eng.infer_device_async(buf, 5003 - 100 * i, out, 5003, h_nc.data_ptr() + 8 * i)
        eng.sync()
        # synchronous device calls: the first eager, the third a graph replay; the finish kernel's last CTA hands the
        # statistics to pinned host memory (one-chunk call)
        for i in range(3):
            k = eng.infer_device(buf, 5003, out, 5003)
        assert eng.stats()["n_calls"] == k
        eng.measure_screen(buf, 5003, reps=2)
        # reads -> calls: pileup chained to the READ pipeline with one host wait
        eng.infer_reads(rs, bo, ba, pe, np.arange(950, 3600, dtype=np.int64))
        with D.DeviceGroup([0, 0]) as grp:
            grp.infer(buf.cpu().numpy().view(np.uint16).reshape(-1, 3, 4))
    print("sanitize_smoke OK")


if __name__ == "__main__":
    main()
