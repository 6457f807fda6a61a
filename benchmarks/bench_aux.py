#!/usr/bin/env python
"""Secondary measurements (not the driver's bench line): the other BASELINE configs on one B200.
Each result is one JSON line.  CUDA-event timing on the launching stream, 3 warm-ups, inputs > L2
or freshly generated per run.

  config 3 variants : MAP / LRT / dense-output over 10M sites at 30x
  config 5          : 200x, eps 1e-3, mu sweep (underflow path): sites/s, hard-path rate
  config 2          : Mendelian filter, ~5M records per individual
  pileup            : reads -> counts, every-position layout
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "denovo-variant-caller-java_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import denovo_b200 as D  # noqa: E402

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, stream, reps=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def read_mode(n, method, depth, eps, mu, lrt=1.0, dense=False, seed=0x5EED0003, tag=""):
    stream = torch.cuda.Stream()
    with D.TrioEngine(eps, mu, lrt, method, time_kernels=True) as eng:
        eng.set_stream(stream.cuda_stream)
        counts = torch.empty((n, 12), dtype=torch.int16, device="cuda")
        eng.synth_counts_device(seed, 0, n, counts, depth=depth)
        cap = max(n // 20, 1024)
        calls = torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
        d = torch.zeros((n, D.SITE_DT.itemsize), dtype=torch.uint8, device="cuda") if dense else None
        out = {}

        def step():
            with torch.cuda.stream(stream):
                out["n"] = eng.infer_device(counts, n, calls, cap, d_dense=d)
        ms = timed(step, stream)
        st = eng.stats()
    print(json.dumps({"bench": "read_mode" + tag, "method": method, "sites": n, "depth": depth, "eps": eps, "mu": mu,
                      "lrt": lrt, "dense_output": dense, "ms_per_pass": ms, "sites_per_s": n / ms * 1e3,
                      "fused_kernel_ms": st["fused_kernel_ms"], "calls": out["n"], "hard_sites": st["n_hard"],
                      "hard_rate": st["n_hard"] / n}), flush=True)


def mendel(n_variants):
    import synth_inputs as S
    import oracle_lib as O
    t0 = time.time()
    # build a large stream by tiling a seeded block along the position axis (keeps generation fast)
    block = S.make_callset(31, 200_000)
    span = int(block["end"].max()) + 100
    reps = max(1, n_variants // 200_000)
    parts = []
    for r in range(reps):
        b = block.copy()
        b["start"] += r * span
        b["end"] += r * span
        b["variant_id"] += r * 200_000
        parts.append(b)
    calls = np.concatenate(parts)
    gen_s = time.time() - t0
    n = len(calls)
    stream = torch.cuda.Stream()
    with D.TrioEngine() as eng:
        eng.set_stream(stream.cuda_stream)
        d_calls = torch.from_numpy(calls.view(np.uint8).reshape(n, 32)).cuda()
        n_child = int((calls["person"] == 0).sum())
        d_out = torch.zeros((n_child + 1, 24), dtype=torch.uint8, device="cuda")
        res = {}

        def step():
            with torch.cuda.stream(stream):
                res["n"] = eng.mendel_filter_device(d_calls, n, d_out, n_child + 1)
        ms = timed(step, stream, reps=5)
        got = d_out[:res["n"]].cpu().numpy().view(D.PCALL_DT).reshape(-1)
    # CPU streaming oracle on one block (bit-exact check + baseline)
    t0 = time.time()
    n_ref, want = O.mendel_stream(block)
    cpu_s = time.time() - t0
    per_block = res["n"] // reps
    ok = bool(np.array_equal(got[:n_ref]["position"], want["position"]) and np.array_equal(got[:n_ref]["is_denovo"], want["is_denovo"])
              and per_block == n_ref)
    per_person = [int((calls["person"] == p).sum()) for p in (0, 1, 2)]
    print(json.dumps({"bench": "mendel_filter", "records": n, "records_per_person_child_mom_dad": per_person,
                      "ms_per_pass": ms, "records_per_s": n / ms * 1e3, "algorithmic_GBps": n * 32 / ms / 1e6,
                      "hbm_peak_GBps": HBM, "frac_of_hbm": n * 32 / ms / 1e6 / HBM, "position_calls": res["n"],
                      "denovo_calls": int(got["is_denovo"].sum()), "bit_exact_vs_oracle_block": ok,
                      "cpu_oracle_records_per_s_1thread": len(block) / cpu_s, "gen_s": gen_s}), flush=True)


def pileup(n_sites, depth=30):
    import synth_inputs as S
    import oracle_lib as O
    rs, bo, ba, pe = S.make_reads(41, 1, n_sites, depth=depth)
    pos = np.arange(1, n_sites + 1, dtype=np.int64)
    stream = torch.cuda.Stream()
    with D.TrioEngine() as eng:
        eng.set_stream(stream.cuda_stream)
        d = [torch.from_numpy(x).cuda() for x in (rs, bo, ba, pe, pos)]
        d_counts = torch.zeros((n_sites, 12), dtype=torch.int16, device="cuda")

        def step():
            with torch.cuda.stream(stream):
                eng.pileup_device(d[0], d[1], d[2], d[3], len(rs), d[4], n_sites, d_counts)
        ms = timed(step, stream, reps=5)
        got = d_counts.cpu().numpy().view(np.uint16).reshape(n_sites, 3, 4)
        # secondary metric of SURVEY.md 8(d): reads -> counts -> calls, chained on the device
        cap = max(n_sites // 20, 1024)
        d_calls = torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
        chain = {}

        def step2():
            with torch.cuda.stream(stream):
                eng.pileup_device(d[0], d[1], d[2], d[3], len(rs), d[4], n_sites, d_counts)
                chain["calls"] = eng.infer_device(d_counts, n_sites, d_calls, cap)
        ms_chain = timed(step2, stream, reps=5)
    m = 20000
    sel = slice(n_sites // 2, n_sites // 2 + m)
    lo, hi = pos[sel][0] - 200, pos[sel][-1] + 1
    keep = (rs >= lo) & (rs <= hi)
    idx = np.nonzero(keep)[0]
    bo2 = np.zeros(len(idx) + 1, np.int64)
    np.cumsum(np.diff(bo)[idx], out=bo2[1:])
    ba2 = np.concatenate([ba[bo[i]:bo[i + 1]] for i in idx])
    t0 = time.time()
    rc, want = O.pileup(rs[idx], bo2, ba2, pe[idx], pos[sel])
    cpu_s = time.time() - t0
    alg_bytes = len(ba) + len(rs) * 17 + n_sites * (8 + 24)
    print(json.dumps({"bench": "pileup", "sites": n_sites, "reads": len(rs), "bases": int(len(ba)), "ms_per_pass": ms,
                      "sites_per_s": n_sites / ms * 1e3, "algorithmic_bytes_per_site": alg_bytes / n_sites,
                      "algorithmic_GBps": alg_bytes / ms / 1e6, "hbm_peak_GBps": HBM,
                      "frac_of_hbm": alg_bytes / ms / 1e6 / HBM,
                      "bit_exact_vs_oracle_window": bool(rc == 0 and np.array_equal(got[sel], want)),
                      "reads_to_calls_ms": ms_chain, "reads_to_calls_sites_per_s": n_sites / ms_chain * 1e3,
                      "reads_to_calls_calls": int(chain["calls"])}), flush=True)


def genome(total=3_095_693_983, chunk=1 << 26, depth=30):
    """BASELINE config 4 on ONE GPU: every position of an hg19-sized axis, counts generated procedurally
    per 67M-site chunk on the device (74 GB of counts are never materialised at once); inference time
    only (generation excluded), plus the generator's own time."""
    stream = torch.cuda.Stream()
    with D.TrioEngine(1e-2, 1e-8, 1.0, "BAYES", time_kernels=True) as eng:
        eng.set_stream(stream.cuda_stream)
        counts = torch.empty((chunk, 12), dtype=torch.int16, device="cuda")
        cap = chunk // 50
        calls = torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
        infer_ms = gen_ms = 0.0
        n_calls = hard = 0
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        for first in range(0, total, chunk):
            m = min(chunk, total - first)
            with torch.cuda.stream(stream):
                e0.record(stream)
                eng.synth_counts_device(0x5EED0004, first, m, counts, depth=depth)
                e1.record(stream)
                n_calls += eng.infer_device(counts, m, calls, cap)
                e2.record(stream)
            torch.cuda.synchronize()
            gen_ms += e0.elapsed_time(e1)
            infer_ms += e1.elapsed_time(e2)
            hard += eng.stats()["n_hard"]
    print(json.dumps({"bench": "genome_every_position_1gpu", "sites": total, "depth": depth, "infer_ms": infer_ms,
                      "sites_per_s": total / infer_ms * 1e3, "generator_ms": gen_ms, "calls": n_calls,
                      "hard_sites": hard}), flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["read", "stress", "mendel", "pileup", "genome"]
    N = 10_000_000
    if "read" in what:
        read_mode(N, "BAYES", 30, 1e-2, 1e-8)
        read_mode(N, "MAP", 30, 1e-2, 1e-8)
        read_mode(N, "LRT", 30, 1e-2, 1e-8, lrt=10.0)
        read_mode(N, "BAYES", 30, 1e-2, 1e-8, dense=True, tag="_dense")
    if "stress" in what:
        for mu in (1e-10, 1e-8, 1e-6):
            read_mode(N, "BAYES", 200, 1e-3, mu, seed=0x5EED0005, tag="_config5")
    if "mendel" in what:
        mendel(6_000_000)
    if "pileup" in what:
        pileup(2_000_000)
    if "genome" in what:
        genome()
