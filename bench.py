#!/usr/bin/env python
"""bench.py — trio candidate sites/sec, READ-mode BAYES inference at 30x (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]       our arm (one process per GPU under torchrun)
  python bench.py --impl reference ...                       reference arm: the reference's CPU algorithm
                                                            (C restatement, no JVM in this image) on host cores

One "step" = one pass of the hot path over one batch: packed pileup counts resident in HBM
-> compacted call records in HBM on rank 0 (N>1: stored there by every rank's kernels over NVLink, the
gather of include/denovo_b200.h: dnv_gather; the gathered records are verified against single-context
inference after the timed region).  Everything timed goes through the C ABI; Python only owns buffers.
Prints ONE JSON line on rank 0; `aux` in that line holds the other BASELINE configs (driver-run).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "denovo-variant-caller-java_b200", "python"))

METRIC = "trio candidate sites/sec (READ-mode BAYES, 30x)"
SITES_PER_GPU = 10_000_000
SEED = 0x5EED0003
DEPTH, EPS, MU, LRT = 30, 1e-2, 1e-8, 1.0
GROUP = 64  # asynchronous steps per host synchronisation (and per gather round)
NBUF = int(os.environ.get('BENCH_NBUF', '4'))    # distinct input buffers (distinct shards of the axis) the steps cycle through; also gather generations
# ALGORITHMIC work per site:
#   integer screen (dominant kernel, HBM-bound): 24 B of counts in + 1 bit of decision out
#   list kernel (only the sites the screen cannot decide): sparse factored FP64 form, FMA = 2
BYTES_PER_SITE = 24.125
FLOP_PER_SITE = 451.0
# DRAM traffic of one screen launch over 10M sites from the ncu --set full capture committed under profiles/
# (dram__bytes_read.sum + dram__bytes_write.sum), per site
NCU_TRAFFIC_BYTES_PER_SITE = 25.08  # profiles/r02_read_pipeline_ncu.md: 246.51 MB read + 4.32 MB written per 10 M sites


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML in a thread DURING the timed region
    (the region lasts tens of milliseconds, far below nvidia-smi's polling period)."""

    def __init__(self, device):
        self.device, self.sm, self.power, self.reasons, self.stop_flag = device, [], [], 0, False
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES if it holds indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = device
            try:
                if vis:
                    idx = int(vis.split(",")[device])
            except ValueError:
                idx = device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def _loop(self):
        # NVML queries go through the same kernel driver as the launches of the timed loop and were measured to
        # stall them (a 20-step region ran 15 % slower with a query every 0.5 ms): the first sample is taken 0.4 ms
        # into the region, the following ones every 4 ms — a 20-step region (1.2 ms) sees one, longer ones several
        nv = self.nv
        time.sleep(0.0004)
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                if len(self.sm) % 4 == 1:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.004)

    def start(self):
        if self.ok:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self.stop_flag = True
        self.t.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(k for k, v in names.items() if self.reasons & v), "samples": len(sm),
                "power_w_max": max(self.power) if self.power else None}


def cpu_reference_run(n_sites, threads, faithful=True):
    """The reference's algorithm (direct form, reference operation order) on host cores."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    counts = O.synth_counts(SEED, 0, n_sites, depth=DEPTH, seq_err=EPS)
    prm = O.params(EPS, MU, LRT)
    O.infer(prm, counts[:1000], faithful=faithful, threads=threads)
    t0 = time.perf_counter()
    res = O.infer(prm, counts, faithful=faithful, threads=threads)
    dt = time.perf_counter() - t0
    return n_sites / dt, int(res["denovo_bayes"].sum())


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    sample = 200_000
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    counts = O.synth_counts(SEED, 0, sample, depth=DEPTH, seq_err=EPS)
    prm = O.params(EPS, MU, LRT)
    for _ in range(args.warmup):
        O.infer(prm, counts[:20000], faithful=True, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.infer(prm, counts, faithful=True, threads=threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "sites/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "READ-mode BAYES inference, synthetic trio pileups at 30x, eps 1e-2, mu 1e-8 "
                                   "(BASELINE configs[2]); %d-site sample per step" % sample,
                       "same_workload_as_our_arm": "a bounded sample of it: the CPU path costs the same per site at any batch "
                                                   "size (one independent task per site, M/ReadCaller.java:91-95), so sites/s compares"},
            "cpu_baseline": {"value": value, "unit": "sites/s", "cores": threads, "kind": "port",
                             "sample": "%d sites/step x %d steps; C restatement of the reference's direct form "
                                       "(3 log + 1 exp per genotype triple, as DenovoBayesNet.java:220-233,269-273); "
                                       "no JVM in this image so the Java itself cannot run" % (sample, args.steps)},
            "e2e": {"value": value, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def timed_sync_calls(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def aux_read_config(D, torch, local, method, depth, eps, mu, lrt, n, seed, steps=12):
    """One BASELINE READ configuration on this GPU: asynchronous batches over NBUF distinct buffers (device
    resident), CUDA events on the context's stream."""
    stream = torch.cuda.Stream(device=local)
    with D.TrioEngine(eps, mu, lrt, method, device=local) as eng:
        eng.set_stream(stream.cuda_stream)
        bufs = []
        for b in range(2):
            c = torch.empty((n, 12), dtype=torch.int16, device="cuda")
            eng.synth_counts_device(seed, b * n, n, c, depth=depth, seq_err=eps)
            bufs.append(c)
        cap = max(n // 20, 1024)
        calls = [torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda") for _ in range(2)]
        h = torch.zeros(2, dtype=torch.int64).pin_memory()

        def run(k):
            for i in range(k):
                eng.infer_device_async(bufs[i % 2], n, calls[i % 2], cap, h.data_ptr() + 8 * (i % 2))
            eng.flush()
        with torch.cuda.stream(stream):
            run(4)
            eng.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run(steps)
            e1.record(stream)
            eng.sync()
        st = eng.stats()
        ms = e0.elapsed_time(e1) / steps
    return {"method": method, "depth": depth, "eps": eps, "mu": mu, "lrt": lrt, "sites": n, "ms_per_step": ms,
            "sites_per_s": n / ms * 1e3, "calls_per_step": st["n_calls"] // steps, "hard_sites_per_step": st["n_hard"] // steps,
            "undecided_per_step": (st["n_sites"] - st["n_screened_out"]) // steps}


def aux_undecided(D, torch, np, local, fractions, n=10_000_000, steps=8):
    """The screen-hostile case: a fraction of the sites replaced by adversarial de-novo-pattern pileups
    (parents homozygous, child heterozygous, Poisson depths) that the integer screen cannot decide, so they all
    reach the FP64 list kernel.  Parity: a 100k-site sample of every mix against the CPU oracle."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    out = []
    stream = torch.cuda.Stream(device=local)
    with D.TrioEngine(EPS, MU, LRT, "BAYES", device=local) as eng:
        eng.set_stream(stream.cuda_stream)
        base = torch.empty((n, 12), dtype=torch.int16, device="cuda")
        eng.synth_counts_device(SEED, 0, n, base, depth=DEPTH)
        g = torch.Generator(device="cuda")
        g.manual_seed(1234)
        # de novo pattern: parents n_X ~ Poisson(30) of allele X, child split between X and Y
        x = torch.randint(0, 4, (n,), device="cuda", generator=g)
        y = (x + torch.randint(1, 4, (n,), device="cuda", generator=g)) % 4
        dep = torch.poisson(torch.full((n, 3), float(DEPTH), device="cuda"), generator=g).to(torch.int16)
        alt = (dep[:, 2].float() * (0.25 + 0.5 * torch.rand(n, device="cuda", generator=g))).to(torch.int16)
        hostile = torch.zeros((n, 3, 4), dtype=torch.int16, device="cuda")
        idx = torch.arange(n, device="cuda")
        hostile[idx, 0, x] = dep[:, 0]
        hostile[idx, 1, x] = dep[:, 1]
        hostile[idx, 2, x] = dep[:, 2] - alt
        hostile[idx, 2, y] += alt
        hostile = hostile.reshape(n, 12)
        u = torch.rand(n, device="cuda", generator=g)
        cap = n
        calls = torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
        h = torch.zeros(1, dtype=torch.int64).pin_memory()
        for f in fractions:
            mix = torch.where((u < f)[:, None], hostile, base).contiguous()
            with torch.cuda.stream(stream):
                for _ in range(2):  # every lane allocates its scratch on its first batch; after the first sync the context
                    for _ in range(4):  # may switch to the all-sites path (screen-hostile input), which allocates its own
                        eng.infer_device_async(mix, n, calls, cap, h)
                    eng.sync()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(steps):
                    eng.infer_device_async(mix, n, calls, cap, h)
                eng.flush()
                e1.record(stream)
                eng.sync()
            st = eng.stats()
            ms = e0.elapsed_time(e1) / steps
            m = 100_000
            sample = mix[:m].cpu().numpy().view(np.uint16).reshape(m, 3, 4)
            ref = O.infer(O.params(EPS, MU, LRT), sample, threads=os.cpu_count() or 1)
            got = calls[:int(h[0])].cpu().numpy().view(D.CALL_DT).reshape(-1)
            ok = bool(np.array_equal(got["site"][got["site"] < m], np.nonzero(ref["denovo_bayes"])[0]))
            out.append({"hostile_fraction": f, "path": "all-sites fused kernel (adaptive)" if st["n_screened_out"] == 0 else "screen + list",
                        "ms_per_step": ms, "sites_per_s": n / ms * 1e3, "calls_per_step": st["n_calls"] // steps,
                        "parity_100k_sample_vs_oracle": ok})
            if not ok:
                raise SystemExit("undecided sweep: parity FAILED at hostile fraction %g" % f)
    return out


def aux_mendel(D, torch, np, local):
    """BASELINE config 2: VARIANT-mode Mendelian filter over ~11.8 M (variant, call) records (~2.3 M child SNPs,
    ~4.7 M records per parent), device resident; bit-exact check of one 200k-variant block against the oracle."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import synth_inputs as S
    block = S.make_callset(31, 200_000)
    span = int(block["end"].max()) + 100
    reps = 30
    parts = []
    for r in range(reps):
        b = block.copy()
        b["start"] += r * span
        b["end"] += r * span
        b["variant_id"] += r * 200_000
        parts.append(b)
    calls = np.concatenate(parts)
    n = len(calls)
    stream = torch.cuda.Stream(device=local)
    with D.TrioEngine(device=local) as eng:
        eng.set_stream(stream.cuda_stream)
        d_calls = torch.from_numpy(calls.view(np.uint8).reshape(n, 32)).cuda()
        n_child = int((calls["person"] == 0).sum())
        d_out = torch.zeros((n_child + 1, 24), dtype=torch.uint8, device="cuda")
        res = {}

        def step():
            res["n"] = eng.mendel_filter_device(d_calls, n, d_out, n_child + 1)
        with torch.cuda.stream(stream):
            for _ in range(3):
                step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(5):
                step()
            e1.record(stream)
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        got = d_out[:res["n"]].cpu().numpy().view(D.PCALL_DT).reshape(-1)
    n_ref, want = O.mendel_stream(block)
    ok = bool(res["n"] == reps * n_ref and all(np.array_equal(got[:n_ref][f], want[f]) for f in ("position", "child", "mom", "dad", "is_denovo")))
    hbm, _ = load_peaks()
    return {"records": n, "ms_per_pass": ms, "records_per_s": n / ms * 1e3, "algorithmic_GBps": n * 32 / ms / 1e6,
            "frac_of_hbm": n * 32 / ms / 1e6 / hbm, "position_calls": int(res["n"]), "denovo_calls": int(got["is_denovo"].sum()),
            "bit_exact_vs_oracle_block": ok}


def tile_reads(np, rs, bo, ba, pe, span, reps):
    """`reps` copies of a read set side by side on the position axis (copy k shifted by k * span), still sorted by
    (person, start): a large input without generating every base on the CPU."""
    lens = np.diff(bo)
    out_rs, out_len, out_ba, out_pe = [], [], [], []
    for p in range(3):
        sel = np.nonzero(pe == p)[0]
        if len(sel) == 0:
            continue
        lo0, hi = sel[0], sel[-1] + 1
        first_pos = int(rs[lo0:hi].max()) - span + 1   # copies after the first keep the reads that start at or after
        for k in range(reps):                          # their own first position: the concatenation stays sorted
            lo = lo0 if k == 0 else lo0 + int(np.searchsorted(rs[lo0:hi], max(first_pos, 1), side="left"))
            out_rs.append(rs[lo:hi] + k * span)
            out_len.append(lens[lo:hi])
            out_ba.append(ba[bo[lo]:bo[hi]])
            out_pe.append(pe[lo:hi])
    rs2, ba2, pe2 = np.concatenate(out_rs), np.concatenate(out_ba), np.concatenate(out_pe)
    bo2 = np.zeros(len(rs2) + 1, np.int64)
    np.cumsum(np.concatenate(out_len), out=bo2[1:])
    return rs2, bo2, ba2, pe2


def aux_pileup(D, torch, np, local, n_sites=8_000_000, reps=4):
    """Secondary metric of SURVEY.md 8(d): reads -> counts (pileup kernel) and reads -> counts -> calls chained on
    the device, every-position layout, 30x, read length 100 (8 M positions: four copies of a 2 M-position read set
    side by side)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import synth_inputs as S
    rs, bo, ba, pe = S.make_reads(41, 1, n_sites // reps, depth=DEPTH)
    rs, bo, ba, pe = tile_reads(np, rs, bo, ba, pe, n_sites // reps, reps)
    pos = np.arange(1, n_sites + 1, dtype=np.int64)
    stream = torch.cuda.Stream(device=local)
    with D.TrioEngine(device=local) as eng:
        eng.set_stream(stream.cuda_stream)
        d = [torch.from_numpy(x).cuda() for x in (rs, bo, ba, pe, pos)]
        d_counts = torch.zeros((n_sites, 12), dtype=torch.int16, device="cuda")
        cap = max(n_sites // 20, 1024)
        d_calls = torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
        chain = {}

        def pile():
            eng.pileup_device(d[0], d[1], d[2], d[3], len(rs), d[4], n_sites, d_counts)

        def both():
            eng.pileup_device(d[0], d[1], d[2], d[3], len(rs), d[4], n_sites, d_counts)
            chain["calls"] = eng.infer_device(d_counts, n_sites, d_calls, cap)

        def reads_api():  # dnv_infer_reads_device: pileup pipeline chained to the READ pipeline inside the library
            chain["api_calls"] = eng.infer_reads_device(d[0], d[1], d[2], d[3], len(rs), d[4], n_sites, d_calls, cap)

        def fused():      # the same entry point with the single fused kernel (DNV_READS_FUSED=1)
            chain["fused_calls"] = eng.infer_reads_device(d[0], d[1], d[2], d[3], len(rs), d[4], n_sites, d_calls, cap)
        times = []
        for fn in (pile, both, reads_api, fused):
            os.environ["DNV_READS_FUSED"] = "1" if fn is fused else "0"
            with torch.cuda.stream(stream):
                for _ in range(3):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record(stream)
                for _ in range(5):
                    fn()
                e1.record(stream)
                torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / 5)
        os.environ.pop("DNV_READS_FUSED", None)
        pile()
        got = d_counts.cpu().numpy().view(np.uint16).reshape(n_sites, 3, 4)
    m = 20000
    sel = slice(n_sites // 2, n_sites // 2 + m)
    keep = (rs >= pos[sel][0] - 200) & (rs <= pos[sel][-1] + 1)
    idx = np.nonzero(keep)[0]
    bo2 = np.zeros(len(idx) + 1, np.int64)
    np.cumsum(np.diff(bo)[idx], out=bo2[1:])
    ba2 = np.concatenate([ba[bo[i]:bo[i + 1]] for i in idx])
    rc, want = O.pileup(rs[idx], bo2, ba2, pe[idx], pos[sel])
    alg = len(ba) + len(rs) * 17 + n_sites * (8 + 24)
    hbm, _ = load_peaks()
    return {"sites": n_sites, "reads": len(rs), "bases": int(len(ba)), "pileup_ms": times[0], "pileup_sites_per_s": n_sites / times[0] * 1e3,
            "algorithmic_bytes_per_site": alg / n_sites, "algorithmic_GBps": alg / times[0] / 1e6, "frac_of_hbm": alg / times[0] / 1e6 / hbm,
            "bit_exact_vs_oracle_window": bool(rc == 0 and np.array_equal(got[sel], want)),
            "pileup_then_infer_two_calls_ms": times[1], "reads_to_calls_calls": int(chain["calls"]),
            "reads_to_calls_ms": times[2], "reads_to_calls_sites_per_s": n_sites / times[2] * 1e3,
            "single_kernel_reads_to_calls_ms": times[3], "single_kernel_reads_to_calls_sites_per_s": n_sites / times[3] * 1e3,
            "all_paths_same_calls": bool(chain["fused_calls"] == chain["calls"] == chain["api_calls"])}


def genome_walk(D, torch, np, dist, eng, stream, gather, rank, world, local, chunk=1 << 26, seed=0x5EED0004):
    """BASELINE config 4: every position of an hg19-sized axis (3,095,693,983 sites, 30x), interval-sharded over the
    ranks (STRONG scaling: the genome is fixed), counts generated procedurally per chunk on the device (the 74 GB
    of counts are never materialised), call records gathered on rank 0.  Generation is excluded from the timed
    inference (separate events); returns per-rank numbers, max-reduced by the caller."""
    from denovo_b200 import sharding
    total = 3_095_693_983
    lo, hi = sharding.shard_range(total, world, rank)
    bufs = [torch.empty((chunk, 12), dtype=torch.int16, device="cuda") for _ in range(2)]
    cap = chunk // 50
    calls = [torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda") for _ in range(2)] if gather is None else None
    h = torch.zeros(2, dtype=torch.int64).pin_memory()
    infer_ms = gen_ms = 0.0
    n_calls = 0
    per_gen = []
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    k = 0
    with torch.cuda.stream(stream):
        # untimed: four chunk-sized batches so that every lane of the context has its scratch for 2^26-site chunks
        # (a one-time cudaMalloc per lane and process: 86 ms for the first walk of a fresh context, 18 ms for the next)
        warm_calls = torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
        eng.synth_counts_device(seed, lo, min(chunk, hi - lo), bufs[0], depth=DEPTH)
        for _ in range(4):
            eng.infer_device_async(bufs[0], min(chunk, hi - lo), warm_calls, cap, h.data_ptr(), site_base=lo)
        eng.flush()
        eng.sync()
        del warm_calls
        for first in range(lo, hi, chunk):
            m = min(chunk, hi - first)
            b = k % 2
            evs[0].record(stream)
            eng.synth_counts_device(seed, first, m, bufs[b], depth=DEPTH)
            evs[1].record(stream)
            if gather is not None:
                # every rank has generated its chunk before any rank starts the timed inference: rank 0's wait then
                # measures the gather, not the other ranks' generators
                torch.cuda.synchronize()
                dist.barrier()
            evs[3].record(stream)
            if gather is None:
                eng.infer_device_async(bufs[b], m, calls[b], cap, h.data_ptr() + 8 * b, site_base=first)
                eng.flush()
            else:
                gather.submit(bufs[b], m, first, b)
                gather.arrive()
                if rank == 0:
                    gather.wait()
            evs[2].record(stream)
            eng.sync()
            if gather is not None:
                dist.barrier()  # rank 0 has every rank's records of this chunk round
                if rank == 0:
                    got, per = gather.fetch(b, cap=world * gather.cap)
                    n_calls += len(got)
                dist.barrier()
            else:
                n_calls += int(h[b])
            gen_ms += evs[0].elapsed_time(evs[1])
            infer_ms += evs[3].elapsed_time(evs[2])
            k += 1
    return {"sites": hi - lo, "infer_ms": infer_ms, "gen_ms": gen_ms, "calls": n_calls, "chunks": k}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--sites", type=int, default=SITES_PER_GPU)
    ap.add_argument("--method", default="BAYES")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the other BASELINE configs (aux object of the line)")
    ap.add_argument("--no-screen", action="store_true", help="evaluate every site with the fused FP64 kernel")
    ap.add_argument("--undecided", type=float, nargs="*", default=None,
                    help="screen-hostile sweep: fractions of adversarial de-novo-pattern sites (default in aux: 0.01 0.1 0.5 1.0)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # one hardware queue per lane stream
    import numpy as np
    import torch
    import torch.distributed as dist
    import denovo_b200 as D

    # NCCL may print a version banner on stdout: keep stdout for the ONE JSON line
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.sites
    K, W = args.steps, max(args.warmup, 3)

    # eng: the product context (lanes inside: consecutive batches overlap on the GPU); eng_t: a serial context whose
    # kernels are bracketed by CUDA events one by one (DNV_F_TIME_KERNELS) for the roofline of the dominant kernel
    stream = torch.cuda.Stream(device=local)
    eng = D.TrioEngine(EPS, MU, LRT, args.method, device=local, no_screen=args.no_screen)
    eng.set_stream(stream.cuda_stream)
    eng_t = D.TrioEngine(EPS, MU, LRT, args.method, device=local, time_kernels=True, no_screen=args.no_screen)
    eng_t.set_stream(stream.cuda_stream)

    # ---- synthetic shard of this rank: NBUF distinct buffers = NBUF consecutive intervals of the axis
    # (interval sharding: rank r owns sites [r * NBUF * n, (r + 1) * NBUF * n)); step i reads buffer i % NBUF
    def shard_first(r, b):
        return (r * NBUF + b) * n
    d_counts = []
    for b in range(NBUF):
        c = torch.empty((n, 12), dtype=torch.int16, device="cuda")
        eng.synth_counts_device(SEED, shard_first(rank, b), n, c, depth=DEPTH)
        d_counts.append(c)
    cap = max(n // 100, 1024)
    torch.cuda.synchronize()

    # N>1: every rank's kernels store records + counts straight into rank 0's HBM (dnv_gather over IPC-mapped peer
    # memory; the 64-byte handle is the only thing torch.distributed carries); generation = buffer index
    gather = None
    if world > 1:
        gather = D.Gather(eng, world, rank, cap, gens=NBUF)
        gather.exchange(dist)
        d_calls = None
    else:
        d_calls = [torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda") for _ in range(NBUF)]
    h_nc = torch.zeros(NBUF, dtype=torch.int64).pin_memory()
    torch.cuda.set_stream(stream)

    def run_steps(k_steps):
        """k_steps pipelined steps; one host synchronisation (and one gather round) per GROUP steps."""
        acc = {"n_launches": 0, "n_screened_out": 0, "n_hard": 0, "n_calls": 0, "n_sites": 0}
        done = 0
        while done < k_steps:
            m = min(GROUP, k_steps - done)
            for i in range(done, done + m):
                b = i % NBUF
                if gather is not None:
                    gather.submit(d_counts[b], n, shard_first(rank, b), b)
                else:
                    eng.infer_device_async(d_counts[b], n, d_calls[b], cap, h_nc.data_ptr() + 8 * b, site_base=shard_first(0, b))
            if gather is not None:
                gather.arrive()      # device-side: flush + release flag in rank 0's HBM
                if rank == 0:
                    gather.wait()    # rank 0's stream waits until every rank's records of the round have landed
            else:
                eng.flush()
            done += m
            if done < k_steps:
                eng.sync()
                st = eng.stats()
                for key in acc:
                    acc[key] += st[key]
        return acc

    def finish(acc):
        eng.sync()
        st = eng.stats()
        for key in acc:
            acc[key] += st[key]
        return acc

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    finish(run_steps(max(W, 2 * NBUF)))  # scratch of every lane is allocated on its first batch: all of them warm up
    sampler = ClockSampler(local)
    barrier()
    if rank == 0 and not os.environ.get('BENCH_NO_SAMPLER'):
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    acc = run_steps(K)
    e1.record(stream)     # after flush / wait on the same stream: everything of the K steps is inside the bracket
    acc = finish(acc)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = (sampler.stop() if not os.environ.get('BENCH_NO_SAMPLER') else None) if rank == 0 else None
    launches = acc["n_launches"]
    screened_out = acc["n_screened_out"] // K
    n_hard = acc["n_hard"] // K
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    per_rank_ms = [ms]
    if world > 1:
        allt = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank_ms = [float(x.item()) for x in allt]
    ms_max = max(per_rank_ms)
    value = world * n * K / (ms_max * 1e-3)

    # ---- parity of what the timed region produced: rank 0 regenerates every rank's shard of each generation,
    # infers it with the serial context and compares the gathered records BYTE FOR BYTE
    verify = {"verified": None}
    if rank == 0:
        tmp_counts = torch.empty((n, 12), dtype=torch.int16, device="cuda")
        tmp_calls = torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda")
        gens = sorted(set((K - 1 - j) % NBUF for j in range(min(K, NBUF))))
        n_rec = 0
        for b in gens:
            want = []
            for r in range(world):
                eng_t.synth_counts_device(SEED, shard_first(r, b), n, tmp_counts, depth=DEPTH)
                k = eng_t.infer_device(tmp_counts, n, tmp_calls, cap)
                w = tmp_calls[:k].cpu().numpy().view(D.CALL_DT).reshape(-1).copy()
                w["site"] += shard_first(r, b)
                want.append(w)
            want = np.concatenate(want)
            if gather is not None:
                got, per = gather.fetch(b, cap=world * cap)
            else:
                got = d_calls[b][:int(h_nc[b])].cpu().numpy().view(D.CALL_DT).reshape(-1)
            if got.tobytes() != want.tobytes():
                print("GATHER PARITY FAILED: generation %d, %d records gathered, %d expected" % (b, len(got), len(want)), file=sys.stderr)
                os.dup2(real_stdout, 1)
                print(json.dumps({"metric": METRIC, "error": "gathered records differ from single-context inference"}), flush=True)
                return 1
            n_rec += len(got)
        total_calls = n_rec // len(gens)
        verify = {"verified": True, "generations_checked": len(gens), "records_checked": n_rec,
                  "how": "rank 0 regenerates every rank's shard, infers it on a serial context and compares the "
                         "records of the last %d step(s) byte for byte" % len(gens)}
        del tmp_counts, tmp_calls
    else:
        total_calls = 0
    barrier()

    # ---- the same step through the synchronous entry point (one call = submit + wait + count on the host), on the
    # product context (lanes + graph replay), and the kernel times on the serial, event-bracketed context
    cptr, clptr = d_counts[0], (d_calls[0] if d_calls is not None else torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda"))
    cp_raw, cl_raw = cptr.data_ptr(), clptr.data_ptr()  # raw addresses: the loop times the C entry point, not torch's accessors
    sync_s = timed_sync_calls(lambda: eng.infer_device(cp_raw, n, cl_raw, cap), 30)
    Kt = 20
    for _ in range(3):
        eng_t.infer_device(cptr, n, clptr, cap)
    fused_ms = screen_ms = 0.0
    for _ in range(Kt):
        eng_t.infer_device(cptr, n, clptr, cap)
        st = eng_t.stats()
        fused_ms += st["fused_kernel_ms"]
        screen_ms += st["screen_kernel_ms"]
    k_ms, s_ms = fused_ms / Kt, screen_ms / Kt
    # the dominant kernel alone: 20 back-to-back launches on one stream between one pair of CUDA events (the event pairs
    # above bracket ONE launch on an idle stream each, so they also hold that launch's latency)
    s_b2b_ms = eng_t.measure_screen(cptr, n, 20) if s_ms > 0 else 0.0

    # ---- end to end through the host API: pinned host counts -> calls in pinned host memory
    # the host packs counts as uint8 (12 B/site) when every per-allele depth is below 256, else uint16: an INPUT
    # FORMAT of the ABI (dnv_infer_counts8), packed once outside the timed region like any other input preparation
    u8 = bool(int(d_counts[0].max().item()) < 256)
    if u8:
        h_counts = torch.empty((n, 12), dtype=torch.uint8).pin_memory()
        h_counts.copy_(d_counts[0].to(torch.uint8))
    else:
        h_counts = torch.empty((n, 12), dtype=torch.int16).pin_memory()
        h_counts.copy_(d_counts[0])
    h_calls = torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8).pin_memory()
    eng.set_stream(None)
    for _ in range(2):
        e2e_calls = eng.infer_host_into(h_counts, n, h_calls, cap, u8=u8)
    Ke = max(3, K // 4)
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_calls = eng.infer_host_into(h_counts, n, h_calls, cap, u8=u8)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n * Ke / float(te.item())
    eng.set_stream(stream.cuda_stream)

    # ---- aux: the other BASELINE configs, driver-run in the same line
    aux = None
    if not args.no_aux:
        aux = {}
        if world > 1:
            # config 4 (strong scaling over the ranks) with its gather
            gw = D.Gather(eng, world, rank, (1 << 26) // 50, gens=2)
            gw.exchange(dist)
            g = genome_walk(D, torch, np, dist, eng, stream, gw, rank, world, local)
            tt = torch.tensor([g["infer_ms"], g["gen_ms"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            gw.close()
            if rank == 0:
                aux["config4_genome_every_position"] = {"sites": 3_095_693_983, "n_gpus": world, "scaling": "strong",
                                                        "infer_ms_max_over_ranks": float(tt[0]), "sites_per_s": 3_095_693_983 / float(tt[0]) * 1e3,
                                                        "generator_ms_max_over_ranks": float(tt[1]), "calls_gathered_on_rank0": g["calls"],
                                                        "chunk_sites": 1 << 26}
        elif rank == 0:
            g = genome_walk(D, torch, np, dist, eng, stream, None, 0, 1, local)
            aux["config4_genome_every_position"] = {"sites": g["sites"], "n_gpus": 1, "scaling": "strong", "infer_ms_max_over_ranks": g["infer_ms"],
                                                    "sites_per_s": g["sites"] / g["infer_ms"] * 1e3, "generator_ms_max_over_ranks": g["gen_ms"],
                                                    "calls_gathered_on_rank0": g["calls"], "chunk_sites": 1 << 26}
        # config 5 on every rank (weak: 10 M sites per GPU), config 3 variants / config 2 / pileup on rank 0
        c5 = [aux_read_config(D, torch, local, "BAYES", 200, 1e-3, mu, 1.0, n, 0x5EED0005) for mu in (1e-10, 1e-6)]
        if world > 1:
            for c in c5:
                tt = torch.tensor([c["ms_per_step"]], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                c["ms_per_step"] = float(tt[0])
                c["sites_per_s"] = world * n / float(tt[0]) * 1e3
                c["n_gpus"] = world
        if rank == 0:
            aux["config5_200x_eps1e-3"] = c5
            aux["config3_MAP"] = aux_read_config(D, torch, local, "MAP", DEPTH, EPS, MU, 1.0, n, SEED)
            aux["config3_LRT_T10"] = aux_read_config(D, torch, local, "LRT", DEPTH, EPS, MU, 10.0, n, SEED)
            aux["config2_mendel_filter"] = aux_mendel(D, torch, np, local)
            aux["pileup_and_reads_to_calls"] = aux_pileup(D, torch, np, local)
            aux["undecided_sweep"] = aux_undecided(D, torch, np, local, args.undecided if args.undecided else [0.01, 0.1, 0.25, 0.5, 1.0])
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel
    fp64_peak, _ = eng_t.measure_fp64_peak()
    hbm_peak, peak_src = load_peaks()
    if s_ms > 0:
        # BAYES/LRT: the integer screen streams every site's counts and is the dominant kernel (HBM-bound)
        n_fp64 = n - screened_out
        ach_gbs = BYTES_PER_SITE * n / (s_b2b_ms * 1e-3) / 1e9
        f_ms = max(k_ms - s_ms, 1e-9)
        ach_tf = FLOP_PER_SITE * n_fp64 / (f_ms * 1e-3) / 1e12
        step_gbs = BYTES_PER_SITE * n / (ms_max / K * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                    "traffic": NCU_TRAFFIC_BYTES_PER_SITE * n, "kernel": "trio_screen_kernel", "kernel_ms": s_b2b_ms,
                    "kernel_ms_how": "CUDA events around 20 back-to-back launches of the kernel alone on one stream (dnv_measure_screen), / 20",
                    "kernel_ms_single_launch": s_ms, "frac_single_launch": BYTES_PER_SITE * n / (s_ms * 1e-3) / 1e9 / hbm_peak,
                    # kernel times from the serial context (each kernel alone, comparable with the serialised ncu launch
                    # list); in the value loop the list / finish kernels of one batch run BESIDE the screen of the
                    # next, so ms_per_step is below their sum
                    "kernel_share_of_step": s_ms / k_ms, "step_kernels_ms": k_ms, "bytes_per_site": BYTES_PER_SITE,
                    "peak_source": peak_src, "whole_step_GBps": step_gbs, "whole_step_frac": step_gbs / hbm_peak,
                    "sites_decided_by_screen": int(screened_out), "sites_to_fp64_kernel": int(n_fp64),
                    "fp64_kernel": {"kernel": "trio_list_kernel + trio_calls_finish_kernel (fused FP64 posterior, ordered compaction, call records)", "kernel_ms": f_ms, "sites": int(n_fp64),
                                    "flop_per_site": FLOP_PER_SITE, "achieved_tflops": ach_tf,
                                    "fp64_peak_tflops": fp64_peak,
                                    "peak_source": "dependent-free DFMA chains on all SMs, measured in this run"}}
    else:
        # screen disabled: the fused FP64 kernel sees every site (FP64-pipe-bound)
        ach_tf = FLOP_PER_SITE * n / (k_ms * 1e-3) / 1e12
        roofline = {"bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": ach_tf / fp64_peak if fp64_peak else None, "traffic": None,
                    "kernel": "trio_fast_kernel", "kernel_ms": k_ms, "kernel_share_of_step": min(1.0, k_ms / (ms_max / K)),
                    "flop_per_site": FLOP_PER_SITE,
                    "peak_source": "dependent-free DFMA chains on all SMs, measured in this run "
                                   "(MEASURED_PEAKS.json has no FP64 figure)"}

    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = 1_000_000
        v, cpu_calls = cpu_reference_run(sample, threads, faithful=True)
        v_hoist, _ = cpu_reference_run(sample, threads, faithful=False)
        cpu = {"value": v, "unit": "sites/s", "cores": threads, "kind": "port",
               "sample": "first %d sites of rank 0's shard, direct form with the reference's 3 log + 1 exp per "
                         "triple (C restatement; no JVM here)" % sample,
               "value_tables_hoisted": v_hoist}

    line = {"metric": METRIC, "value": value, "unit": "sites/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "READ-mode %s inference, %d synthetic trio candidate sites per GPU at 30x, eps 1e-2, "
                                   "mu 1e-8 (BASELINE configs[2]); interval-sharded; calls gathered to rank 0 %s"
                                   % (args.method, n, "by the kernels' peer stores over NVLink into rank 0's HBM + one release flag per "
                                      "rank and round (dnv_gather; no data collective)" if world > 1 else "(single GPU)"),
                       "sites_per_gpu": n, "l2": "input 240 MB per pass > 126 MB L2, and %d distinct input buffers in rotation (no flush needed)" % NBUF,
                       "calls_per_step": int(total_calls), "hard_sites_per_step": int(n_hard),
                       "step": "asynchronous batches (dnv_infer_counts_device_async / _gather) on ONE context: %s lanes inside the library, one dnv_sync per %d steps; kernel times from a serial, event-bracketed context" % (os.environ.get("DNV_LANES", "3"), GROUP),
                       "ms_per_step_synchronous_api": sync_s * 1e3, "per_rank_ms_per_step": [x / K for x in per_rank_ms],
                       "gather_verified": verify["verified"], "gather_verification": verify,
                       "warmup_steps_run": max(W, 2 * NBUF), "same_config_as_reference_arm": "the reference arm times a bounded 200k-site sample of this workload per step (CPU throughput per site does not depend on the batch size)"},
            "e2e": {"value": e2e_value, "unit": "sites/s", "h2d_bytes_per_step": n * (12 if u8 else 24),
                    "input_format": "uint8 counts, 12 B/site (dnv_infer_counts8; packed once outside the timed region — it is the caller's input format)" if u8 else "uint16 counts, 24 B/site",
                    "d2h_bytes_per_step": int(e2e_calls) * D.CALL_DT.itemsize + 64, "steps": Ke,
                    "timer": "host wall clock around dnv_infer_counts8 (pinned host buffers), max over ranks",
                    "limiter": "PCIe: one x16 Gen5 link per GPU; at N>1 every rank pulls its pinned buffer through the same host NUMA node"},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "aux": aux}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
