#!/usr/bin/env python
"""bench.py — trio candidate sites/sec, READ-mode BAYES inference at 30x (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]       our arm (one process per GPU under torchrun)
  python bench.py --impl reference ...                       reference arm: the reference's CPU algorithm
                                                            (C restatement, no JVM in this image) on host cores

One "step" = one pass of the hot path over one batch: packed pileup counts resident in HBM
-> compacted call records in HBM (+ for N>1 the NCCL gather of calls to rank 0).
Prints ONE JSON line on rank 0.
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
GROUP = 64  # asynchronous steps per host synchronisation (and per inter-GPU barrier); <= 64 per context
# ALGORITHMIC work per site:
#   integer screen (dominant kernel, HBM-bound): 24 B of counts in + 1 bit of decision out
#   list kernel (only the sites the screen cannot decide): sparse factored FP64 form, FMA = 2
BYTES_PER_SITE = 24.125
FLOP_PER_SITE = 451.0
# DRAM traffic of one screen launch over 10M sites from the ncu --set full capture committed as
# profiles/r01_read_pipeline_ncu.md (dram__bytes_read.sum + dram__bytes_write.sum), per site
NCU_TRAFFIC_BYTES_PER_SITE = 26.05


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
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:  # noqa: BLE001
                break
            time.sleep(0.0005)

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
                                   "(BASELINE configs[2]); %d-site sample per step" % sample},
            "cpu_baseline": {"value": value, "unit": "sites/s", "cores": threads, "kind": "port",
                             "sample": "%d sites/step x %d steps; C restatement of the reference's direct form "
                                       "(3 log + 1 exp per genotype triple, as DenovoBayesNet.java:220-233,269-273); "
                                       "no JVM in this image so the Java itself cannot run" % (sample, args.steps)},
            "e2e": {"value": value, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--sites", type=int, default=SITES_PER_GPU)
    ap.add_argument("--method", default="BAYES")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-screen", action="store_true", help="evaluate every site with the fused FP64 kernel")
    ap.add_argument("--contexts", type=int, default=4, help="contexts (each on its own stream) the batches are fed to in turn")
    ap.add_argument("--single-context", action="store_true", help="same as --contexts 1")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

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

    # eng: the context used for generation, kernel timing (event brackets) and the host-API leg;
    # pipe: `depth` untimed contexts, each on its own stream, fed in turn — the production way to walk a stream
    # of batches (the latency-bound list / finish kernels of one batch run beside the bandwidth-bound screen of
    # the next ones)
    eng = D.TrioEngine(EPS, MU, LRT, args.method, device=local, time_kernels=True, no_screen=args.no_screen)
    stream = torch.cuda.Stream(device=local)          # main stream: timing events, inter-GPU barrier
    eng.set_stream(stream.cuda_stream)
    depth = 1 if args.single_context else max(1, args.contexts)
    pstreams = [torch.cuda.Stream(device=local) for _ in range(depth)]
    pipe = D.BatchPipeline(depth=depth, streams=[st.cuda_stream for st in pstreams], seq_err_rate=EPS,
                           denovo_mut_rate=MU, lrt_threshold=LRT, method=args.method, device=local,
                           no_screen=args.no_screen)

    # ---- synthetic shard of this rank (interval sharding: rank r owns sites [r*n, (r+1)*n))
    d_counts = torch.empty((n, 12), dtype=torch.int16, device="cuda")
    eng.synth_counts_device(SEED, rank * n, n, d_counts, depth=DEPTH)
    cap = max(n // 100, 1024)
    d_calls = [torch.zeros((cap, D.CALL_DT.itemsize), dtype=torch.uint8, device="cuda") for _ in range(depth)]
    torch.cuda.synchronize()

    from denovo_b200 import sharding

    # N>1: the list kernel of every rank writes its call records and their count straight into rank 0's
    # HBM over NVLink (symmetric memory, one generation of slots per step of a group); no data
    # collective, one signal-pad barrier per group of GROUP steps.
    peer = None
    if world > 1:
        try:
            peer = sharding.PeerGather(dist, cap, D.CALL_DT.itemsize, slots=GROUP)
        except Exception as e:  # noqa: BLE001
            print("symmetric memory unavailable (%r): falling back to the NCCL gather" % (e,), file=sys.stderr)

    # One step = one batch through the asynchronous device entry point (dnv_infer_counts_device_async):
    # the whole pipeline is enqueued, the batch's call count lands in a pinned host slot (N=1) or straight
    # in rank 0's HBM next to the records (N>1), and the host only waits once per GROUP steps (dnv_sync).
    torch.cuda.set_stream(stream)
    h_nc = torch.zeros(64, dtype=torch.int64).pin_memory()
    counts_ptr = d_counts.data_ptr()
    calls_ptrs = [t.data_ptr() for t in d_calls]
    nc_ptrs = [h_nc.data_ptr() + 8 * i for i in range(64)]
    if peer:  # per-generation destinations in rank 0's HBM, computed once
        rec_ptrs = [peer.records_ptr(i) for i in range(GROUP)]
        cnt_ptrs = [peer.count_ptr(i) for i in range(GROUP)]

    def join():
        """The main stream waits for both pipeline streams (no host synchronisation)."""
        pipe.flush()
        for st in pstreams:
            stream.wait_stream(st)

    def run_steps(k_steps):
        """k_steps pipelined steps; returns summed library stats."""
        acc = {"n_launches": 0, "n_screened_out": 0, "n_hard": 0, "n_calls": 0, "n_sites": 0}
        done = 0
        while done < k_steps:
            m = min(GROUP, k_steps - done)
            if world > 1 and not peer:  # no symmetric memory: NCCL gather, needs the count on the host every step
                m = 1
                pipe.submit(counts_ptr, n, calls_ptrs[0], cap, nc_ptrs[0])
                st = pipe.sync()
                sharding.gather_records(d_calls[0][:int(h_nc[0])], dist)
            else:
                for i in range(m):
                    if peer:
                        pipe.submit(counts_ptr, n, rec_ptrs[i], cap, cnt_ptrs[i])
                    else:
                        pipe.submit(counts_ptr, n, calls_ptrs[i % depth], cap, nc_ptrs[i])
                if peer:
                    join()
                    peer.barrier()  # records + counts of every rank and step of the group now sit in rank 0's HBM
                st = pipe.sync()
            for key in acc:
                acc[key] += st[key]
            done += m
        return acc

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run_steps(max(W, 2 * depth))  # every context allocates its scratch on its first batch: all of them warm up
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if peer:
        peer.barrier()  # device-side line-up: every rank's timed region starts when all ranks are ready
    e0.record(stream)
    for st in pstreams:
        st.wait_event(e0)
    acc = run_steps(K)
    join()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = acc["n_launches"]
    screened_out = acc["n_screened_out"] // K
    n_hard = acc["n_hard"] // K
    if peer:
        total_calls = sum(peer.counts(gen=(K - 1) % GROUP)) if rank == 0 else 0
    else:
        total_calls = acc["n_calls"] // K
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n * K / (ms_max * 1e-3)

    # kernel times (CUDA-event brackets inside the library, DNV_F_TIME_KERNELS) and the same step through
    # the synchronous entry point (the host reads n_calls after every batch), on the timed context
    Kt = 20
    for _ in range(3):
        eng.infer_device(counts_ptr, n, calls_ptrs[0], cap)
    fused_ms = screen_ms = 0.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(Kt):
        eng.infer_device(counts_ptr, n, calls_ptrs[0], cap)
        st = eng.stats()
        fused_ms += st["fused_kernel_ms"]
        screen_ms += st["screen_kernel_ms"]
    sync_ms = (time.perf_counter() - t0) * 1e3 / Kt
    fused_ms, screen_ms = fused_ms * K / Kt, screen_ms * K / Kt  # the roofline code below divides by K

    # ---- end to end through the host API: pinned host counts -> calls in pinned host memory
    # the host packs counts as uint8 (12 B/site) when every per-allele depth is below 256, else uint16
    u8 = bool(int(d_counts.max().item()) < 256)
    if u8:
        h_counts = torch.empty((n, 12), dtype=torch.uint8).pin_memory()
        h_counts.copy_(d_counts.to(torch.uint8))
    else:
        h_counts = torch.empty((n, 12), dtype=torch.int16).pin_memory()
        h_counts.copy_(d_counts)
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel
    fp64_peak, _ = eng.measure_fp64_peak()
    hbm_peak, peak_src = load_peaks()
    k_ms = fused_ms / K          # the kernels of a step: integer screen + list kernel + finish kernel
    s_ms = screen_ms / K         # the screen alone
    if s_ms > 0:
        # BAYES/LRT: the integer screen streams every site's counts and is the dominant kernel (HBM-bound)
        n_fp64 = n - screened_out
        ach_gbs = BYTES_PER_SITE * n / (s_ms * 1e-3) / 1e9
        f_ms = max(k_ms - s_ms, 1e-9)
        ach_tf = FLOP_PER_SITE * n_fp64 / (f_ms * 1e-3) / 1e12
        roofline = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                    "traffic": NCU_TRAFFIC_BYTES_PER_SITE * n, "kernel": "trio_screen_kernel", "kernel_ms": s_ms,
                    # share of the step's kernel time (screen + list + finish kernel, timed alone on the timed
                    # context: comparable with the serialised ncu launch list); the value loop overlaps the
                    # list kernel of one batch with the screen of the next, so ms_per_step < their sum
                    "kernel_share_of_step": s_ms / k_ms, "step_kernels_ms": k_ms, "bytes_per_site": BYTES_PER_SITE,
                    "peak_source": peak_src,
                    "sites_decided_by_screen": int(screened_out), "sites_to_fp64_kernel": int(n_fp64),
                    "fp64_kernel": {"kernel": "trio_list_kernel + trio_calls_finish_kernel (fused FP64 posterior, ordered compaction, call records)", "kernel_ms": f_ms, "sites": int(n_fp64),
                                    "flop_per_site": FLOP_PER_SITE, "achieved_tflops": ach_tf,
                                    "fp64_peak_tflops": fp64_peak,
                                    "peak_source": "dependent-free DFMA chains on all SMs, measured in this run"}}
    else:
        # MAP (or screen disabled): the fused FP64 kernel sees every site (FP64-pipe-bound)
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
                                   % (args.method, n, "by peer stores over NVLink from the list kernel + one signal-pad "
                                      "barrier per group of steps (no data collective)" if peer else ("with NCCL" if world > 1 else "(single GPU)")),
                       "sites_per_gpu": n, "l2": "input 240 MB per pass > 126 MB L2 (no flush needed)",
                       "calls_per_step": int(total_calls), "hard_sites_per_step": int(n_hard),
                       "step": "asynchronous batches (dnv_infer_counts_device_async) fed alternately to %d contexts on %d streams, one dnv_sync per %d steps; kernel times from a separate timed context" % (depth, depth, GROUP),
                       "ms_per_step_synchronous_api": sync_ms, "contexts": depth,
                       "warmup_steps_run": max(W, 2 * depth)},
            "e2e": {"value": e2e_value, "unit": "sites/s", "h2d_bytes_per_step": n * (12 if u8 else 24),
                    "input_format": "uint8 counts, 12 B/site (dnv_infer_counts8)" if u8 else "uint16 counts, 24 B/site",
                    "d2h_bytes_per_step": int(e2e_calls) * D.CALL_DT.itemsize + 64, "steps": Ke,
                    "timer": "host wall clock around dnv_infer_counts (pinned host buffers), max over ranks"},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
