"""Host-side mirror of the reference's per-site engine seam, batched.

  reference                                        here
  ------------------------------------------------ ----------------------------------
  new BayesInfer(shared)  (M/BayesInfer.java:39-42) TrioEngine(seq_err_rate=..., ...)
  BayesInfer.infer(map, method) per site (:50-69)   TrioEngine.infer(counts)  per batch
  new ReadSummary(reads, pos) (M/ReadSummary.java)  TrioEngine.pileup(...)
  VariantsBuffer + PositionCall (M/VariantsBuffer)  TrioEngine.mendel_filter(calls)

  shared PrintWriter of N per-site tasks            Gather (one process per GPU) / DeviceGroup (one process)
    (M/ReadCaller.java:91-95,122-129)

Every method goes through the C ABI (include/denovo_b200.h); host arrays are numpy,
device arrays are torch CUDA tensors (torch is used only to own device memory).  Batching — the overlap of
consecutive batches, the interval shards, the gather — lives below the ABI; this file only forwards.
"""
import ctypes as C
import math

import numpy as np

from . import _lib as L

GENOTYPES = ["AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT"]
ALLELES = "ACGT"


def poisson_thresholds(depth, kmax=None):
    """2^64-scaled ascending Poisson CDF thresholds (depth = #{k: thr[k] <= u}).
    Plain IEEE double recurrence in a fixed operation order, so the Python host, the C++
    host (host/denovo_host.hpp: poisson_thresholds) and the tests build identical integers."""
    import math
    depth = float(depth)
    if kmax is None:
        kmax = int(depth + 12.0 * math.sqrt(depth) + 20.0)
    term, cdf, thr = math.exp(-depth), 0.0, []
    for k in range(kmax):
        cdf += term
        x = cdf * 18446744073709551616.0
        thr.append((1 << 64) - 1 if x >= 18446744073709551616.0 else int(x))
        term = term * depth / (k + 1)
    return np.array(thr, np.uint64)


class TrioEngine:
    """One CUDA device, one set of (seq_err_rate, denovo_mut_rate, lrt_threshold, method)."""

    def __init__(self, seq_err_rate=1e-2, denovo_mut_rate=1e-8, lrt_threshold=1.0, method="BAYES", device=0,
                 force_direct=False, time_kernels=False, no_screen=False):
        self.lib = L.load()
        m = L.METHODS[method.upper()] if isinstance(method, str) else int(method)
        self.params = L.Params(seq_err_rate, denovo_mut_rate, lrt_threshold, m, device,
                               (L.F_FORCE_DIRECT if force_direct else 0) | (L.F_TIME_KERNELS if time_kernels else 0)
                               | (L.F_NO_SCREEN if no_screen else 0), 0)
        self.method = m
        self.device = device
        h = C.c_void_p()
        L.check(self.lib.dnv_create(C.byref(self.params), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.dnv_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- tables / stats
    def tables(self):
        rel, bl, isden = np.zeros(1000), np.zeros(40), np.zeros(1000, np.uint8)
        L.check(self.lib.dnv_get_tables(self._h, L.ptr(rel), L.ptr(bl), L.ptr(isden)))
        return {"R": rel, "BL": bl.reshape(10, 4), "isden": isden}

    def stats(self):
        s = L.Stats()
        L.check(self.lib.dnv_get_stats(self._h, C.byref(s)))
        return {"n_sites": s.n_sites, "n_calls": s.n_calls, "n_hard": s.n_hard, "n_launches": s.n_launches,
                "fused_kernel_ms": s.fused_kernel_ms, "screen_kernel_ms": s.screen_kernel_ms,
                "n_screened_out": s.n_screened_out}

    def set_stream(self, cuda_stream):
        """Run on the caller's CUDA stream (int handle, e.g. torch.cuda.Stream().cuda_stream); None restores."""
        L.check(self.lib.dnv_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    # ---- READ mode
    def infer(self, counts, dense=False, cap=None):
        """counts: host uint16 [n,3,4] (DAD,MOM,CHILD x A,C,G,T); a uint8 array selects the compact
        12-byte-per-site entry point.  Returns (calls CALL_DT[], dense SITE_DT[] or None)."""
        u8 = isinstance(counts, np.ndarray) and counts.dtype == np.uint8
        counts = np.ascontiguousarray(counts, np.uint8 if u8 else np.uint16).reshape(-1, 12)
        n = counts.shape[0]
        cap = n if cap is None else cap
        calls = np.zeros(max(cap, 1), L.CALL_DT)
        d = np.zeros(n, L.SITE_DT) if dense else None
        nc = C.c_int64(0)
        fn = self.lib.dnv_infer_counts8 if u8 else self.lib.dnv_infer_counts
        L.check(fn(self._h, L.ptr(counts), n, L.ptr(d) if dense else None,
                   L.ptr(calls) if cap > 0 else None, cap, C.byref(nc)))
        return calls[:nc.value], d

    def infer_host_into(self, counts_ptr, n, calls_ptr, cap, u8=False):
        """Zero-allocation host call (pinned buffers owned by the caller); returns n_calls."""
        nc = C.c_int64(0)
        fn = self.lib.dnv_infer_counts8 if u8 else self.lib.dnv_infer_counts
        L.check(fn(self._h, L.ptr(counts_ptr), n, None, L.ptr(calls_ptr), cap, C.byref(nc)))
        return nc.value

    def infer_device(self, d_counts, n, d_calls, cap, d_dense=None, u8=False):
        """Device-resident buffers (torch CUDA tensors or raw pointers); returns n_calls."""
        nc = C.c_int64(0)
        fn = self.lib.dnv_infer_counts8_device if u8 else self.lib.dnv_infer_counts_device
        L.check(fn(self._h, L.ptr(d_counts), n, L.ptr(d_dense), L.ptr(d_calls), cap, C.byref(nc)))
        return nc.value

    def infer_device_async(self, d_counts, n, d_calls, cap, h_n_calls, u8=False, site_base=0):
        """Enqueue one batch without waiting; h_n_calls = pinned host int64 slot (tensor / array / address)
        that receives the batch's call count in stream order.  site_base is added to the records' site
        indices (interval shards).  Pair with flush() / sync()."""
        fn = self.lib.dnv_infer_counts8_device_async_at if u8 else self.lib.dnv_infer_counts_device_async_at
        L.check(fn(self._h, L.ptr(d_counts), n, site_base, L.ptr(d_calls), cap, L.ptr(h_n_calls)))

    def flush(self):
        """Device-side join: the context's stream waits for every enqueued batch (no host wait)."""
        L.check(self.lib.dnv_flush(self._h))

    def sync(self):
        """Wait for every enqueued batch; stats() then holds the sums over those batches."""
        L.check(self.lib.dnv_sync(self._h))

    def pileup(self, read_start, base_off, bases, person, site_pos):
        read_start = np.ascontiguousarray(read_start, np.int64)
        base_off = np.ascontiguousarray(base_off, np.int64)
        bases = np.ascontiguousarray(bases, np.uint8)
        person = np.ascontiguousarray(person, np.uint8)
        site_pos = np.ascontiguousarray(site_pos, np.int64)
        out = np.zeros((site_pos.shape[0], 3, 4), np.uint16)
        L.check(self.lib.dnv_pileup(self._h, L.ptr(read_start), L.ptr(base_off), L.ptr(bases), L.ptr(person),
                                    read_start.shape[0], L.ptr(site_pos), site_pos.shape[0], L.ptr(out)))
        return out

    def pileup_device(self, d_read_start, d_base_off, d_bases, d_person, n_reads, d_site_pos, n_sites, d_counts):
        L.check(self.lib.dnv_pileup_device(self._h, L.ptr(d_read_start), L.ptr(d_base_off), L.ptr(d_bases),
                                           L.ptr(d_person), n_reads, L.ptr(d_site_pos), n_sites, L.ptr(d_counts)))

    def infer_reads(self, read_start, base_off, bases, person, site_pos, cap=None):
        """reads -> calls in one fused pass (dnv_infer_reads): the records of infer(pileup(...))."""
        read_start = np.ascontiguousarray(read_start, np.int64)
        base_off = np.ascontiguousarray(base_off, np.int64)
        bases = np.ascontiguousarray(bases, np.uint8)
        person = np.ascontiguousarray(person, np.uint8)
        site_pos = np.ascontiguousarray(site_pos, np.int64)
        n = site_pos.shape[0]
        cap = n if cap is None else cap
        calls = np.zeros(max(cap, 1), L.CALL_DT)
        nc = C.c_int64(0)
        L.check(self.lib.dnv_infer_reads(self._h, L.ptr(read_start), L.ptr(base_off), L.ptr(bases), L.ptr(person),
                                         read_start.shape[0], L.ptr(site_pos), n, L.ptr(calls) if cap > 0 else None, cap,
                                         C.byref(nc)))
        return calls[:nc.value]

    def infer_reads_device(self, d_read_start, d_base_off, d_bases, d_person, n_reads, d_site_pos, n_sites, d_calls, cap):
        nc = C.c_int64(0)
        L.check(self.lib.dnv_infer_reads_device(self._h, L.ptr(d_read_start), L.ptr(d_base_off), L.ptr(d_bases), L.ptr(d_person),
                                                n_reads, L.ptr(d_site_pos), n_sites, L.ptr(d_calls), cap, C.byref(nc)))
        return nc.value

    # ---- VARIANT mode
    def mendel_filter(self, calls, cap=None):
        calls = np.ascontiguousarray(calls, L.VCALL_DT)
        n = calls.shape[0]
        cap = n if cap is None else cap
        out = np.zeros(max(cap, 1), L.PCALL_DT)
        no = C.c_int64(0)
        L.check(self.lib.dnv_mendel_filter(self._h, L.ptr(calls), n, L.ptr(out) if cap > 0 else None, cap,
                                           C.byref(no)))
        return out[:no.value]

    def mendel_filter_device(self, d_calls, n, d_out, cap):
        no = C.c_int64(0)
        L.check(self.lib.dnv_mendel_filter_device(self._h, L.ptr(d_calls), n, L.ptr(d_out), cap, C.byref(no)))
        return no.value

    # ---- synthetic source + calibration
    def synth_counts_device(self, seed, first, n, d_counts, depth=30, seq_err=None, thr=None):
        thr = poisson_thresholds(depth) if thr is None else thr
        err_thr = int((self.params.seq_err_rate if seq_err is None else seq_err) * (1 << 32))
        L.check(self.lib.dnv_synth_counts_device(self._h, seed, first, n, L.ptr(thr), thr.shape[0], err_thr,
                                                 L.ptr(d_counts)))

    def synth_counts_at(self, seed, index, depth=30, seq_err=None, thr=None):
        """Host arrays: counts of the synthetic trio at arbitrary global site indices."""
        thr = poisson_thresholds(depth) if thr is None else thr
        index = np.ascontiguousarray(index, np.int64)
        err_thr = int((self.params.seq_err_rate if seq_err is None else seq_err) * (1 << 32))
        out = np.zeros((index.shape[0], 3, 4), np.uint16)
        L.check(self.lib.dnv_synth_counts_at(self._h, seed, L.ptr(index), index.shape[0], L.ptr(thr), thr.shape[0],
                                             err_thr, L.ptr(out)))
        return out

    def measure_fp64_peak(self):
        tf, ms = C.c_double(0), C.c_double(0)
        L.check(self.lib.dnv_measure_fp64_peak(self._h, C.byref(tf), C.byref(ms)))
        return tf.value, ms.value

    def measure_screen(self, d_counts, n, reps=20, u8=False):
        """Average duration (ms) of `reps` back-to-back launches of the integer screen alone over device counts."""
        ms = C.c_double(0)
        L.check(self.lib.dnv_measure_screen(self._h, L.ptr(d_counts), 12 if u8 else 24, n, reps, C.byref(ms)))
        return ms.value


class Gather:
    """One rank's end of the gather of call records to rank `dst` (include/denovo_b200.h: dnv_gather).
    The records and counts are stored by the producing kernels straight into the destination's HBM; this class
    only owns the handle.  Between processes the destination's 64-byte handle has to reach the other ranks:
    `exchange(dist)` does that with one torch.distributed broadcast (plumbing, any backend)."""

    def __init__(self, engine, world, rank, cap, gens=1, dst=0):
        self.eng, self.lib = engine, engine.lib
        self.world, self.rank, self.dst, self.cap, self.gens = world, rank, dst, int(cap), int(gens)
        h = C.c_void_p()
        L.check(self.lib.dnv_gather_create(engine._h, world, rank, dst, self.cap, self.gens, C.byref(h)))
        self._h = h

    def export_handle(self):
        buf = (C.c_ubyte * 64)()
        L.check(self.lib.dnv_gather_export(self._h, buf))
        return bytes(buf)

    def import_handle(self, handle):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        L.check(self.lib.dnv_gather_import(self._h, buf))

    def attach(self, dst_end):
        L.check(self.lib.dnv_gather_attach(self._h, dst_end._h))

    def exchange(self, dist):
        """One process per GPU: broadcast the destination's handle and map it on the other ranks."""
        import torch
        t = torch.zeros(64, dtype=torch.uint8)
        if self.rank == self.dst:
            t = torch.frombuffer(bytearray(self.export_handle()), dtype=torch.uint8).clone()
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = t.to(dev)
        dist.broadcast(t, src=self.dst)
        if self.rank != self.dst:
            self.import_handle(bytes(t.cpu().numpy().tobytes()))

    def submit(self, d_counts, n, site_base, gen, u8=False):
        """Enqueue one batch whose records (global site indices) and count land in this rank's slot of `gen`."""
        fn = self.lib.dnv_infer_counts8_device_gather if u8 else self.lib.dnv_infer_counts_device_gather
        L.check(fn(self.eng._h, L.ptr(d_counts), n, site_base, self._h, gen))

    def arrive(self):
        L.check(self.lib.dnv_gather_arrive(self._h))

    def wait(self):
        L.check(self.lib.dnv_gather_wait(self._h))

    def fetch(self, gen=0, cap=None):
        """Destination: the gathered records of one generation (rank order = genomic order) and per-rank counts."""
        cap = self.world * self.cap if cap is None else cap
        out = np.zeros(max(cap, 1), L.CALL_DT)
        per = np.zeros(self.world, np.int64)
        n = C.c_int64(0)
        L.check(self.lib.dnv_gather_fetch(self._h, gen, L.ptr(out), cap, C.byref(n), L.ptr(per)))
        return out[:n.value], per

    def close(self):
        if getattr(self, "_h", None):
            self.lib.dnv_gather_destroy(self._h)
            self._h = None

    __del__ = close


class DeviceGroup:
    """N devices driven by this one process (dnv_group): host batches are split by interval, every device infers
    its shard, the records come back gathered in input order.  What the single-process hosts (CLI, Java) use."""

    def __init__(self, devices, seq_err_rate=1e-2, denovo_mut_rate=1e-8, lrt_threshold=1.0, method="BAYES"):
        self.lib = L.load()
        m = L.METHODS[method.upper()] if isinstance(method, str) else int(method)
        prm = L.Params(seq_err_rate, denovo_mut_rate, lrt_threshold, m, 0, 0, 0)
        dev = np.ascontiguousarray(devices, np.int32)
        h = C.c_void_p()
        L.check(self.lib.dnv_group_create(C.byref(prm), L.ptr(dev), len(dev), C.byref(h)))
        self._h = h

    def infer(self, counts, cap=None):
        u8 = isinstance(counts, np.ndarray) and counts.dtype == np.uint8
        counts = np.ascontiguousarray(counts, np.uint8 if u8 else np.uint16).reshape(-1, 12)
        n = counts.shape[0]
        cap = n if cap is None else cap
        calls = np.zeros(max(cap, 1), L.CALL_DT)
        nc = C.c_int64(0)
        fn = self.lib.dnv_group_infer_counts8 if u8 else self.lib.dnv_group_infer_counts
        L.check(fn(self._h, L.ptr(counts), n, L.ptr(calls) if cap > 0 else None, cap, C.byref(nc)))
        return calls[:nc.value]

    def infer_into(self, counts_ptr, n, calls_ptr, cap, u8=False):
        """Zero-allocation call over caller-owned (pinned) buffers; returns n_calls."""
        nc = C.c_int64(0)
        fn = self.lib.dnv_group_infer_counts8 if u8 else self.lib.dnv_group_infer_counts
        L.check(fn(self._h, L.ptr(counts_ptr), n, L.ptr(calls_ptr), cap, C.byref(nc)))
        return nc.value

    def stats(self):
        s = L.Stats()
        L.check(self.lib.dnv_group_get_stats(self._h, C.byref(s)))
        return {"n_sites": s.n_sites, "n_calls": s.n_calls, "n_hard": s.n_hard, "n_launches": s.n_launches,
                "n_screened_out": s.n_screened_out}

    def close(self):
        if getattr(self, "_h", None):
            self.lib.dnv_group_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
