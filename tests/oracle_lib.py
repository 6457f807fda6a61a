"""ctypes loader for the CPU oracle (oracle/liboracle.so).  Test infrastructure:
only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GENOTYPES = ["AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT"]

RESULT_DT = np.dtype([("dad", "<i4"), ("mom", "<i4"), ("child", "<i4"),
                      ("denovo_map", "<i4"), ("denovo_bayes", "<i4"), ("denovo_lrt", "<i4"),
                      ("max_ll", "<f8"), ("p", "<f8"), ("lr", "<f8"),
                      ("log_men", "<f8"), ("log_den", "<f8"), ("l_den", "<f8"), ("l_men", "<f8")])
VCALL_DT = np.dtype([("start", "<i8"), ("end", "<i8"), ("variant_id", "<i4"),
                     ("gt0", "i1"), ("gt1", "i1"), ("ngt", "u1"), ("person", "u1"),
                     ("filter_pass", "u1"), ("ref_len", "u1"), ("ref_base", "S1"),
                     ("alt0_len", "u1"), ("alt0_base", "S1"), ("alt1_len", "u1"),
                     ("alt1_base", "S1"), ("pad_", "u1", (1,))])
PCALL_DT = np.dtype([("position", "<i8"), ("child_index", "<i8"), ("child", "u1"),
                     ("mom", "u1"), ("dad", "u1"), ("is_denovo", "u1"), ("pad_", "u1", (4,))])
assert RESULT_DT.itemsize == 80 and VCALL_DT.itemsize == 32 and PCALL_DT.itemsize == 24


class Params(C.Structure):
    _fields_ = [("seq_err_rate", C.c_double), ("denovo_mut_rate", C.c_double),
                ("lrt_threshold", C.c_double)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        src_m = max(os.path.getmtime(os.path.join(ORACLE_DIR, f))
                    for f in ("denovo_oracle.c", "denovo_oracle.h", "synth_ref.c"))
        if not os.path.exists(so) or os.path.getmtime(so) < src_m:
            build()
        _lib = C.CDLL(so)
        _lib.orc_infer.restype = C.c_int
        _lib.orc_pileup.restype = C.c_int64
        _lib.orc_mendel_stream.restype = C.c_int64
        _lib.orc_pileup_sorted.restype = C.c_int64
        _lib.synth_ref_reads_at.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def params(seq_err=1e-2, mut=1e-8, lrt=1.0):
    return Params(seq_err, mut, lrt)


def tables(prm):
    L = lib()
    isden = np.zeros(1000, np.uint8)
    P = np.zeros(10)
    T = np.zeros(1000)
    BL = np.zeros(40)
    R = np.zeros(1000)
    L.orc_is_denovo_table(_p(isden))
    L.orc_parent_prior(_p(P))
    L.orc_child_cpt(C.byref(prm), _p(T))
    L.orc_base_loglik(C.byref(prm), _p(BL))
    L.orc_relationship_loglik(C.byref(prm), _p(R))
    return {"isden": isden, "P": P, "T": T, "BL": BL.reshape(10, 4), "R": R}


def read_loglik(prm, counts4):
    c = np.ascontiguousarray(counts4, np.uint16)
    out = np.zeros(10)
    lib().orc_read_loglik(C.byref(prm), _p(c), _p(out))
    return out


def infer(prm, counts, faithful=False, threads=1):
    """counts: [n,3,4] uint16 (DAD, MOM, CHILD × A,C,G,T) -> RESULT_DT[n]"""
    c = np.ascontiguousarray(counts, np.uint16).reshape(-1, 12)
    out = np.zeros(c.shape[0], RESULT_DT)
    rc = lib().orc_infer(C.byref(prm), _p(c), C.c_int64(c.shape[0]), _p(out),
                         C.c_int(int(faithful)), C.c_int(threads))
    assert rc == 0
    return out


def pileup(read_start, base_off, bases, person, site_pos):
    read_start = np.ascontiguousarray(read_start, np.int64)
    base_off = np.ascontiguousarray(base_off, np.int64)
    bases = np.ascontiguousarray(bases, np.uint8)
    person = np.ascontiguousarray(person, np.uint8)
    site_pos = np.ascontiguousarray(site_pos, np.int64)
    out = np.zeros((site_pos.shape[0], 3, 4), np.uint16)
    rc = lib().orc_pileup(_p(read_start), _p(base_off), _p(bases), _p(person),
                          C.c_int64(read_start.shape[0]), _p(site_pos),
                          C.c_int64(site_pos.shape[0]), _p(out))
    return rc, out


def pileup_sorted(read_start, base_off, bases, person, site_pos):
    """orc_pileup for reads sorted by (person, start): binary-searched windows, for big inputs."""
    read_start = np.ascontiguousarray(read_start, np.int64)
    base_off = np.ascontiguousarray(base_off, np.int64)
    bases = np.ascontiguousarray(bases, np.uint8)
    person = np.ascontiguousarray(person, np.uint8)
    site_pos = np.ascontiguousarray(site_pos, np.int64)
    out = np.zeros((site_pos.shape[0], 3, 4), np.uint16)
    rc = lib().orc_pileup_sorted(_p(read_start), _p(base_off), _p(bases), _p(person), C.c_int64(read_start.shape[0]),
                                 _p(site_pos), C.c_int64(site_pos.shape[0]), _p(out))
    return rc, out


def synth_reads_at(seed, index, depth=30, seq_err=1e-2, read_len=100, gap_rate=1e-3, thr=None):
    """Synthetic reads covering the candidate sites `index` (global positions): the specification of the host's
    read source.  Returns read_start, base_off, bases (uint8 ASCII), person, sorted by (person, start)."""
    if thr is None:
        thr = poisson_thresholds(depth)
    index = np.ascontiguousarray(index, np.int64)
    err_thr, gap_thr = int(seq_err * (1 << 32)), int(gap_rate * (1 << 32))
    args = (C.c_uint64(seed), _p(index), C.c_int64(index.shape[0]), _p(thr), C.c_int32(thr.shape[0]), C.c_uint32(err_thr),
            C.c_int32(read_len), C.c_uint32(gap_thr))
    n = lib().synth_ref_reads_at(*args, None, None, None, None)
    rs, bo = np.zeros(n, np.int64), np.zeros(n + 1, np.int64)
    ba, pe = np.zeros(n * read_len, np.uint8), np.zeros(n, np.uint8)
    lib().synth_ref_reads_at(*args, _p(rs), _p(bo), _p(ba), _p(pe))
    return rs, bo, ba, pe


def mendel_stream(calls):
    calls = np.ascontiguousarray(calls, VCALL_DT)
    cap = int((calls["person"] == 0).sum()) + 1
    out = np.zeros(cap, PCALL_DT)
    n = lib().orc_mendel_stream(_p(calls), C.c_int64(calls.shape[0]), _p(out), C.c_int64(cap))
    if n < 0:
        return n, out[:0]
    return n, out[:n]


def mendel_admit(calls):
    calls = np.ascontiguousarray(calls, VCALL_DT)
    out = np.zeros(calls.shape[0], np.uint8)
    lib().orc_mendel_admit(_p(calls), C.c_int64(calls.shape[0]), _p(out))
    return out


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


def synth_counts(seed, first, n, depth=30, seq_err=1e-2, thr=None):
    if thr is None:
        thr = poisson_thresholds(depth)
    err_thr = int(seq_err * (1 << 32))
    out = np.zeros((n, 3, 4), np.uint16)
    lib().synth_ref_counts(C.c_uint64(seed), C.c_int64(first), C.c_int64(n), _p(thr),
                           C.c_int32(thr.shape[0]), C.c_uint32(err_thr), _p(out))
    return out


def synth_counts_at(seed, index, depth=30, seq_err=1e-2, thr=None):
    if thr is None:
        thr = poisson_thresholds(depth)
    index = np.ascontiguousarray(index, np.int64)
    err_thr = int(seq_err * (1 << 32))
    out = np.zeros((index.shape[0], 3, 4), np.uint16)
    lib().synth_ref_counts_at(C.c_uint64(seed), _p(index), C.c_int64(index.shape[0]), _p(thr),
                              C.c_int32(thr.shape[0]), C.c_uint32(err_thr), _p(out))
    return out
