"""ctypes binding of include/denovo_b200.h (libdenovo_b200.so).

The library is CUDA-only: loading works anywhere the CUDA runtime is linked
statically, but `dnv_create` fails with DNV_E_CUDA when no device is present —
there is no CPU fallback and nothing here imports the test oracle.
"""
import ctypes as C
import os
import subprocess

import numpy as np

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(PKG_ROOT, "csrc")
LIB_PATH = os.environ.get("DNV_B200_LIB") or os.path.join(PKG_ROOT, "lib", "libdenovo_b200.so")

DNV_OK, DNV_E_INVALID, DNV_E_CUDA, DNV_E_CAPACITY, DNV_E_BADBASE, DNV_E_UNSORTED, DNV_E_NOMEM = 0, -1, -2, -3, -4, -5, -6
METHODS = {"MAP": 0, "BAYES": 1, "LRT": 2}
F_FORCE_DIRECT = 1
F_TIME_KERNELS = 2
F_NO_SCREEN = 4

EXPORTS = ["dnv_create", "dnv_destroy", "dnv_last_error", "dnv_abi_version", "dnv_get_tables", "dnv_get_stats",
           "dnv_set_stream",
           "dnv_infer_counts", "dnv_infer_counts_device", "dnv_infer_counts8", "dnv_infer_counts8_device",
           "dnv_infer_counts_device_async", "dnv_infer_counts8_device_async",
           "dnv_infer_counts_device_async_at", "dnv_infer_counts8_device_async_at", "dnv_flush", "dnv_sync",
           "dnv_alloc_pinned", "dnv_free_pinned",
           "dnv_gather_create", "dnv_gather_destroy", "dnv_gather_export", "dnv_gather_import", "dnv_gather_attach",
           "dnv_gather_slot", "dnv_infer_counts_device_gather", "dnv_infer_counts8_device_gather", "dnv_gather_arrive",
           "dnv_gather_wait", "dnv_gather_fetch", "dnv_group_create", "dnv_group_destroy", "dnv_group_infer_counts",
           "dnv_group_infer_counts8", "dnv_group_size", "dnv_group_ctx", "dnv_group_get_stats", "dnv_pileup", "dnv_pileup_device", "dnv_infer_reads", "dnv_infer_reads_device",
           "dnv_mendel_filter", "dnv_mendel_filter_device", "dnv_synth_counts_device", "dnv_synth_counts_at_device",
           "dnv_synth_counts_at", "dnv_measure_fp64_peak", "dnv_measure_screen"]


class Params(C.Structure):
    _fields_ = [("seq_err_rate", C.c_double), ("denovo_mut_rate", C.c_double), ("lrt_threshold", C.c_double),
                ("method", C.c_int32), ("device", C.c_int32), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("n_sites", C.c_int64), ("n_calls", C.c_int64), ("n_hard", C.c_int64), ("n_launches", C.c_int64),
                ("fused_kernel_ms", C.c_double), ("screen_kernel_ms", C.c_double), ("n_screened_out", C.c_int64)]


CALL_DT = np.dtype([("site", "<i8"), ("counts", "<u2", (12,)), ("dad", "u1"), ("mom", "u1"), ("child", "u1"),
                    ("is_denovo", "u1"), ("flags", "<u4"), ("max_ll", "<f8"), ("p", "<f8"), ("lr", "<f8"),
                    ("log_men", "<f8"), ("log_den", "<f8")])
SITE_DT = np.dtype([("p", "<f8"), ("lr", "<f8"), ("log_men", "<f8"), ("log_den", "<f8"), ("max_ll", "<f8"),
                    ("dad", "u1"), ("mom", "u1"), ("child", "u1"), ("is_denovo", "u1"), ("flags", "<u4")])
VCALL_DT = np.dtype([("start", "<i8"), ("end", "<i8"), ("variant_id", "<i4"), ("gt0", "i1"), ("gt1", "i1"),
                     ("ngt", "u1"), ("person", "u1"), ("filter_pass", "u1"), ("ref_len", "u1"), ("ref_base", "S1"),
                     ("alt0_len", "u1"), ("alt0_base", "S1"), ("alt1_len", "u1"), ("alt1_base", "S1"),
                     ("pad_", "u1", (1,))])
PCALL_DT = np.dtype([("position", "<i8"), ("child_index", "<i8"), ("child", "u1"), ("mom", "u1"), ("dad", "u1"),
                     ("is_denovo", "u1"), ("pad_", "u1", (4,))])
assert CALL_DT.itemsize == 80 and SITE_DT.itemsize == 48 and VCALL_DT.itemsize == 32 and PCALL_DT.itemsize == 24


class DenovoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("denovo_b200 error %d: %s" % (code, msg))
        self.code = code


def build(verbose=False):
    """Compile every CUDA source for sm_100a into lib/libdenovo_b200.so (in-tree)."""
    subprocess.check_call(["make", "-C", CSRC] + ([] if verbose else ["-s"]))
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, u32, u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_uint64
    L.dnv_last_error.restype = C.c_char_p
    L.dnv_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.dnv_destroy.argtypes = [vp]
    L.dnv_get_tables.argtypes = [vp, vp, vp, vp]
    L.dnv_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.dnv_set_stream.argtypes = [vp, vp]
    L.dnv_infer_counts.argtypes = [vp, vp, i64, vp, vp, i64, C.POINTER(i64)]
    L.dnv_infer_counts_device.argtypes = [vp, vp, i64, vp, vp, i64, C.POINTER(i64)]
    L.dnv_infer_counts8.argtypes = [vp, vp, i64, vp, vp, i64, C.POINTER(i64)]
    L.dnv_infer_counts8_device.argtypes = [vp, vp, i64, vp, vp, i64, C.POINTER(i64)]
    L.dnv_infer_counts_device_async.argtypes = [vp, vp, i64, vp, i64, vp]
    L.dnv_infer_counts8_device_async.argtypes = [vp, vp, i64, vp, i64, vp]
    L.dnv_infer_counts_device_async_at.argtypes = [vp, vp, i64, i64, vp, i64, vp]
    L.dnv_infer_counts8_device_async_at.argtypes = [vp, vp, i64, i64, vp, i64, vp]
    L.dnv_flush.argtypes = [vp]
    L.dnv_sync.argtypes = [vp]
    L.dnv_gather_create.argtypes = [vp, i32, i32, i32, i64, i32, C.POINTER(vp)]
    L.dnv_gather_destroy.argtypes = [vp]
    L.dnv_gather_export.argtypes = [vp, vp]
    L.dnv_gather_import.argtypes = [vp, vp]
    L.dnv_gather_attach.argtypes = [vp, vp]
    L.dnv_gather_slot.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(vp)]
    L.dnv_infer_counts_device_gather.argtypes = [vp, vp, i64, i64, vp, i32]
    L.dnv_infer_counts8_device_gather.argtypes = [vp, vp, i64, i64, vp, i32]
    L.dnv_gather_arrive.argtypes = [vp]
    L.dnv_gather_wait.argtypes = [vp]
    L.dnv_gather_fetch.argtypes = [vp, i32, vp, i64, C.POINTER(i64), vp]
    L.dnv_group_create.argtypes = [C.POINTER(Params), vp, i32, C.POINTER(vp)]
    L.dnv_group_destroy.argtypes = [vp]
    L.dnv_group_infer_counts.argtypes = [vp, vp, i64, vp, i64, C.POINTER(i64)]
    L.dnv_group_infer_counts8.argtypes = [vp, vp, i64, vp, i64, C.POINTER(i64)]
    L.dnv_group_size.argtypes = [vp]
    L.dnv_group_ctx.argtypes = [vp, i32, C.POINTER(vp)]
    L.dnv_group_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.dnv_alloc_pinned.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.dnv_free_pinned.argtypes = [vp]
    L.dnv_pileup.argtypes = [vp, vp, vp, vp, vp, i64, vp, i64, vp]
    L.dnv_pileup_device.argtypes = [vp, vp, vp, vp, vp, i64, vp, i64, vp]
    L.dnv_infer_reads.argtypes = [vp, vp, vp, vp, vp, i64, vp, i64, vp, i64, C.POINTER(i64)]
    L.dnv_infer_reads_device.argtypes = [vp, vp, vp, vp, vp, i64, vp, i64, vp, i64, C.POINTER(i64)]
    L.dnv_mendel_filter.argtypes = [vp, vp, i64, vp, i64, C.POINTER(i64)]
    L.dnv_mendel_filter_device.argtypes = [vp, vp, i64, vp, i64, C.POINTER(i64)]
    L.dnv_synth_counts_device.argtypes = [vp, u64, i64, i64, vp, i32, u32, vp]
    L.dnv_synth_counts_at_device.argtypes = [vp, u64, vp, i64, vp, i32, u32, vp]
    L.dnv_synth_counts_at.argtypes = [vp, u64, vp, i64, vp, i32, u32, vp]
    L.dnv_measure_fp64_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.dnv_measure_screen.argtypes = [vp, vp, C.c_int, i64, C.c_int, C.POINTER(C.c_double)]
    for name in EXPORTS:
        if name != "dnv_last_error":
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc, allow=()):
    if rc != DNV_OK and rc not in allow:
        raise DenovoError(rc, load().dnv_last_error().decode("utf-8", "replace"))
    return rc


def ptr(a):
    """numpy array -> void*, torch tensor -> data_ptr, int passes through."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())
