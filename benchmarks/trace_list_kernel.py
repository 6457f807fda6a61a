"""Device-side timeline of trio_list_kernel (developer tool, needs a GPU).

Builds a copy of the library with -DDNV_LIST_TRACE (thread 0 of every CTA stamps %globaltimer at the
phase boundaries) into build/ and prints min / median / max over the CTAs.
    python benchmarks/trace_list_kernel.py            (under gpurun)
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "denovo-variant-caller-java_b200", "csrc")
LIB = os.path.join(ROOT, "build", "libdenovo_b200_trace.so")
if not os.path.exists(LIB) or "--rebuild" in sys.argv:
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                           "-Xcompiler", "-fPIC,-ffp-contract=off", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
                           "--expt-relaxed-constexpr", "-DDNV_LIST_TRACE", "-shared", "-o", LIB] +
                          [os.path.join(CSRC, f) for f in ("api.cu", "trio.cu", "synth.cu", "pileup.cu", "mendel.cu")] +
                          ["-x", "cu", os.path.join(CSRC, "tables.cpp")])
if "--build-only" in sys.argv:
    sys.exit(0)
import ctypes as C
os.environ["DNV_B200_LIB"] = LIB
sys.path.insert(0, os.path.join(ROOT, "denovo-variant-caller-java_b200", "python"))
import numpy as np, torch
import denovo_b200 as D
n = 10_000_000
eng = D.TrioEngine(1e-2, 1e-8, 1.0, "BAYES")
d = torch.empty((n, 12), dtype=torch.int16, device="cuda")
eng.synth_counts_device(0x5EED0003, 0, n, d, depth=30)
cap = 200000
calls = torch.zeros((cap, 80), dtype=torch.uint8, device="cuda")
for _ in range(5):
    eng.infer_device(d, n, calls, cap)
torch.cuda.synchronize()
G = 444
out = np.zeros((G, 8), np.uint64)
eng.lib.dnv_debug_list_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
print("rc", eng.lib.dnv_debug_list_trace(eng._h, out.ctypes.data_as(C.c_void_p), G))
t = out[:, :6].astype(np.int64)
t0 = t[:, 0].min()
rel = (t - t0) / 1000.0
names = ["start", "prologue_done", "first_tile_compute_done", "loop_done", None, "end"]
for i, nm in enumerate(names):
    if nm is None:
        continue
    col = rel[:, i]
    print("%-20s min %7.2f  med %7.2f  max %7.2f us" % (nm, col.min(), np.median(col), col.max()))
