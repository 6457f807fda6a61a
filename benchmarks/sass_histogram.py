#!/usr/bin/env python
"""SASS opcode histogram of the library's kernels (run on the CPU box: cuobjdump only), so that the instructions the
design leans on are on record: UBLKCP / SYNCS (TMA bulk copies + mbarriers), IDP (dp2a / dp4a dot products of the
screen), ATOMS / RED (shared-memory atomics of the scatter pileup), DFMA / DADD / DMUL (FP64 posterior), ACQBULK etc.
    python benchmarks/sass_histogram.py > profiles/r02_sass_opcodes.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "denovo-variant-caller-java_b200", "lib", "libdenovo_b200.so")
WANT = [("trio_screen_kernel<u16, 2 buffers>", "trio_screen_kernelILb0ELi2E"), ("trio_screen_kernel<u8, 2 buffers>", "trio_screen_kernelILb1ELi2E"),
        ("trio_list_kernel<u16, BAYES/LRT>", "trio_list_kernelILb0ELb0E"), ("trio_list_kernel<u16, MAP>", "trio_list_kernelILb0ELb1E"),
        ("trio_calls_finish_kernel<BAYES/LRT>", "trio_calls_finish_kernelILb0E"), ("trio_fast_kernel (all sites, BAYES)", "trio_fast_kernelILb1ELb0ELb0ELi3ELb0E"),
        ("pileup_pipe_kernel<8> (producer / consumer pipeline)", "pileup_pipe_kernelILi8E"), ("pileup_windows_kernel", "pileup_windows_kernel"),
        ("pileup_stage_kernel<8>", "pileup_stage_kernelILi8E"), ("pileup_scatter_kernel", "pileup_scatter_kernel"),
        ("pileup_screen_kernel (fused)", "pileup_screen_kernel"), ("mendel_admit_compact_kernel", "mendel_admit_compact_kernel"),
        ("mendel_join_emit_kernel", "mendel_join_emit_kernel"), ("gather_arrive_kernel", "gather_arrive_kernel"),
        ("gather_wait_kernel", "gather_wait_kernel")]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = {}, None
for l in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", l)
    if m and cur:
        funcs[cur][m.group(1)] += 1
print("# r02 — SASS opcode histograms (sm_100a, `cuobjdump -sass` of the committed sources' build)\n")
print("Made by `benchmarks/sass_histogram.py` on the CPU box. Static instruction counts per kernel (not executed counts).\n")
for title, key in WANT:
    hit = [f for f in funcs if key in f]
    if not hit:
        print("## %s\n\n(not found)\n" % title)
        continue
    c = funcs[hit[0]]
    tot = sum(c.values())
    print("## %s\n\n`%s` — %d instructions\n" % (title, hit[0][:110], tot))
    print(", ".join("%s %d" % (k, v) for k, v in c.most_common(28)))
    marks = {k: c[k] for k in ("UBLKCP", "SYNCS", "IDP", "ATOMS", "RED", "ATOMG", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "VOTE", "BAR", "ERRBAR", "MEMBAR", "ACQBULK", "CCTL") if c.get(k)}
    print("\nmarkers: " + ", ".join("%s %d" % kv for kv in marks.items()) + "\n")
