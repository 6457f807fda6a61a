#!/usr/bin/env python
"""Per-source-line stall samples of one kernel: joins `ncu --page source --csv` (SASS rows with sample counts)
with `nvdisasm --print-line-info` of the cubin inside the library (line of every SASS offset).
    python benchmarks/ncu_lines.py <report.ncu-rep> <kernel regex> <mangled name prefix> [top]"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
import os

rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "denovo-variant-caller-java_b200", "lib", "libdenovo_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
ins = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    on, cur = False, None
    for l in dis.splitlines():
        if l.startswith(".text."):
            on = l.startswith(".text." + mangled)
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins[int(m.group(1), 16)] = (cur, m.group(2).strip())
    if ins:
        break
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
base, seen = None, set()
lines, linstr = collections.Counter(), collections.Counter()
for r in rows[2:]:
    try:
        a = int(r[0], 16)
    except (ValueError, IndexError):
        continue
    if a in seen:
        continue
    seen.add(a)
    base = a if base is None else base
    cur = ins.get(a - base, (("?", 0), ""))[0] or ("?", 0)
    lines[cur] += int(r[ia])
    linstr[cur] += int(r[ii])
tot, toti = sum(lines.values()), sum(linstr.values())
print("%d samples, %d warp instructions, %d SASS instructions mapped" % (tot, toti, len(ins)))
src = {}
for k, v in lines.most_common(top):
    if k[0] not in src:
        p = os.path.join(ROOT, "denovo-variant-caller-java_b200", "csrc", k[0])
        src[k[0]] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[k[0]][k[1] - 1].strip()[:90] if 0 < k[1] <= len(src[k[0]]) else ""
    print("%5.1f%% samples %5.1f%% instr  %s:%d  %s" % (100 * v / tot, 100 * linstr[k] / toti, k[0], k[1], text))
