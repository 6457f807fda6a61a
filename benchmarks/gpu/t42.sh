#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== tests"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== sync call"
python benchmarks/time_sync_call.py
DNV_NO_HOST_COLLECT=1 python benchmarks/time_sync_call.py
DNV_SYNC_MIN_CHUNK=1048576 python benchmarks/time_sync_call.py
python benchmarks/time_sync_call.py 40000000 100
DNV_SYNC_MIN_CHUNK=100000000 python benchmarks/time_sync_call.py 40000000 100
echo "== aux pileup (8 M)"
python - <<'PY'
import sys, json
sys.path.insert(0, '.'); sys.path.insert(0, 'denovo-variant-caller-java_b200/python')
import numpy as np, torch
import denovo_b200 as D, bench
print(json.dumps(bench.aux_pileup(D, torch, np, 0)))
PY
} > gpurun_out/t42.txt 2>&1
tail -30 gpurun_out/t42.txt
