#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_pileup_mendel.py tests/test_host_cli.py -x -q -m gpu 2>&1 | tail -5
echo "== sync call, 100 M sites: three chunks (default) / whole"
python benchmarks/time_sync_call.py 100000000 30
DNV_SYNC_MIN_CHUNK=1000000000 python benchmarks/time_sync_call.py 100000000 30
python benchmarks/time_sync_call.py 67108864 30
DNV_SYNC_MIN_CHUNK=1000000000 python benchmarks/time_sync_call.py 67108864 30
echo "== aux pileup (8 M) and mendel"
python - <<'PY'
import sys, json
sys.path.insert(0, '.'); sys.path.insert(0, 'denovo-variant-caller-java_b200/python')
import numpy as np, torch
import denovo_b200 as D, bench
print(json.dumps(bench.aux_pileup(D, torch, np, 0)))
print(json.dumps(bench.aux_mendel(D, torch, np, 0)))
PY
} > gpurun_out/t43.txt 2>&1
tail -30 gpurun_out/t43.txt
