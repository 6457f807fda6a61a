#!/bin/bash
# round-2 late changes: whole synchronous call in one chunk, pileup pre-passes folded, reads -> calls with one host wait
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_pileup_mendel.py tests/test_gpu_read.py -x -q -m gpu 2>&1 | tail -5
echo "== sync call time line"; python benchmarks/ktrace.py 6 10000000 sync | grep -v "^ *[0-9]"
echo "== pileup"; python benchmarks/time_pileup.py 8000000
echo "== aux pileup"; python benchmarks/bench_aux.py pileup
} > gpurun_out/t41.txt 2>&1
tail -40 gpurun_out/t41.txt
