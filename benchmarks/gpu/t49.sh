#!/bin/bash
# final check of the round's last commit: smoke, the whole GPU suite, one bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -2
echo "== tests"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
} > gpurun_out/t49.txt 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu_v4.json 2> gpurun_out/r02_bench_1gpu_v4.err; echo "bench rc=$?" >> gpurun_out/t49.txt
tail -12 gpurun_out/t49.txt
