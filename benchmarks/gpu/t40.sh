#!/bin/bash
# where does a synchronous call's time go: 3 lanes + graph, 3 lanes eager, 1 lane, 2 lanes
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== 3 lanes, graph"; python benchmarks/ktrace.py 6 10000000 sync
echo "== 3 lanes, eager"; DNV_GRAPHS=0 python benchmarks/ktrace.py 6 10000000 sync
echo "== 1 lane"; DNV_LANES=1 python benchmarks/ktrace.py 6 10000000 sync
echo "== 2 lanes, graph"; DNV_LANES=2 python benchmarks/ktrace.py 6 10000000 sync
} > gpurun_out/t40.txt 2>&1
tail -80 gpurun_out/t40.txt
