#!/bin/bash
cd $GRAFT_REPO_ROOT
V=denovo-variant-caller-java_b200/lib/variants
for rep in 1 2; do
for name in ms0 ms40 ms100 ms250; do
  export DNV_B200_LIB=$PWD/$V/$name.so
  echo "$name $(timeout 120 python benchmarks/bench_aux.py mendel 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_pass'], d.get('bit_exact_vs_oracle_block'), d['position_calls'])")"
done; done 2>&1 | tee gpurun_out/t48.txt
