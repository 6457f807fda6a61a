#!/bin/bash
cd $GRAFT_REPO_ROOT
V=denovo-variant-caller-java_b200/lib/variants
for rep in 1 2; do
for name in fb3 fb4 fb5; do
  export DNV_B200_LIB=$PWD/$V/$name.so
  echo "$name $(timeout 200 python bench.py --no-screen --no-cpu-baseline --no-aux --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['value']/1e9, d['ms_per_step'], r.get('kernel'), r.get('kernel_ms'), r.get('frac'))")"
done; done 2>&1 | tee gpurun_out/t50.txt
