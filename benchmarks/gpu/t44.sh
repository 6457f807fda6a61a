#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu_v3.json 2> gpurun_out/r02_bench_1gpu_v3.err
echo "rc=$?"; tail -3 gpurun_out/r02_bench_1gpu_v3.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_1gpu_v3.json'))
print(d['value'], d['ms_per_step'], d['config']['ms_per_step_synchronous_api'], d['e2e']['value'])
print(json.dumps(d['roofline'])[:900])
print(json.dumps(d['aux']['pileup_and_reads_to_calls']))
print(json.dumps(d['aux']['config2_mendel_filter']))
"
