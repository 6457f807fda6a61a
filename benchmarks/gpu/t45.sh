#!/bin/bash
# late round-2 evidence of the final code on ONE B200
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 100 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/r02_bench_1gpu_k100_v3.json 2>/dev/null
export CUDA_DEVICE_MAX_CONNECTIONS=32
python benchmarks/ktrace.py 6 10000000 sync > gpurun_out/r02_ktrace_sync_call.txt 2>&1
python benchmarks/time_sync_call.py > gpurun_out/r02_sync_call.txt 2>&1
DNV_NO_HOST_COLLECT=1 python benchmarks/time_sync_call.py >> gpurun_out/r02_sync_call.txt 2>&1
DNV_SYNC_MIN_CHUNK=1048576 python benchmarks/time_sync_call.py >> gpurun_out/r02_sync_call.txt 2>&1
unset CUDA_DEVICE_MAX_CONNECTIONS
python benchmarks/time_pileup.py 8000000 > gpurun_out/plain_p3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 9 -c 6 --csv --log-file gpurun_out/r02_launches_pileup_v3.csv \
    python benchmarks/time_pileup.py 8000000 > gpurun_out/ncu_p3.log 2>&1
python bench.py --steps 4 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/plain_l3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 160 --csv --log-file gpurun_out/r02_launches_bench_v3.csv \
    python bench.py --steps 4 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/ncu_l3.log 2>&1
tail -2 gpurun_out/r02_sync_call.txt gpurun_out/plain_p3.log gpurun_out/ncu_p3.log gpurun_out/ncu_l3.log
tail -4 gpurun_out/r02_ktrace_sync_call.txt
