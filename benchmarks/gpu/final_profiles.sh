# round-2 evidence run on ONE B200: bench line, launch list, ncu full captures (each after its plain run exited 0),
# device-side time lines.  Outputs under gpurun_out/; summarised into profiles/ on the build box.
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python bench.py --steps 100 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/r02_bench_1gpu_k100.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
export CUDA_DEVICE_MAX_CONNECTIONS=32
DNV_LANES=1 python benchmarks/ktrace.py 6 > gpurun_out/r02_ktrace_serial.txt 2>&1
python benchmarks/ktrace.py 18 > gpurun_out/r02_ktrace_3lanes.txt 2>&1
DNV_NO_LEVEL3=1 python benchmarks/ktrace.py 18 > gpurun_out/r02_ktrace_3lanes_no_level3.txt 2>&1
unset CUDA_DEVICE_MAX_CONNECTIONS
python bench.py --steps 4 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/plain_l.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 4 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python benchmarks/prof_read.py 6 > gpurun_out/plain_r.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"trio_list|trio_screen|trio_calls_finish" -s 9 -c 3 -o gpurun_out/r02_read_serial \
    python benchmarks/prof_read.py 6 > gpurun_out/ncu_r.log 2>&1
python benchmarks/bench_aux.py mendel > gpurun_out/plain_m.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"mendel" -s 6 -c 2 -o gpurun_out/r02_mendel \
    python benchmarks/bench_aux.py mendel > gpurun_out/ncu_m.log 2>&1
python benchmarks/time_pileup.py 8000000 > gpurun_out/plain_p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"pileup_pipe" -s 3 -c 1 -o gpurun_out/r02_pileup_pipe \
    python benchmarks/time_pileup.py 8000000 > gpurun_out/ncu_p.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 8 --csv --log-file gpurun_out/r02_launches_pileup.csv \
    python benchmarks/time_pileup.py 8000000 > gpurun_out/ncu_p2.log 2>&1
tail -2 gpurun_out/ncu_l.log gpurun_out/ncu_r.log gpurun_out/ncu_m.log gpurun_out/ncu_p.log gpurun_out/plain_p.log
