# round-2 evidence run on ONE B200: bench line, launch list, ncu full captures (each after its plain run exited 0)
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python bench.py --steps 4 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/plain_l.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 4 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python benchmarks/prof_read.py 6 > gpurun_out/plain_r.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"trio_list|trio_screen|trio_calls_finish" -s 9 -c 3 -o gpurun_out/r02_read_serial \
    python benchmarks/prof_read.py 6 > gpurun_out/ncu_r.log 2>&1
python benchmarks/bench_aux.py mendel > gpurun_out/plain_m.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"mendel" -s 6 -c 2 -o gpurun_out/r02_mendel \
    python benchmarks/bench_aux.py mendel > gpurun_out/ncu_m.log 2>&1
tail -2 gpurun_out/ncu_l.log gpurun_out/ncu_r.log gpurun_out/ncu_m.log
