python benchmarks/bench_aux.py mendel > gpurun_out/plain_m.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"mendel" -s 6 -c 2 -o gpurun_out/r02_mendel_b \
    python benchmarks/bench_aux.py mendel > gpurun_out/ncu_m.log 2>&1
tail -3 gpurun_out/plain_m.log
