timeout 900 python -m pytest tests/test_gpu_pileup_mendel.py -m gpu -x -q -k "mendel" 2>&1 | tail -3
python benchmarks/bench_aux.py mendel 2>&1 | tail -1 | cut -c1-420
ncu --set full --clock-control none --import-source on -k regex:"mendel" -s 6 -c 2 -o gpurun_out/r02_mendel_c \
    python benchmarks/bench_aux.py mendel > gpurun_out/ncu_m.log 2>&1
