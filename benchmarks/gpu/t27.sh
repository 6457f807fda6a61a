timeout 120 python benchmarks/time_pileup.py
timeout 120 python benchmarks/time_pileup.py 8000000
timeout 900 python -m pytest tests/test_gpu_pileup_mendel.py -m gpu -x -q 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:pileup_pipe -s 3 -c 1 -o gpurun_out/pp8c python benchmarks/time_pileup.py 8000000 > gpurun_out/ncu8.log 2>&1
