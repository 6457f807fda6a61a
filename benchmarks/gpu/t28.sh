timeout 120 python benchmarks/time_pileup.py
timeout 120 python benchmarks/time_pileup.py 8000000
timeout 900 python -m pytest tests/test_gpu_pileup_mendel.py -m gpu -x -q -k pileup 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/pile_launches.csv python benchmarks/time_pileup.py 8000000 > gpurun_out/ncu8.log 2>&1
