for s in 8 16; do DNV_PILEUP_S=$s timeout 120 python benchmarks/time_pileup.py; done
DNV_PILEUP_S=8 timeout 120 python benchmarks/time_pileup.py 8000000
timeout 600 python -m pytest tests/test_gpu_pileup_mendel.py -m gpu -x -q 2>&1 | tail -3
DNV_PILEUP_S=8 ncu --set full --clock-control none --import-source on -k regex:pileup_pipe -s 3 -c 1 -o gpurun_out/pp8b python benchmarks/time_pileup.py > gpurun_out/ncu8.log 2>&1
