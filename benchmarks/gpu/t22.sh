for s in 8 16 32; do DNV_PILEUP_S=$s python benchmarks/time_pileup.py; done
python -m pytest tests/test_gpu_pileup_mendel.py -m gpu -x -q 2>&1 | tail -3
