for k in stage pipe; do for s in 8 16; do DNV_PILEUP_KERNEL=$k DNV_PILEUP_S=$s timeout 120 python benchmarks/time_pileup.py; done; done
timeout 600 python -m pytest tests/test_gpu_pileup_mendel.py -m gpu -x -q 2>&1 | tail -3
