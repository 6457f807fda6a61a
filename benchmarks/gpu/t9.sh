python -m pytest tests/test_gpu_read.py -m gpu -x -q 2>&1 | grep -v "^$" | tail -30
DNV_LANES=3 DNV_SCREEN_CTAS=2 python benchmarks/ktrace.py 10
DNV_LANES=3 DNV_SCREEN_CTAS=3 python benchmarks/ktrace.py 10
DNV_LANES=4 DNV_SCREEN_CTAS=3 python benchmarks/ktrace.py 12
